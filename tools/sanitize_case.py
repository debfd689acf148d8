"""Small end-to-end case for compute-sanitizer runs (memcheck / racecheck): every kernel shape, the long-list and
ceiling passes, the device-pointer API with a foreign id, the sharded path (three shards on this GPU) and all-pairs."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle_lib import Oracle
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import Engine, ShardedEngine, variant_names
o = Oracle()
e = Engine(0)
for n, nq, k in ((60_000, 70, 10), (30_011, 9, 150), (5000, 33, 7), (20_000, 300, 50)):
    f = synth.adversarial(n) if n == 30_011 else synth.features(n)
    e.load_features(f)
    q = synth.query_indices(nq, n)
    want = o.query_index(f, q, k, threads=4)
    for v in [-1] + list(range(len(variant_names()))):
        e.set_option("variant", v)
        gi, gs = e.query_by_index(q, k)
        assert np.array_equal(gi, want[0]) and np.array_equal(gs.view(np.uint32), want[1].view(np.uint32)), (n, nq, k, v)
    e.set_option("variant", -1)
f = synth.features(4000)
e.load_features(f)
q = np.array([1, 77, 3999], np.int32)
gi, gs = e.query_by_index(q, 2500)  # three passes under ceilings
want = o.query_index(f, q, 2500, threads=4)
assert np.array_equal(gi, want[0]) and np.array_equal(gs.view(np.uint32), want[1].view(np.uint32))
gi, gs = e.all_pairs_topk(0, 4000, 5)
want = o.query_index(f, np.arange(4000, dtype=np.int32), 5, threads=4)
assert np.array_equal(gi, want[0])
with ShardedEngine([0, 0, 0]) as se:
    se.load_features(f)
    gi, gs = se.query_by_index(q, 20)
    want = o.query_index(f, q, 20, threads=4)
    assert np.array_equal(gi, want[0]) and np.array_equal(gs.view(np.uint32), want[1].view(np.uint32))
print("sanitize case ok")
