"""Small end-to-end case for compute-sanitizer runs (memcheck / racecheck), all kernel shapes."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle_lib import Oracle
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import Engine, variant_names
o = Oracle()
e = Engine(0)
for n, nq, k in ((60_000, 70, 10), (30_011, 9, 150), (5000, 33, 7)):
    f = synth.adversarial(n) if n == 30_011 else synth.features(n)
    e.load_features(f)
    q = synth.query_indices(nq, n)
    want = o.query_index(f, q, k, threads=4)
    for v in [-1] + list(range(len(variant_names()))):
        e.set_option("variant", v)
        gi, gs = e.query_by_index(q, k)
        assert np.array_equal(gi, want[0]) and np.array_equal(gs.view(np.uint32), want[1].view(np.uint32)), (n, nq, k, v)
    e.set_option("variant", -1)
print("sanitize case ok")
