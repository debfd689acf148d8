import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import Engine
from oracle_lib import Oracle
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
f = synth.features(n)
e = Engine(0); e.load_features(f)
q = ((np.arange(nq, dtype=np.int64) * 7919 + 13) % n).astype(np.int32)
ref = None
for it in range(6):
    gi, gs = e.query_by_index(q, k)
    if ref is None:
        ref = (gi.copy(), gs.copy())
    else:
        bad = np.where((gi != ref[0]).any(axis=1))[0]
        print("run", it, "rows differing from run 0:", bad.size, bad[:10])
        for b in bad[:3]:
            print("  q", b, q[b], gi[b], ref[0][b], gs[b], ref[1][b])
o = Oracle()
sel = np.arange(0, nq, max(1, nq // 48))
wi, ws = o.query_index(f, q[sel], k, threads=o.max_threads)
bad = np.where((wi != ref[0][sel]).any(axis=1))[0]
print("vs oracle: mismatching rows", bad.size, "of", sel.size)
for b in bad[:5]:
    print("  q", sel[b], q[sel[b]], "got", ref[0][sel[b]], "want", wi[b], ref[1][sel[b]], ws[b])
print("stats: hits", e.stat("filter_hits"), "settles", e.stat("settles"), "rescans", e.stat("rescans"), "refilters", e.stat("refilters"))
