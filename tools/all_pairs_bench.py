"""BASELINE config 5 on one GPU: all-pairs top-10 neighbour table for 1M songs (10^12 scored pairs), host table out.

    python tools/all_pairs_bench.py [songs] [key=value engine options ...]"""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import Engine
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
k = 10
f = synth.features(n)
e = Engine(0); e.load_features(f)
opts = {}
for kv in sys.argv[2:]:
    if "=" in kv:
        opts[kv.split("=")[0]] = int(kv.split("=")[1])
        e.set_option(kv.split("=")[0], int(kv.split("=")[1]))
e.all_pairs_topk(0, 20000, k)  # warm-up
t0 = time.perf_counter()
gi, gs = e.all_pairs_topk(0, n, k)
dt = time.perf_counter() - t0
assert (gi >= 0).all() and (gi != np.arange(n)[:, None]).all()
print(json.dumps({"config": f"all-pairs top-{k}, {n} songs, 1 GPU, host table out", "options": opts, "seconds": dt,
                  "song_pairs_per_s": float(n) * n / dt, "queries_per_s": n / dt,
                  "frac_of_fp32_roofline": 24.0 * n * n / dt / 1e12 / 74.45}))
