"""BASELINE config 2: 1M songs x 12, batch of 1024 queries, top-10 -- this engine vs the
reference's own cuBLAS SGEMV path rebuilt for sm_100a (oracle/_ref/libref_gpu.so: the
unmodified Recommender.cu; its batch mode is 1024 sequential recommendByIndex calls)."""
import sys, time, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle_lib import Reference
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import Engine

n, nq, k = 1_000_000, 1024, 10
f = synth.features(n)
q = synth.query_indices(nq, n)
eng = Engine(0); eng.load_features(f)
for _ in range(3): gi, gs = eng.query_by_index(q, k)
t0 = time.perf_counter()
for _ in range(10): gi, gs = eng.query_by_index(q, k)
t_ours = (time.perf_counter() - t0) / 10
ref = Reference(f, gpu=True)
assert ref.gpu_enabled(), "reference fell back to its CPU path"
ref.batch(q[:32], k)
t0 = time.perf_counter()
ri = ref.batch(q, k)
t_ref = time.perf_counter() - t0
same = int((ri == gi).all(axis=1).sum())
print(json.dumps({"config": "1M songs x 12, 1024 queries, top-10, host buffers in and out",
                  "ours_ms_per_batch": t_ours * 1e3, "reference_cublas_path_ms_per_batch": t_ref * 1e3,
                  "speedup": t_ref / t_ours, "identical_ordered_lists": f"{same}/{nq}"}))
