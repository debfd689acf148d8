"""This engine vs the reference's own cuBLAS SGEMV path rebuilt for sm_100a (oracle/_ref/libref_gpu.so: the unmodified
Recommender.cu; its batch mode is sequential recommendByIndex calls), host buffers in and out, same box, same data.

    python tools/ref_gpu_compare.py [songs] [queries] [reference_sample]

BASELINE config 2 is 1e6 songs x 1024 queries; a 1e7 spot check times a few reference queries only."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle_lib import Reference
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import Engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
sample = int(sys.argv[3]) if len(sys.argv) > 3 else nq
k = 10
f = synth.features(n)
q = synth.query_indices(nq, n)
eng = Engine(0); eng.load_features(f)
for _ in range(3): gi, gs = eng.query_by_index(q, k)
t0 = time.perf_counter()
for _ in range(10): gi, gs = eng.query_by_index(q, k)
t_ours = (time.perf_counter() - t0) / 10
for _ in range(3): eng.query_by_index(q[:1], k)
t0 = time.perf_counter()
for i in range(50): eng.query_by_index(q[i:i + 1], k)
t_one = (time.perf_counter() - t0) / 50
ref = Reference(f, gpu=True)
assert ref.gpu_enabled(), "reference fell back to its CPU path"
ref.batch(q[:8], k)
t0 = time.perf_counter()
ri = ref.batch(q[:sample], k)
t_ref = (time.perf_counter() - t0) / sample
same = int((ri == gi[:sample]).all(axis=1).sum())
print(json.dumps({"config": f"{n} songs x 12, {nq} queries, top-{k}, host buffers in and out",
                  "ours_ms_per_batch": t_ours * 1e3, "ours_ms_single_query_call": t_one * 1e3,
                  "reference_cublas_path_ms_per_query": t_ref * 1e3,
                  "reference_cublas_path_ms_per_batch": t_ref * 1e3 * nq, "reference_sample": sample,
                  "speedup_per_batch": t_ref * nq / t_ours, "speedup_single_query": t_ref / t_one,
                  "identical_ordered_lists": f"{same}/{sample}"}))
