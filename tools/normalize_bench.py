"""Development timing of the preprocessing normalisation kernels against the HBM roofline."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from spotify_recommender_b200.engine import Engine

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
e = Engine(0)
g = torch.Generator(device="cuda").manual_seed(3)
d_raw = torch.rand((n, 11), device="cuda", generator=g)
d_genre = torch.randint(0, 114, (n,), device="cuda", dtype=torch.int32, generator=g)
d_out = torch.empty((n, 12), device="cuda")
d_mm = torch.empty(22, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    e.normalize_features_dev(d_raw, d_genre, n, 114, d_out, d_mm, stream=st)
torch.cuda.synchronize()
reps = 20
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ms = 0.0
for _ in range(reps):
    flush.zero_()  # L2 flush between repetitions
    t0.record()
    e.normalize_features_dev(d_raw, d_genre, n, 114, d_out, d_mm, stream=st)
    t1.record()
    torch.cuda.synchronize()
    ms += t0.elapsed_time(t1)
ms /= reps
alg = n * (44 + 44 + 4 + 48)  # min/max pass reads 44 B/song; normalise pass reads 48, writes 48
peak = 6458.7
try:
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbps", peak)
except Exception:
    pass
out = {"songs": n, "ms_per_call": ms, "algorithmic_bytes": alg, "achieved_gbps": alg / ms / 1e6, "peak_gbps": peak,
       "frac": alg / ms / 1e6 / peak, "note": "memset + minmax_kernel + normalize_kernel, CUDA events on the launching stream, L2 flushed between repetitions"}
# CPU oracle beside it (one thread, the reference's loop is OpenMP-parallel only in the second pass)
from oracle_lib import Oracle
o = Oracle()
m = min(n, 2_000_000)
raw = d_raw[:m].cpu().numpy()
gen = d_genre[:m].cpu().numpy()
t = time.perf_counter()
o.minmax_normalize(raw, gen, 114)
dt = time.perf_counter() - t
out["cpu_oracle_songs_per_s_1thread"] = m / dt
out["gpu_songs_per_s"] = n / (ms * 1e-3)
print(json.dumps(out))
