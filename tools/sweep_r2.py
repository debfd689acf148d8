"""Development sweep (not the contract bench): time per call and per scan launch over batch sizes, kernel shapes and
store sizes, to see where the mid-size batches lose their time.  `SR_ENGINE_SO=<path>` selects another build of the
engine (e.g. one made with SR_NVCC_EXTRA=-DSR_SCAN_TIMING)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import Engine, variant_names

sizes = [int(float(x)) for x in (sys.argv[1] if len(sys.argv) > 1 else "1e7,1e6").split(",")]
nqs = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,4,16,32,40,64,128,256,512,1024,1280").split(",")]
ks = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "10").split(",")]
variants = [int(x) for x in (sys.argv[4] if len(sys.argv) > 4 else "-1,4,5").split(",")]
opts = {}
for kv in (sys.argv[5] if len(sys.argv) > 5 else "").split(","):
    if "=" in kv:
        opts[kv.split("=")[0]] = int(kv.split("=")[1])
names = variant_names()
e = Engine(0)
peak = 148 * 128 * 2 * 1.965e9 / 1e12
for n in sizes:
    f = synth.features(n)
    e.load_features(f)
    for key, v in opts.items():
        e.set_option(key, v)
    for k in ks:
        for nq in nqs:
            q = torch.from_numpy(synth.query_indices(nq, n)).cuda()
            oi = torch.empty((nq, k), dtype=torch.int32, device="cuda")
            for v in variants:
                try:
                    e.set_option("variant", v)
                    reps = 30 if nq <= 64 else 10
                    for _ in range(3):
                        e.query_by_index_dev(q, nq, k, oi, None, 0)
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(reps):
                        e.query_by_index_dev(q, nq, k, oi, None, 0)
                    b.record()
                    torch.cuda.synchronize()
                    call_ms = a.elapsed_time(b) / reps
                    e.set_option("profile", 1); e.set_option("reset", 1)
                    for _ in range(reps):
                        e.query_by_index_dev(q, nq, k, oi, None, 0)
                    e.synchronize(); torch.cuda.synchronize()
                    scan_ms = e.timing("scan")[0] / reps
                    other = {x: round(e.timing(x)[0] / reps * 1e3, 1) for x in ("prep", "bound", "sample", "finalize")}
                    e.set_option("profile", 0)
                    t_fp = 24.0 * n * nq / (peak * 1e12) * 1e3
                    t_hbm = 48.0 * n / 6458.7e9 * 1e3
                    roof = max(t_fp, t_hbm)
                    rec = {"n": n, "nq": nq, "k": k, "variant": "auto" if v < 0 else names[v], "used": names[e.stat("variant")],
                           "call_ms": round(call_ms, 4), "scan_ms": round(scan_ms, 4), "roof_ms": round(roof, 4),
                           "frac_call": round(roof / call_ms, 3), "frac_scan": round(roof / scan_ms, 3), "other_us": other,
                           "hits_per_q": round(e.stat("filter_hits") / reps / nq, 1)}
                    if e.stat("cta_cycles"):
                        cta = e.stat("cta_cycles")
                        rec["cycles_pct"] = {x: round(100.0 * e.stat(x + "_cycles") / cta, 1) for x in ("hot", "settle", "wait", "prologue", "final_settle", "flush", "join")}
                    print(json.dumps(rec), flush=True)
                except Exception as exc:
                    print(json.dumps({"n": n, "nq": nq, "k": k, "variant": v, "error": str(exc)}), flush=True)
            e.set_option("variant", -1)
