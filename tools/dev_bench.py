"""Development timing sweep over kernel shapes (not the contract bench: see bench.py)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import Engine, variant_names

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [10, 100]
variants = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else list(range(len(variant_names())))
vname = lambda v: "auto" if v < 0 else variant_names()[v]
import torch
t0 = time.time()
import os
if os.environ.get("SR_DATA") == "synth":
    feats = synth.features(n)
else:
    g = torch.Generator(device="cuda").manual_seed(1)
    feats = torch.rand((n, 12), device="cuda", generator=g)
    feats = torch.floor(feats * 1000) / 1000
e = Engine(0)
e.load_features(feats)
import os
for kv in os.environ.get("SR_OPTS", "").split(","):
    if "=" in kv:
        e.set_option(kv.split("=")[0], int(kv.split("=")[1]))
print("store ready", time.time() - t0, "s; fp32 peak FFMA2 %.1f FFMA %.1f unfused %.1f TF" % (e.measure_fp32(1), e.measure_fp32(0), e.measure_fp32(2)), flush=True)
q = synth.query_indices(nq, n)
dq = torch.from_numpy(q).cuda()
for k in ks:
    oi = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    os_ = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    for v in variants:
        e.set_option("variant", v)
        e.set_option("profile", 1)
        for it in range(3):
            if it == 1:
                e.set_option("reset", 1)
            e.query_by_index_dev(dq, nq, k, oi, os_)
        e.synchronize()
        ms, cnt = e.timing("scan")
        tot = sum(e.timing(x)[0] for x in ("prep", "sample", "bound", "scan", "finalize")) / 2
        pairs = float(n) * nq
        tf = pairs * 24 / (ms / 2 * 1e-3) / 1e12
        print("k=%d %-18s scan %.3f ms/batch (%d launches) total %.3f ms  scan %.1f TFLOP/s = %.1f%% of 74.4; hits/q %.0f settles/q %.2f rescans %d rescored/q %.0f refilters %d" % (
            k, vname(v), ms / 2, cnt // 2, tot, tf, 100 * tf / 74.4, e.stat("filter_hits") / 2 / nq, e.stat("settles") / 2 / nq,
            e.stat("rescans"), e.stat("rescored") / 2 / nq, e.stat("refilters")), "| bound %.3f sample %.3f finalize %.3f prep %.3f" % (e.timing("bound")[0] / 2, e.timing("sample")[0] / 2, e.timing("finalize")[0] / 2, e.timing("prep")[0] / 2), flush=True)
        if e.stat("cta_cycles"):
            cta, hot, st, wt = e.stat("cta_cycles"), e.stat("hot_cycles"), e.stat("settle_cycles"), e.stat("wait_cycles")
            print("   in-kernel cycles (thread 0 of every CTA): hot loop %.1f%%  settle phases %.1f%%  tile barrier wait %.1f%%  other (prologue, flush, joins) %.1f%%; CTA-cycles/launch/SM %.0f" % (
                100.0 * hot / cta, 100.0 * st / cta, 100.0 * wt / cta, 100.0 * (cta - hot - st - wt) / cta, cta / max(1, cnt) / 148), flush=True)
            print("   of which: " + "  ".join("%s %.1f%%" % (k, 100.0 * e.stat(k + "_cycles") / cta) for k in ("prologue", "final_settle", "flush", "join")), flush=True)
