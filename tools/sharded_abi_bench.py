"""The single-process multi-GPU C ABI (sr_sharded_*) on every visible GPU: BASELINE config 4 (100 M songs row-sharded,
8192-query batches, top-100) through host buffers, checked against the oracle on sampled queries.  SR_SONGS overrides
the store size."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from oracle_lib import Oracle
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import ShardedEngine
n = int(float(os.environ.get("SR_SONGS", "1e8"))); nq, k = 8192, 100
devs = [int(x) for x in os.environ.get("SR_DEVICES", "").split(",") if x] or list(range(torch.cuda.device_count()))
f = synth.features(n)
se = ShardedEngine(devs)
t0 = time.perf_counter(); se.load_features(f); t_load = time.perf_counter() - t0
qs = [(((np.arange(nq, dtype=np.int64) + b * nq) * 7919 + 13) % n).astype(np.int32) for b in range(6)]
for b in range(2): se.query_by_index(qs[b], k)
t0 = time.perf_counter()
for b in range(2, 6): gi, gs = se.query_by_index(qs[b], k)
dt = (time.perf_counter() - t0) / 4
sel = np.unique(np.linspace(0, nq - 1, 32).astype(np.int64))
o = Oracle()
wi, ws = o.query_index(f, qs[5][sel], k, threads=len(os.sched_getaffinity(0)))
bad = int(((gi[sel] != wi).any(axis=1) | (gs[sel].view(np.uint32) != ws.view(np.uint32)).any(axis=1)).sum())
print(json.dumps({"config": f"sr_sharded_* (one process, {len(devs)} GPUs, peer-access gather + fused merge): {n} songs row-sharded, "
                            f"{nq}-query batches, top-{k}, host buffers in and out", "devices": devs, "ms_per_batch": dt * 1e3,
                  "song_pairs_per_s": float(n) * nq / dt, "load_seconds": t_load,
                  "parity_check": {"queries": int(sel.size), "mismatches": bad}}))
