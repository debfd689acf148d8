#!/usr/bin/env python
"""bench.py -- the contract benchmark of the cosine-similarity + top-K hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json metric "top-10 queries/s @10M songs & song-pairs/s"): a store of
10,000,000 synthetic Spotify-schema songs x 12 features PER GPU (row-sharded, weak scaling:
N GPUs hold N x 10M songs), batches of 4096 in-store query songs, top-10.  One "step" = one
pass of the hot path over one batch: every query scored against every song of the
(sharded) store, exact top-10 per query; at N > 1 the step includes the query-row
exchange, the NCCL all-gather of the per-shard candidates and the merge.

  value  song-pairs/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e    the same through the public host-buffer API: H2D of the query ids from pinned
         memory and D2H of the result lists inside the timed region
  roofline / cpu_baseline: see DESIGN.md "measurement"

`--impl reference` times the reference's own CPU implementation of the path
(oracle/_ref/libref_cpu.so = unmodified Recommender.cu built with -DDISABLE_CUDA; the
oracle port when that is absent) with all host threads on a bounded sample of the same
workload.  Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SONGS_PER_GPU = 10_000_000
BATCH = 4096
TOPK = 10
FLOP_PER_PAIR = 24          # 12 multiplies + 12 adds
BYTES_PER_SONG = 48         # 12 x FP32 per song per pass
METRIC = "song-pairs/s (top-10, 4096-query batches, 10M songs per GPU; queries/s @10M = value / 1e7)"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.2] or [l for (_, l) in self.lines]
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline_port(feats: np.ndarray, n_total: int, k: int, budget_s: float = 12.0) -> dict:
    """The oracle port (oracle/cosine_topk_oracle.c), OpenMP over all host cores, on a
    bounded sample of the batch: `cores` queries per round, rounds until ~budget_s."""
    from oracle_lib import Oracle
    from spotify_recommender_b200 import synth
    o = Oracle()
    cores = o.max_threads
    q = synth.query_indices(BATCH, n_total)
    q = q[q < feats.shape[0]][: max(cores, 8)]
    o.query_index(feats, q[:2], k, threads=cores)  # touch
    done, t0 = 0, time.perf_counter()
    while True:
        o.query_index(feats, q, k, threads=cores)
        done += q.size
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": done * float(feats.shape[0]) / dt, "unit": "song-pairs/s", "cores": cores, "kind": "port",
            "sample": f"{done} of the batch's {BATCH} queries x {feats.shape[0]} songs, top-{k}, "
                      f"oracle/cosine_topk_oracle.c with {cores} OpenMP threads, {dt:.1f} s"}


def run_reference(args) -> None:
    """--impl reference: the reference's own CPU implementation on the host cores."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle_lib import Oracle, Reference
    from spotify_recommender_b200 import synth
    n = SONGS_PER_GPU  # one GPU's shard of the workload; the CPU arm does not shard
    feats = synth.features(n)
    qall = synth.query_indices(BATCH, n)
    if Reference.available():
        ref = Reference(feats)
        cores = int(ref.L.ref_max_threads())
        kind = "reference"
        run = lambda q: ref.batch(q, TOPK, threads=cores)
        what = "oracle/_ref/libref_cpu.so (unmodified reference Recommender.cu, -DDISABLE_CUDA), one recommendByIndex per thread"
    else:
        o = Oracle()
        cores = o.max_threads
        kind = "port"
        run = lambda q: o.query_index(feats, q, TOPK, threads=cores)
        what = "oracle/cosine_topk_oracle.c (oracle/_ref absent)"
    per_step = max(cores, 8)
    for w in range(args.warmup):
        run(qall[w * per_step:(w + 1) * per_step])
    t0 = time.perf_counter()
    for s in range(args.steps):
        lo = ((args.warmup + s) * per_step) % (BATCH - per_step)
        run(qall[lo:lo + per_step])
    dt = time.perf_counter() - t0
    value = args.steps * per_step * float(n) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "song-pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n} songs x 12, top-{TOPK}; each step = {per_step} of the batch's {BATCH} queries "
                               f"(bounded sample), {what}", "songs": n, "queries_per_step": per_step, "top_k": TOPK},
        "queries_per_s_at_10M": value / 1e7,
        "cpu_baseline": {"value": value, "unit": "song-pairs/s", "cores": cores, "kind": kind,
                         "sample": f"{per_step} queries per step x {n} songs, {cores} host threads"},
        "e2e": {"value": value, "unit": "song-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args) -> None:
    global BATCH, TOPK
    if args.batch:
        BATCH = args.batch
    if args.topk:
        TOPK = args.topk
    import torch
    import torch.distributed as dist
    from spotify_recommender_b200 import synth
    from spotify_recommender_b200.engine import Engine, variant_names
    from spotify_recommender_b200.sharded import ShardedRecommender, shard_bounds

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with torch.distributed.run --nproc-per-node N (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_total = int(args.songs_total) if args.songs_total else SONGS_PER_GPU * world
    eng = Engine(local_rank)
    if args.variant is not None:
        eng.set_option("variant", args.variant)
    sh = ShardedRecommender(eng, n_total, device=dev)
    lo, hi = shard_bounds(n_total, world, rank)
    feats = synth.features(n_total, lo, hi)          # this rank's row shard, generated on the host
    sh.load_shard(feats)
    n_batches = args.warmup + args.steps
    q_host = [((np.arange(BATCH, dtype=np.int64) + b * BATCH) * 7919 + 13) % n_total for b in range(n_batches)]
    q_host = [q.astype(np.int32) for q in q_host]
    q_dev = [torch.from_numpy(q).to(dev) for q in q_host]
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ("value") -----------------------------------------------
    for b in range(args.warmup):
        sh.query_by_index_dev(q_dev[b], TOPK)
    barrier()
    eng.set_option("profile", 1)
    eng.set_option("reset", 1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    barrier()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for b in range(args.warmup, n_batches):
        out_i, out_s = sh.query_by_index_dev(q_dev[b], TOPK)
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t_wall0, t_wall1)
    launches = eng.stat("kernel_launches")
    scan_ms, scan_n = eng.timing("scan")
    other = {k: eng.timing(k)[0] / args.steps for k in ("prep", "sample", "bound", "finalize", "merge")}
    eng.set_option("profile", 0)
    checksum = int(out_i.to(torch.int64).sum().item())  # the step's result is really read

    # ---- end to end through the host-buffer API ("e2e") -----------------------------------
    sh.query_by_index(q_host[0], TOPK)
    barrier()
    t0 = time.perf_counter()
    for b in range(args.warmup, n_batches):
        h_i, h_s = sh.query_by_index(q_host[b], TOPK)
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    barrier()
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e2e.item())
    assert int(h_i.astype(np.int64).sum()) == checksum, "host-path and device-path results differ"

    pairs_per_step = float(BATCH) * float(n_total)
    value = pairs_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = pairs_per_step * args.steps / e2e_s

    # ---- roofline of the dominant kernel (scan): FP32-bound at Q = 4096 ----------
    props = torch.cuda.get_device_properties(dev)
    sm_count = props.multi_processor_count
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
    fp32_peak = sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12          # FFMA lanes x 2 flop x max clock
    pairs_per_launch = float(hi - lo) * BATCH * args.steps / max(scan_n, 1)
    scan_avg_ms = scan_ms / max(scan_n, 1)
    achieved = FLOP_PER_PAIR * pairs_per_launch / (scan_avg_ms * 1e-3) / 1e12
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "scan_kernel_summary.json")) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    except Exception:
        pass
    kernel_shape = variant_names()[eng.stat("variant")]  # of the timed batches (the HBM sweep below uses another)
    roofline = {
        "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
        "traffic": traffic, "kernel": "scan_kernel<%s>" % kernel_shape,
        "launches_timed": scan_n, "avg_launch_ms": scan_avg_ms,
        "peak_is": f"{sm_count} SMs x 128 FFMA lanes x 2 flop x {sm_max_mhz:.0f} MHz (nominal FP32, no tensor cores; "
                   "north_star: min(HBM, FP32) roofline; this batch is FP32-bound, Q* = 23)",
        "peak_measured_ffma2": eng.measure_fp32(1),
        "algorithmic": f"{FLOP_PER_PAIR} flop/pair x {pairs_per_launch:.3e} pairs per launch",
        "other_kernels_ms_per_step": other,
    }
    # the HBM-bound regime of the same kernels (Q < Q* = 23 queries: one pass over the store per call)
    hbm = None
    if rank == 0 and world == 1:
        try:
            hbm_peak = float(peaks.get("hbm_gbs") or 6650.0)
            hbm = {"bound": "hbm", "peak": hbm_peak, "unit": "GB/s",
                   "peak_is": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6650 GB/s",
                   "algorithmic": f"{BYTES_PER_SONG} B/song x {hi - lo} songs per call", "cases": []}
            for nq_small in (1, 4, 16):
                qs = q_dev[0][:nq_small].contiguous()
                os_ = torch.empty((nq_small, TOPK), dtype=torch.int32, device=dev)
                for _ in range(3):
                    eng.query_by_index_dev(qs, nq_small, TOPK, os_, None, stream.cuda_stream)
                torch.cuda.synchronize()
                eng.set_option("profile", 1)
                eng.set_option("reset", 1)
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(20):
                    eng.query_by_index_dev(qs, nq_small, TOPK, os_, None, stream.cuda_stream)
                b_.record(stream)
                torch.cuda.synchronize()
                t_call = a.elapsed_time(b_) / 20 * 1e-3
                t_scan = eng.timing("scan")[0] / 20 * 1e-3
                eng.set_option("profile", 0)
                nbytes = BYTES_PER_SONG * float(hi - lo)
                hbm["cases"].append({"queries": nq_small, "ms_per_call": t_call * 1e3, "scan_kernel_ms": t_scan * 1e3,
                                     "achieved_call": nbytes / t_call / 1e9, "frac_call": nbytes / t_call / 1e9 / hbm_peak,
                                     "achieved_scan_kernel": nbytes / t_scan / 1e9,
                                     "frac_scan_kernel": nbytes / t_scan / 1e9 / hbm_peak})
        except Exception as exc:  # never lose the headline over the side measurement
            hbm = {"error": str(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_port(feats, n_total, TOPK)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "song-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{hi - lo} songs x 12 features per GPU ({n_total} total, row-sharded), "
                                   f"batch of {BATCH} in-store queries, exact top-{TOPK}",
                       "songs_total": n_total, "songs_per_gpu": hi - lo, "queries_per_batch": BATCH,
                       "top_k": TOPK, "parallelism": f"row-shard x{world}" + (" + NCCL all-gather + merge" if world > 1 else ""),
                       "l2": "store (2 x 480 MB per GPU) is larger than the 126 MB L2; every step uses a fresh query batch",
                       "kernel_shape": kernel_shape},
            "queries_per_s_at_10M": value / 1e7,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "song-pairs/s", "h2d_bytes_per_step": BATCH * 4,
                    "d2h_bytes_per_step": BATCH * TOPK * 8, "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "roofline_hbm_regime": hbm,
            "cpu_baseline": cpu,
            "result_checksum": checksum,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--variant", type=int, default=None, help="scan kernel shape (development)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--songs-total", type=float, default=None,
                    help="non-contract runs: total songs, row-sharded over the GPUs (e.g. 1e8 for BASELINE config 4)")
    ap.add_argument("--batch", type=int, default=None, help="non-contract runs: queries per batch")
    ap.add_argument("--topk", type=int, default=None, help="non-contract runs: K")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
