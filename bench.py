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

Every case is followed by a parity check: 32 sampled queries of the last timed batch are compared bit for bit
(index lists and score bits) with the CPU oracle -- at N > 1 every rank runs the oracle over its own row shard
and rank 0 merges the parts -- and reported as "parity_check": {"queries": 32, "mismatches": 0}.

Extra keys on the same line (DESIGN.md "measurement"):
  config3_top100   (N = 1) the same store and batch at top-100 = BASELINE config 3 proper
  strong_scaling   BASELINE config 4 at every N, including N = 1: 100 M songs TOTAL row-sharded over the N GPUs,
                   batches of 8192 queries, top-100 ("scaling": "strong"; T_1 / (N T_N) is the efficiency)
  all_pairs_1M     BASELINE config 5 at every N: the top-10 neighbour table of 1 M songs, store replicated, queries sharded
  roofline_hbm_regime, reference_gpu_path (N = 1): the HBM-bound regime of the same kernels, and the
                   reference's own cuBLAS path rebuilt for sm_100a (oracle/_ref/libref_gpu.so) on BASELINE config 2

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


C4_SONGS_TOTAL = 100_000_000   # BASELINE config 4: 100 M songs over the GPUs, 8192 queries, top-100
C4_BATCH = 8192
C4_TOPK = 100
PARITY_QUERIES = 32


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_cores() -> int:
    """Threads this process may use -- NOT OpenMP's default, which torchrun pins to 1 through OMP_NUM_THREADS."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


def cpu_baseline_port(feats: np.ndarray, n_total: int, k: int, batch: int, budget_s: float = 12.0) -> dict:
    """The oracle port (oracle/cosine_topk_oracle.c), OpenMP over all host cores, on a
    bounded sample of the batch: `cores` queries per round, rounds until ~budget_s."""
    from oracle_lib import Oracle
    from spotify_recommender_b200 import synth
    o = Oracle()
    cores = host_cores()
    q = synth.query_indices(batch, n_total)
    q = q[q < feats.shape[0]][: max(cores, 8)]
    o.query_index(feats, q[:2], k, threads=cores)  # touch
    done, t0 = 0, time.perf_counter()
    while True:
        o.query_index(feats, q, k, threads=cores)
        done += q.size
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    o.query_index(feats, q[:2], k, threads=1)  # the same code on ONE host thread (both are reported)
    dt1 = time.perf_counter() - t1
    return {"value": done * float(feats.shape[0]) / dt, "unit": "song-pairs/s", "cores": cores, "kind": "port",
            "sample": f"{done} of the batch's {batch} queries x {feats.shape[0]} songs, top-{k}, "
                      f"oracle/cosine_topk_oracle.c with {cores} OpenMP threads, {dt:.1f} s",
            "value_single_thread": 2 * float(feats.shape[0]) / dt1, "single_thread_sample": f"2 queries x {feats.shape[0]} songs, 1 thread, {dt1:.2f} s"}


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
    cores = host_cores()  # the same at every N: torchrun's OMP_NUM_THREADS=1 does not apply to an explicit thread count
    if Reference.available():
        ref = Reference(feats)
        kind = "reference"
        run = lambda q: ref.batch(q, TOPK, threads=cores)
        what = "oracle/_ref/libref_cpu.so (unmodified reference Recommender.cu, -DDISABLE_CUDA), one recommendByIndex per thread"
    else:
        o = Oracle()
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


class Job:
    """One rank's view of the run: device, process group, engine."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from spotify_recommender_b200.engine import Engine
        self.torch, self.dist = torch, dist
        self.world = env_int("WORLD_SIZE", 1)
        self.rank = env_int("RANK", 0)
        self.local_rank = env_int("LOCAL_RANK", 0)
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        if args.gpus > 1 and self.world == 1:
            raise SystemExit("for --gpus N > 1 launch with torch.distributed.run --nproc-per-node N (one rank per GPU)")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.eng = Engine(self.local_rank)
        if args.variant is not None:
            self.eng.set_option("variant", args.variant)
        self.stream = torch.cuda.current_stream()
        props = torch.cuda.get_device_properties(self.dev)
        self.sm_count = props.multi_processor_count
        self.peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                self.peaks = json.load(fh)
        except Exception:
            pass

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def parity_check(job: Job, feats, n_total: int, lo: int, hi: int, q_last: np.ndarray, out_i, out_s, k: int) -> dict:
    """PARITY_QUERIES sampled queries of the last timed batch against the CPU oracle, bit for bit.  Every rank
    scores its own row shard with the oracle (all its share of the host cores); rank 0 merges the parts in the
    oracle's own order (sr_oracle_merge_parts) and compares with what the GPUs returned."""
    from oracle_lib import Oracle
    from spotify_recommender_b200 import synth
    torch, dist = job.torch, job.dist
    o = Oracle()
    nq = q_last.size
    sel = np.unique(np.linspace(0, nq - 1, PARITY_QUERIES).astype(np.int64))
    ids = q_last[sel].astype(np.int64)
    qrows = synth.rows(ids, n_total)
    ex = np.where((ids >= lo) & (ids < hi), ids - lo, -1).astype(np.int64)
    threads = max(1, host_cores() // max(1, job.world))
    t0 = time.perf_counter()
    li, ls = o.query_rows(feats, qrows, ex, k, id_base=lo, threads=threads)
    dt = time.perf_counter() - t0
    if job.world > 1:
        d_i, d_s = torch.from_numpy(li).to(job.dev), torch.from_numpy(ls).to(job.dev)
        all_i = torch.empty((job.world,) + tuple(d_i.shape), dtype=torch.int32, device=job.dev)
        all_s = torch.empty((job.world,) + tuple(d_s.shape), dtype=torch.float32, device=job.dev)
        dist.all_gather_into_tensor(all_i.view(-1, k), d_i)
        dist.all_gather_into_tensor(all_s.view(-1, k), d_s)
        wi, ws = o.merge_parts(all_i.cpu().numpy(), all_s.cpu().numpy())
    else:
        wi, ws = li, ls
    gi = out_i[torch.from_numpy(sel).to(job.dev)].cpu().numpy()
    gs = out_s[torch.from_numpy(sel).to(job.dev)].cpu().numpy()
    bad = int(((gi != wi).any(axis=1) | (gs.view(np.uint32) != ws.view(np.uint32)).any(axis=1)).sum())
    return {"queries": int(sel.size), "mismatches": bad, "top_k": k, "songs": n_total,
            "oracle": f"oracle/cosine_topk_oracle.c over {'each rank its own row shard, parts merged on rank 0' if job.world > 1 else 'the whole store'}, "
                      f"{threads} threads per rank, {dt:.2f} s; index lists and score bits compared"}


def measure_case(job: Job, args, n_total: int, batch: int, topk: int, steps: int, warmup: int, with_e2e: bool = True,
                 sample_clocks: bool = False):
    """Load this rank's row shard of an n_total-song store and time `steps` batches.  Returns (result dict, feats)."""
    from spotify_recommender_b200 import synth
    from spotify_recommender_b200.engine import variant_names
    from spotify_recommender_b200.sharded import ShardedRecommender, shard_bounds
    torch, eng = job.torch, job.eng
    sh = ShardedRecommender(eng, n_total, device=job.dev)
    lo, hi = shard_bounds(n_total, job.world, job.rank)
    feats = synth.features(n_total, lo, hi)          # this rank's row shard, generated on the host
    sh.load_shard(feats)
    n_batches = warmup + steps
    q_host = [(((np.arange(batch, dtype=np.int64) + b * batch) * 7919 + 13) % n_total).astype(np.int32) for b in range(n_batches)]
    q_dev = [torch.from_numpy(q).to(job.dev) for q in q_host]

    # ---- device-resident timing ("value") -----------------------------------------------
    for b in range(warmup):
        sh.query_by_index_dev(q_dev[b], topk)
    job.barrier()
    eng.set_option("profile", 1)
    eng.set_option("reset", 1)
    sampler = ClockSampler(job.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    job.barrier()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(job.stream)
    for b in range(warmup, n_batches):
        out_i, out_s = sh.query_by_index_dev(q_dev[b], topk)
    ev1.record(job.stream)
    job.barrier()
    t_wall1 = time.time()
    ms_total = job.max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    launches = eng.stat("kernel_launches")
    scan_ms, scan_n = eng.timing("scan")
    other = {k: eng.timing(k)[0] / steps for k in ("prep", "sample", "bound", "finalize", "merge")}
    eng.set_option("profile", 0)
    out_i, out_s = out_i.clone(), out_s.clone()
    checksum = int(out_i.to(torch.int64).sum().item())  # the step's result is really read
    kernel_shape = variant_names()[eng.stat("variant")]

    res = {"songs_total": n_total, "songs_per_gpu": hi - lo, "queries_per_batch": batch, "top_k": topk, "steps": steps,
           "warmup": warmup, "ms_per_step": ms_total / steps, "value": float(batch) * float(n_total) * steps / (ms_total * 1e-3),
           "unit": "song-pairs/s", "gpu_launches": launches, "kernel_shape": kernel_shape, "result_checksum": checksum,
           "clocks": clocks}

    # ---- end to end through the host-buffer API ("e2e") -----------------------------------
    if with_e2e:
        sh.query_by_index(q_host[0], topk)
        job.barrier()
        t0 = time.perf_counter()
        for b in range(warmup, n_batches):
            h_i, h_s = sh.query_by_index(q_host[b], topk)
        dt = time.perf_counter() - t0
        job.barrier()
        e2e_s = job.max_over_ranks(dt)
        assert int(h_i.astype(np.int64).sum()) == checksum, "host-path and device-path results differ"
        res["e2e"] = {"value": float(batch) * float(n_total) * steps / e2e_s, "unit": "song-pairs/s",
                      "h2d_bytes_per_step": batch * 4, "d2h_bytes_per_step": batch * topk * 8, "ms_per_step": 1e3 * e2e_s / steps}

    # ---- roofline of the dominant kernel (scan): FP32-bound for batches beyond Q* = 23 queries -------
    sm_max_mhz = float(job.peaks.get("sm_max_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
    fp32_peak = job.sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12          # FFMA lanes x 2 flop x max clock
    pairs_per_launch = float(hi - lo) * batch * steps / max(scan_n, 1)
    scan_avg_ms = scan_ms / max(scan_n, 1)
    achieved = FLOP_PER_PAIR * pairs_per_launch / (scan_avg_ms * 1e-3) / 1e12
    res["roofline"] = {
        "bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
        "kernel": "scan_kernel<%s>" % kernel_shape, "launches_timed": scan_n, "avg_launch_ms": scan_avg_ms,
        "peak_is": f"{job.sm_count} SMs x 128 FFMA lanes x 2 flop x {sm_max_mhz:.0f} MHz (nominal FP32, no tensor cores; "
                   "north_star: min(HBM, FP32) roofline; this batch is FP32-bound, Q* = 23)",
        "algorithmic": f"{FLOP_PER_PAIR} flop/pair x {pairs_per_launch:.3e} pairs per launch",
        "step_frac": FLOP_PER_PAIR * float(hi - lo) * batch / (ms_total / steps * 1e-3) / 1e12 / fp32_peak,
        "other_kernels_ms_per_step": other,
    }
    res["parity_check"] = parity_check(job, feats, n_total, lo, hi, q_host[-1], out_i, out_s, topk)
    return res, feats, q_dev


def hbm_regime(job: Job, q_dev, n_local: int, topk: int) -> dict:
    """The HBM-bound regime of the same kernels (Q < Q* = 23 queries: one pass over the store per call)."""
    torch, eng = job.torch, job.eng
    try:
        hbm_peak = float(job.peaks.get("hbm_gbs") or 6650.0)
        hbm = {"bound": "hbm", "peak": hbm_peak, "unit": "GB/s",
               "peak_is": "MEASURED_PEAKS.json hbm_gbs" if job.peaks.get("hbm_gbs") else "fallback 6650 GB/s",
               "algorithmic": f"{BYTES_PER_SONG} B/song x {n_local} songs per call", "cases": []}
        for nq_small in (1, 4, 16):
            qs = q_dev[0][:nq_small].contiguous()
            os_ = torch.empty((nq_small, topk), dtype=torch.int32, device=job.dev)
            for _ in range(3):
                eng.query_by_index_dev(qs, nq_small, topk, os_, None, job.stream.cuda_stream)
            torch.cuda.synchronize()
            eng.set_option("profile", 1)
            eng.set_option("reset", 1)
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(job.stream)
            for _ in range(20):
                eng.query_by_index_dev(qs, nq_small, topk, os_, None, job.stream.cuda_stream)
            b_.record(job.stream)
            torch.cuda.synchronize()
            t_scan = eng.timing("scan")[0] / 20 * 1e-3
            eng.set_option("profile", 0)
            # the call as a user issues it: no per-kernel event brackets in the stream
            for _ in range(3):
                eng.query_by_index_dev(qs, nq_small, topk, os_, None, job.stream.cuda_stream)
            torch.cuda.synchronize()
            t_call = 1e9
            for _ in range(3):  # best of three rounds of 30 calls (a host-side hiccup in one round must not count)
                a.record(job.stream)
                for _ in range(30):
                    eng.query_by_index_dev(qs, nq_small, topk, os_, None, job.stream.cuda_stream)
                b_.record(job.stream)
                torch.cuda.synchronize()
                t_call = min(t_call, a.elapsed_time(b_) / 30 * 1e-3)
            nbytes = BYTES_PER_SONG * float(n_local)
            hbm["cases"].append({"queries": nq_small, "ms_per_call": t_call * 1e3, "scan_kernel_ms": t_scan * 1e3,
                                 "achieved_call": nbytes / t_call / 1e9, "frac_call": nbytes / t_call / 1e9 / hbm_peak,
                                 "achieved_scan_kernel": nbytes / t_scan / 1e9,
                                 "frac_scan_kernel": nbytes / t_scan / 1e9 / hbm_peak})
        return hbm
    except Exception as exc:  # never lose the headline over the side measurement
        return {"error": str(exc)}


def reference_gpu_path(job: Job) -> dict:
    """BASELINE config 2 (1 M songs, 1024 queries, top-10, host buffers) on this engine and on the reference's own GPU
    path -- the unmodified Recommender.cu rebuilt for sm_100a (oracle/_ref/libref_gpu.so): cuBLAS SGEMV + two kernels
    + D2H + host heap per query (Recommender.cu:184-254, :293-315); its batch mode is sequential calls, timed on a
    bounded sample of the batch."""
    try:
        from oracle_lib import Reference
        from spotify_recommender_b200 import synth
        from spotify_recommender_b200.engine import Engine
        if not Reference.available(gpu=True):
            return {"unavailable": "oracle/_ref/libref_gpu.so not built"}
        n, nq, k, sample = 1_000_000, 1024, 10, 64
        f = synth.features(n)
        q = synth.query_indices(nq, n)
        with Engine(job.local_rank) as e2:
            e2.load_features(f)
            for _ in range(3):
                gi, _ = e2.query_by_index(q, k)
            t0 = time.perf_counter()
            for _ in range(10):
                gi, _ = e2.query_by_index(q, k)
            t_ours = (time.perf_counter() - t0) / 10
        ref = Reference(f, gpu=True)
        if not ref.gpu_enabled():
            return {"unavailable": "the reference fell back to its CPU path"}
        ref.batch(q[:8], k)
        t_ref = 1e9
        for _ in range(2):  # (it cudaMallocs and frees per call: 3 .. 40 ms per query between runs; the better round counts)
            t0 = time.perf_counter()
            ri = ref.batch(q[:sample], k)
            t_ref = min(t_ref, (time.perf_counter() - t0) / sample)
        ref.close()
        same = int((ri == gi[:sample]).all(axis=1).sum())
        return {"config": f"{n} songs x 12, {nq} queries, top-{k}, host buffers in and out",
                "ours_ms_per_batch": t_ours * 1e3, "ours_ms_per_query": t_ours * 1e3 / nq,
                "reference_ms_per_query": t_ref * 1e3, "reference_ms_per_batch_extrapolated": t_ref * 1e3 * nq,
                "reference_sample": f"{sample} sequential recommendByIndex calls of the batch's {nq}",
                "speedup_per_batch": t_ref * nq / t_ours, "identical_ordered_lists": f"{same}/{sample}"}
    except Exception as exc:
        return {"error": str(exc)}


def all_pairs_case(job: Job) -> dict:
    """BASELINE config 5: the top-10 neighbour table of 1 M songs (10^12 scored pairs).  The store is replicated, rank r
    owns the query songs [r * ceil(N/G), ...), no exchange until the final gather of the table.  Wall clock
    per rank around the host-output call plus the gather, max over ranks; a sample of rank 0's rows against the oracle."""
    try:
        from oracle_lib import Oracle
        from spotify_recommender_b200 import synth
        from spotify_recommender_b200.sharded import QueryShardedAllPairs
        n, k = 1_000_000, 10
        f = synth.features(n)
        ap = QueryShardedAllPairs(job.eng, device=job.dev)
        ap.load_replicated(f)
        lo, hi = ap.local_range()
        job.eng.all_pairs_topk(lo, min(hi, lo + 16384), k)  # warm-up
        job.barrier()
        t0 = time.perf_counter()
        li, ls = ap.local_topk(k)
        t_local = time.perf_counter() - t0
        gi, gs = ap.gather_table(li, ls)
        dt = job.max_over_ranks(time.perf_counter() - t0)
        t_local = job.max_over_ranks(t_local)
        sel = np.unique(np.linspace(0, n - 1, PARITY_QUERIES).astype(np.int64)).astype(np.int32)
        wi, ws = Oracle().query_index(f, sel, k, threads=max(1, host_cores() // max(1, job.world)))
        bad = int(((gi[sel] != wi).any(axis=1) | (gs[sel].view(np.uint32) != ws.view(np.uint32)).any(axis=1)).sum())
        ok = bool((gi >= 0).all() and (gi != np.arange(n)[:, None]).all())
        return {"workload": f"BASELINE config 5: all-pairs top-{k} over {n} songs ({n}^2 scored pairs), store replicated, queries sharded "
                            f"over {job.world} GPU(s), host table out", "scaling": "strong", "n_gpus": job.world, "seconds": dt,
                "seconds_local_topk": t_local, "value": float(n) * n / dt, "unit": "song-pairs/s",
                "roofline_frac_fp32": FLOP_PER_PAIR * float(n) * n / max(1, job.world) / t_local / 1e12 / (job.sm_count * 128 * 2 * 1.965e9 / 1e12),
                "parity_check": {"queries": int(sel.size), "mismatches": bad, "top_k": k, "songs": n, "table_complete": ok}}
    except Exception as exc:
        return {"error": str(exc)}


def run_ours(args) -> None:
    global BATCH, TOPK
    if args.batch:
        BATCH = args.batch
    if args.topk:
        TOPK = args.topk
    job = Job(args)
    world, rank = job.world, job.rank
    n_total = int(args.songs_total) if args.songs_total else SONGS_PER_GPU * world

    main, feats, q_dev = measure_case(job, args, n_total, BATCH, TOPK, args.steps, args.warmup, sample_clocks=True)
    n_local = main["songs_per_gpu"]
    roofline = main["roofline"]
    # DRAM traffic of one scan launch: a STORED figure from the committed ncu --set full capture of this command
    # (profiles/scan_kernel_summary.json), not measured in this run -- ncu cannot run inside the timed program
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "scan_kernel_summary.json")) as fh:
            summ = json.load(fh)
        traffic = summ.get("dram_bytes_per_launch")
        traffic_src = f"stored: ncu --set full capture {summ.get('source', 'profiles/scan_kernel_summary.json')} (dram__bytes_read.sum + dram__bytes_write.sum of one launch)"
    except Exception:
        pass
    roofline["traffic"] = traffic
    roofline["traffic_source"] = traffic_src
    roofline["peak_measured_ffma2"] = job.eng.measure_fp32(1)

    single = rank == 0 and world == 1
    top100 = None
    if single and not args.quick and TOPK != 100:
        # BASELINE config 3 proper: the same store and batch, top-100
        job.eng.set_option("reset", 1)
        from spotify_recommender_b200 import synth  # noqa: F401
        top100 = measure_top100(job, args, n_total, feats)
    hbm = hbm_regime(job, q_dev, n_local, TOPK) if single else None
    cpu = None
    if single and not args.no_cpu_baseline:
        cpu = cpu_baseline_port(feats, n_total, TOPK, BATCH)
    del feats, q_dev
    refgpu = reference_gpu_path(job) if single and not args.quick else None

    strong = None
    if not args.quick and not args.songs_total:
        st_steps = max(2, min(args.steps, 5))
        c4, f4, q4 = measure_case(job, args, C4_SONGS_TOTAL, C4_BATCH, C4_TOPK, st_steps, 3)
        del f4, q4
        strong = {"scaling": "strong", "n_gpus": world, "value": c4["value"], "unit": c4["unit"], "ms_per_step": c4["ms_per_step"],
                  "steps": st_steps, "warmup": 3,
                  "config": {"workload": f"BASELINE config 4: {C4_SONGS_TOTAL} songs TOTAL x 12 row-sharded over {world} GPU(s) "
                                         f"({c4['songs_per_gpu']} per GPU), batch of {C4_BATCH} in-store queries, exact top-{C4_TOPK}",
                             "parallelism": f"row-shard x{world}" + (" + one NCCL all-gather of packed keys + merge" if world > 1 else "")},
                  "e2e": c4.get("e2e"), "roofline": {k: c4["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "step_frac", "avg_launch_ms", "other_kernels_ms_per_step")},
                  "parity_check": c4["parity_check"], "gpu_launches": c4["gpu_launches"],
                  "efficiency_is": "T(n_gpus = 1) / (n_gpus x T(n_gpus)) over the ms_per_step of this key at each N"}

    allpairs = all_pairs_case(job) if not args.quick and not args.songs_total else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": "song-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{n_local} songs x 12 features per GPU ({n_total} total, row-sharded), "
                                   f"batch of {BATCH} in-store queries, exact top-{TOPK}",
                       "songs_total": n_total, "songs_per_gpu": n_local, "queries_per_batch": BATCH,
                       "top_k": TOPK, "parallelism": f"row-shard x{world}" + (" + one NCCL all-gather of packed keys + merge" if world > 1 else ""),
                       "l2": "store (2 x 480 MB per GPU) is larger than the 126 MB L2; every step uses a fresh query batch",
                       "kernel_shape": main["kernel_shape"]},
            "queries_per_s_at_10M": main["value"] / 1e7,
            "clocks": main["clocks"],
            "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"],
            "roofline": roofline,
            "parity_check": main["parity_check"],
            "config3_top100": top100,
            "strong_scaling": strong,
            "all_pairs_1M": allpairs,
            "roofline_hbm_regime": hbm,
            "reference_gpu_path": refgpu,
            "cpu_baseline": cpu,
            "result_checksum": main["result_checksum"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        job.dist.barrier()
        job.dist.destroy_process_group()
    job.eng.close()


def measure_top100(job: Job, args, n_total: int, feats) -> dict:
    """The loaded store again at top-100 (BASELINE config 3 proper), device-resident timing + parity."""
    torch, eng = job.torch, job.eng
    steps, warmup, k = max(2, min(args.steps, 10)), 3, 100
    q_host = [(((np.arange(BATCH, dtype=np.int64) + b * BATCH) * 7919 + 13) % n_total).astype(np.int32) for b in range(warmup + steps)]
    q_dev = [torch.from_numpy(q).to(job.dev) for q in q_host]
    oi = torch.empty((BATCH, k), dtype=torch.int32, device=job.dev)
    os_ = torch.empty((BATCH, k), dtype=torch.float32, device=job.dev)
    for b in range(warmup):
        eng.query_by_index_dev(q_dev[b], BATCH, k, oi, os_, job.stream.cuda_stream)
    torch.cuda.synchronize()
    eng.set_option("profile", 1)
    eng.set_option("reset", 1)
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(job.stream)
    for b in range(warmup, warmup + steps):
        eng.query_by_index_dev(q_dev[b], BATCH, k, oi, os_, job.stream.cuda_stream)
    b_.record(job.stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b_) / steps
    scan_ms, scan_n = eng.timing("scan")
    other = {kk: eng.timing(kk)[0] / steps for kk in ("prep", "sample", "bound", "finalize")}
    eng.set_option("profile", 0)
    sm_max_mhz = float(job.peaks.get("sm_max_mhz") or 1965.0)
    fp32_peak = job.sm_count * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    flop_step = FLOP_PER_PAIR * float(n_total) * BATCH
    return {"workload": f"{n_total} songs, batch of {BATCH} queries, exact top-{k} (BASELINE config 3)", "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "value": float(n_total) * BATCH / (ms * 1e-3), "unit": "song-pairs/s",
            "roofline": {"bound": "fp32", "peak": fp32_peak, "unit": "TFLOP/s", "frac": flop_step / (scan_ms / steps * 1e-3) / 1e12 / fp32_peak,
                         "step_frac": flop_step / (ms * 1e-3) / 1e12 / fp32_peak, "avg_launch_ms": scan_ms / max(scan_n, 1),
                         "launches_timed": scan_n, "other_kernels_ms_per_step": other},
            "parity_check": parity_check(job, feats, n_total, 0, n_total, q_host[-1], oi, os_, k)}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--variant", type=int, default=None, help="scan kernel shape (development)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="development: the headline case only (no top-100, config 4, reference GPU path)")
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
