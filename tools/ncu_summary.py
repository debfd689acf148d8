"""Summarise one `ncu --set full` kernel capture (.ncu-rep) into profiles/scan_kernel_summary.json.

    python tools/ncu_summary.py gpurun_out/final/scan_full.ncu-rep profiles/scan_kernel_summary.json

Reads the report with `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: v for h, v in zip(hdr, vals)}
u = {h: x for h, x in zip(hdr, units)}


def num(key, default=None):
    v = m.get(key)
    if v in (None, ""):
        return default
    return float(v.replace(",", ""))


def to_bytes(key):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return num(key) * scale[u[key]]


def to_ms(key):
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    return num(key) * scale[u[key]]


issue = num("smsp__issue_active.avg.per_cycle_active") or num("smsp__inst_executed.avg.per_cycle_active")
stalls = {}
pre = "smsp__average_warps_issue_stalled_"
suf = "_per_issue_active.ratio"
for h in hdr:
    if h.startswith(pre) and h.endswith(suf):
        stalls[h[len(pre):-len(suf)]] = round(num(h), 6)
rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
summary = {
    "kernel": m.get("Kernel Name"),
    "command": "python bench.py --steps 2 --warmup 3 --no-cpu-baseline (ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 13 -c 1)",
    "workload": "10M Spotify-schema synthetic songs x 1024-query group (one of 4 scan launches per 4096-query batch), top-10",
    "gpu_time_ms": to_ms("gpu__time_duration.sum"),
    "dram_bytes_read": rd,
    "dram_bytes_write": wr,
    "dram_bytes_per_launch": rd + wr,
    "algorithmic_flop_per_launch": 24.0 * 1e7 * 1024,
    "algorithmic_bytes_single_pass": 48.0 * 1e7,
    "sm__pipe_fma_cycles_active_pct": num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    "sm__inst_executed_pipe_fma_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "dram_throughput_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "issue_active_pct": None if issue is None else issue * 100.0,
    "registers_per_thread": int(num("launch__registers_per_thread")),
    "grid": int(num("launch__grid_size")),
    "block": int(num("launch__block_size")),
    "l2_hit_rate_pct": num("lts__t_sector_hit_rate.pct"),
    "tensor_pipe_active_pct": num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0),
    "sm_cycles_active_avg": num("sm__cycles_active.avg"),
    "sm_cycles_elapsed_avg": num("sm__cycles_elapsed.avg"),
    "stalls_per_issue": dict(sorted(stalls.items())),
}
json.dump(summary, open(out, "w"), indent=1)
print(json.dumps(summary, indent=1))
