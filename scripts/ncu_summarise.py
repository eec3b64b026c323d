"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/:
  launches csv (gpu__time_duration.sum per launch)  -> per-kernel totals and the dominant kernel's share
  .ncu-rep of one k_sweep launch (--set full)        -> a JSON of the metrics DESIGN.md quotes + sweep_traffic.json
Usage: python scripts/ncu_summarise.py <tag>   (reads gpurun_out/launches_bench.csv, gpurun_out/prof_sweep_u398.ncu-rep)"""
import csv, io, json, subprocess, sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

src = ROOT / "gpurun_out" / "launches_bench.csv"
lines = [l for l in src.read_text().splitlines() if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void ", "")
    tot[name][0] += 1
    tot[name][1] += float(r["Metric Value"]) / 1e6
all_ms = sum(v[1] for v in tot.values())
out = ROOT / "profiles" / f"{tag}_launches_bench_py.csv"
out.write_text("\n".join(lines) + "\n")
summary = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv python bench.py --steps 2 --warmup 3 "
                      "--no-cpu-baseline --no-api --no-tc  (the 3 184-unit strong-scaling batch)", "total_device_ms": all_ms,
           "kernels": {k: {"launches": v[0], "ms": v[1], "share": v[1] / all_ms} for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])}}
(ROOT / "profiles" / f"{tag}_launches_bench_py_summary.json").write_text(json.dumps(summary, indent=1))
print(json.dumps(summary, indent=1)[:1200])

rep = ROOT / "gpurun_out" / "prof_sweep_u398.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    head, units, vals = rr[0], rr[1], rr[2]
    keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
            "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    d = {}
    for k in keep:
        if k in head:
            i = head.index(k)
            d[k] = f"{vals[i]} {units[i]}".strip()
    (ROOT / "profiles" / f"{tag}_ksweep_u398_ncu_full_summary.json").write_text(json.dumps(d, indent=1))
    def num(k):
        i = head.index(k)
        v = float(vals[i].replace(",", ""))
        u = units[i].lower()
        return v * (1e6 if u.startswith("mbyte") else 1e3 if u.startswith("kbyte") else 1e9 if u.startswith("gbyte") else 1.0)
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    kname = d.get("Kernel Name", "k_sweep").split("(")[0].replace("void ", "")
    tr = {"kernel": kname, "workload": "config 2 (one pullback pair): 398 units x 36000 candidates, N=M=520",
          "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch_config2": rd + wr,
          "dram_bytes_per_unit": (rd + wr) / 398.0,
          "source": f"ncu --set full --clock-control none, profiles/{tag}_ksweep_u398_ncu_full_summary.json",
          "algorithmic_bytes_per_launch": 398 * (2 * 520 * 16 + 36000 * 4),
          "algorithmic_bytes_note": "per unit: both point sets once (f64 x, y) + one FP32 distance per candidate"}
    (ROOT / "profiles" / "sweep_traffic.json").write_text(json.dumps(tr, indent=1))
    print(json.dumps(d, indent=1))
