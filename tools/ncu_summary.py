#!/usr/bin/env python
"""Writes the judged summaries of one ncu report under profiles/:
   python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/TAG_name  <evals per launch>
-> profiles/TAG_name_details.txt (ncu --page details), profiles/TAG_name_raw_metrics.txt (selected raw counters + per-eval
   FP64 instruction counts + region breakdown from the source page)."""
import csv, io, subprocess, sys
rep, out, evals = sys.argv[1], sys.argv[2], float(sys.argv[3])
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
open(out + "_details.txt", "w").write(det)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active", "sm__cycles_elapsed.max",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]
lines = ["# selected raw counters of %s (ncu --set full --clock-control none); evals per launch = %g" % (rep.split("/")[-1], evals)]
val = {}
for i, n in enumerate(h):
    val[n] = v[i]
    if n in keep or "pcsamp_warps_issue_stalled" in n and "not_issued" not in n:
        lines.append("%-82s %-14s %s" % (n, u[i], v[i]))
cyc = float(val["sm__cycles_elapsed.max"].replace(",", ""))
per = {}
for op in ("dfma", "dmul", "dadd"):
    per[op] = float(val["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op].replace(",", "")) * cyc / evals
tot = float(val["smsp__inst_executed.sum"].replace(",", "")) * float(val.get("smsp__thread_inst_executed_per_inst_executed.ratio", "32").replace(",", "")) / evals
lines.append("# per eval: dfma %.2f  dmul %.2f  dadd %.2f  (FP64-pipe arithmetic %.2f)  all thread instructions %.1f" %
             (per["dfma"], per["dmul"], per["dadd"], sum(per.values()), tot))
dram = float(val["dram__bytes_read.sum"].replace(",", "")) + float(val["dram__bytes_write.sum"].replace(",", ""))
lines.append("# dram bytes per launch: read %s %s + write %s %s" % (val["dram__bytes_read.sum"], u[h.index("dram__bytes_read.sum")], val["dram__bytes_write.sum"], u[h.index("dram__bytes_write.sum")]))
reg = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_regions.py"), rep], capture_output=True, text=True).stdout
lines.append("# region breakdown (tools/ncu_regions.py): share of stall samples / executed instructions per backward-branch loop")
lines += ["# " + l[:230] for l in reg.splitlines() if l.startswith(("loop", "stalls"))]
open(out + "_raw_metrics.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-12:]))
