#!/usr/bin/env python
"""Summarise an ncu report of one kernel: headline counters, stall reasons, and the share of samples /
executed instructions of each backward-branch loop (innermost first).
   python tools/ncu_regions.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys, re, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[2]
def get(n):
    return v[h.index(n)] if n in h else "n/a"
for n in ["gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
          "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active"]:
    print("%-70s %s" % (n, get(n)))
st = []
for i, n in enumerate(h):
    if "pcsamp_warps_issue_stalled" in n and "not_issued" not in n:
        try: st.append((float(v[i].replace(",", "")), n.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError: pass
st.sort(reverse=True); tot = sum(x for x, _ in st)
print("stalls:", ", ".join("%s %.1f%%" % (n, 100 * x / tot) for x, n in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ia, isrc, iss, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(data[0][ia], 16)
recs = [(int(r[ia], 16) - base, r[isrc].strip(), int(r[iss]), int(r[iex])) for r in data]
ts, te = sum(r[2] for r in recs), sum(r[3] for r in recs)
loops = []
for off, t, _, _ in recs:
    m = re.search(r"\bBRA\b.*0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16) - base
        if 0 <= tgt < off: loops.append((tgt, off))
for lo, hi in sorted(loops, key=lambda x: x[1] - x[0]):
    s = sum(r[2] for r in recs if lo <= r[0] <= hi); e = sum(r[3] for r in recs if lo <= r[0] <= hi)
    if s / ts < 0.01: continue
    it = max(r[3] for r in recs if lo <= r[0] <= hi)
    c = collections.Counter()
    for off, t, _, ex in recs:
        if lo <= off <= hi:
            p = t.split(); c[(p[1] if p[0].startswith("@") else p[0]).split(".")[0]] += ex
    print("loop %05x-%05x: samples %5.1f%%  inst %5.1f%%  per trip: %s" % (lo, hi, 100 * s / ts, 100 * e / te,
          ", ".join("%s %.1f" % (k, n / it) for k, n in c.most_common(12))))
