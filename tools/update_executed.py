#!/usr/bin/env python
"""Rebuild profiles/executed_fp64.json from the ncu summaries of one profiling round (tools/profile_all.sh <tag> all):

   python tools/update_executed.py r02x

reads profiles/<tag>_{ts_cfg2,ts_cfg5,tf_cfg2,tf_cfg5,ladder_cfg2,ladder_cfg5}_raw_metrics.txt (whatever exists), takes the
EXECUTED FP64 instruction counts per eval, the DRAM bytes per launch and the FP64-pipe utilisation from them, and stamps the
file with the git commit and the hash of the kernel sources the captures were taken at (bench.py checks the hash)."""
import json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash          # noqa: E402

tag = sys.argv[1]
P = os.path.join(ROOT, "profiles")
path = os.path.join(P, "executed_fp64.json")
out = json.load(open(path)) if os.path.exists(path) else {}
KERNELS = {"ts": "qo_mc_ts_kernel", "tf": "qo_mc_tf_kernel", "ladder": "qo_mc_ladder_kernel"}
for short, kname in KERNELS.items():
    for cfg in ("cfg2", "cfg5"):
        fn = os.path.join(P, "%s_%s_%s_raw_metrics.txt" % (tag, short, cfg))
        if not os.path.exists(fn):
            continue
        txt = open(fn).read()
        m = re.search(r"# per eval: dfma ([0-9.]+)\s+dmul ([0-9.]+)\s+dadd ([0-9.]+)", txt)
        rd = re.search(r"dram__bytes_read.sum\s+(\S+)\s+([0-9.,]+)", txt)
        wr = re.search(r"dram__bytes_write.sum\s+(\S+)\s+([0-9.,]+)", txt)
        pipe = re.search(r"sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active\s+\S+\s+([0-9.]+)", txt)
        ev = re.search(r"evals per launch = ([0-9.e+]+)", txt)
        unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = sum(float(x.group(2).replace(",", "")) * unit.get(x.group(1), 1) for x in (rd, wr) if x)
        out.setdefault(kname, {})[cfg] = {
            "dfma": float(m.group(1)), "dmul": float(m.group(2)), "dadd": float(m.group(3)), "dram_bytes_per_launch": int(dram),
            "fp64_pipe_active_pct": float(pipe.group(1)) if pipe else None,
            "source": "profiles/%s (ncu --set full --clock-control none, %s evals per launch)" % (os.path.basename(fn), ev.group(1) if ev else "?")}
        print(kname, cfg, out[kname][cfg])
out["_doc"] = ("EXECUTED FP64 instructions per (sample, frequency) eval and DRAM bytes per launch of the Monte-Carlo kernels, from ncu "
               "(smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on.sum / evals, dram__bytes_{read,write}.sum); bench.py multiplies "
               "them by the rate it measures.  _src_hash = sha256 over qo-100-tools_b200/csrc/*.{cu,cuh,h} at capture time.")
out["_src_hash"] = kernel_source_hash()
out["_git"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
out["_tag"] = tag
try:                                          # machine-code identity of the profiled build (tools/sass_hash.py)
    from tools.sass_hash import write_lib_hashes
    k = write_lib_hashes()["kernels"]
    out["_sass"] = {n: k[n]["sha"] for n in k}
    out["_sass_git"] = out["_git"]
except Exception as ex:
    print("no SASS hashes:", ex)
json.dump(out, open(path, "w"), indent=1)
print("stamped", out["_git"], out["_src_hash"])
