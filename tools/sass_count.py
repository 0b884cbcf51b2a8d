#!/usr/bin/env python
"""Static SASS census of the hot loops of libqo100net.so (cuobjdump -sass): for every instantiation of the
ladder kernel, the innermost frequency loop's instruction mix per (sample, frequency) point.  Written to
profiles/sass_counts.json; bench.py reads the FP64-pipe instruction count from there to report the FP64
pipe utilisation next to the ALG-v1 roofline fraction.

  python tools/sass_count.py            # needs cuobjdump (CUDA toolkit), no GPU
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qo-100-tools_b200", "lib", "libqo100net.so")
FP64 = ("DFMA", "DMUL", "DADD", "DSETP")


def functions(sass):
    for blk in re.split(r"\n\s*Function : ", sass)[1:]:
        name = blk.split("\n", 1)[0].strip()
        ins = []
        for ln in blk.split("\n"):
            m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        yield name, ins


def loops(ins):
    out = []
    for a, t in ins:
        m = re.search(r"\bBRA\b.*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            out.append((int(m.group(1), 16), a))
    return out


def census(ins, lo, hi):
    c = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi:
            p = t.split()
            op = p[1] if p[0].startswith("@") else p[0]
            c[op.split(".")[0]] += 1
    return c


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    res = {}
    for name, ins in functions(sass):
        if "qo_mc_ladder_kernel" not in name:
            continue
        dn = demangle(name)
        m = re.search(r"<(double|float), (\d+), (\d+), \(?(?:bool\))?(\w+), (\d+), (\d+), (\d+), (\d+)>", dn)
        if not m:
            continue
        fp32 = m.group(1) == "float"
        n, first, cpl, nrows, pp = int(m.group(2)), int(m.group(3)), m.group(4) in ("1", "true"), int(m.group(5)), int(m.group(6))
        # the frequency loop = the loop with the most FP64 instructions that contains no other loop with FP64 work
        best = None
        ls = loops(ins)
        for lo, hi in ls:
            c = census(ins, lo, hi)
            fp = sum(c[k] for k in FP64)
            inner = [(a, b) for a, b in ls if (a, b) != (lo, hi) and lo <= a and b <= hi and sum(census(ins, a, b)[k] for k in FP64) > 20]
            if not inner and (best is None or fp > best[0]):
                best = (fp, c, lo, hi)
        if not best:
            continue
        fp, c, lo, hi = best
        pts = 2 * pp
        # DSETP and the trackers' selects are counted statically for all four spec slots of both the uniform
        # and the edge path; per point the uniform path executes one DSETP per ACTIVE spec
        fp_core = (c["DFMA"] + c["DMUL"] + c["DADD"]) / pts
        if fp32:
            continue        # the census counts FP64-pipe instructions; the FP32 mode runs on the FMA pipe
        res["n%d_first%d_cpl%d%s" % (n, first, int(cpl), "_s11" if nrows == 2 else "")] = {
            "kernel": dn, "points_per_thread_iteration": pts, "loop_sass_instructions": sum(c.values()),
            "dfma_per_eval": c["DFMA"] / pts, "dmul_per_eval": c["DMUL"] / pts, "dadd_per_eval_static": c["DADD"] / pts,
            "mufu_per_eval": c["MUFU"] / pts, "lds_per_eval": c["LDS"] / pts, "ldg_per_eval": c["LDG"] / pts,
            "fp64_pipe_instr_per_eval": fp_core, "mix_static": dict(c.most_common()),
        }
    out = os.path.join(ROOT, "profiles", "sass_counts.json")
    json.dump(res, open(out, "w"), indent=1, sort_keys=True)
    for k in ("n11_first0_cpl0", "n11_first0_cpl1"):
        if k in res:
            r = res[k]
            print(k, "FP64-pipe instr/eval %.1f (DFMA %.1f DMUL %.1f) MUFU %.2f LDS %.2f loop %d SASS" %
                  (r["fp64_pipe_instr_per_eval"], r["dfma_per_eval"], r["dmul_per_eval"], r["mufu_per_eval"], r["lds_per_eval"],
                   r["loop_sass_instructions"]))
    print("wrote", out, len(res), "kernels")


if __name__ == "__main__":
    sys.exit(main())
