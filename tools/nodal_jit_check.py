#!/usr/bin/env python
"""A/B of the nodal kernels on the reference's 5-port bias network (tests/test_nodal.py::hand_netlist): interpreted static plan
against the run-time compiled kernel at several register budgets; counters must be identical, FULL_S planes equal to 1e-12.

  python tools/nodal_jit_check.py [--samples 100000] [--nf 1000] [--out gpurun_out/nodal_jit.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=100000)
    ap.add_argument("--nf", type=int, default=1000)
    ap.add_argument("--minb", default="4,5,6,8")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import qo100net as Q
    import test_nodal as T
    g = np.load(os.path.join(ROOT, "tests", "golden", "touchstone.npz"))
    nd, br, nn, ports = T.build_nodal(Q, g)
    ctx = Q.Context(device=0)
    f = Q.grid_lin(1e8, 3e9, args.nf)
    specs = [(Q.SPEC_S21_MIN_DB, 1, 0, 2.3e9, 2.5e9, -3.0), (Q.SPEC_S21_MAX_DB, 2, 0, 2.3e9, 2.5e9, -25.0)]
    tols = [(i, 0, v, Q.TOL_REL, 0.05 if b[0] == T.NB_C else 0.01) for v, (i, b) in
            enumerate((i, b) for i, b in enumerate(br) if b[0] in (T.NB_R, T.NB_C))]
    hist = dict(hist_bins=64, hist_spec=0, hist_lo=-6.0, hist_hi=0.0)
    n = args.samples

    def run(mode, minb=None):
        os.environ["QO100NET_NODAL"] = mode
        if minb:
            os.environ["QO100NET_NODAL_JIT_MINB"] = str(minb)
        ctx.nodal_mc_run(nd, f, specs, 5, 512, tols, **hist)
        best = None
        for rep in range(3):
            r = ctx.nodal_mc_run(nd, f, specs, 5, n, tols, sample_offset=7, **hist)
            best = r if best is None or r["seconds"] < best["seconds"] else best
        fs = ctx.nodal_mc_run(nd, f[:64], [], 5, 16, tols, mode=Q.MODE_FULL_S)["s"]
        return best, ctx.nodal_last_kernel(), fs

    ref, kname, fs_ref = run("static")
    out = {"workload": "pa-bias 5-port network, 23 unknowns, %d samples x %d points, reduce-only, 2 specs" % (n, args.nf),
           "static": {"kernel": kname, "points_per_s": n * args.nf / ref["seconds"], "n_pass": int(ref["n_pass"])}, "jit": []}
    for mb in [int(x) for x in args.minb.split(",")]:
        r, kn, fs = run("jit", mb)
        out["jit"].append({"kernel": kn, "minb": mb, "points_per_s": n * args.nf / r["seconds"], "n_pass": int(r["n_pass"]),
                           "counters_equal_static": bool(r["n_pass"] == ref["n_pass"] and np.array_equal(r["hist"], ref["hist"])
                                                         and np.array_equal(r["fail_per_spec"], ref["fail_per_spec"])),
                           "full_s_max_abs_diff": float(np.max(np.abs(fs - fs_ref)))})
    os.environ.pop("QO100NET_NODAL", None)
    print(json.dumps(out, indent=1))
    if args.out:
        open(args.out, "w").write(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
