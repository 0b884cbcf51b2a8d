#!/usr/bin/env python
"""Throughput of cascades with a non-rational block in front of a lumped ladder: transfer-function kernel (row vector of the
block per point x polynomials of the ladder) against the opcode interpreter (QO100NET_TF_NO_FRONT=1).

  python tools/front_block_speed.py [--samples 200000] [--out gpurun_out/front.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=200000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import qo100net as Q
    from qo100net import workloads as W
    ctx = Q.Context(device=0)
    w = W.cfg2()
    g = np.load(os.path.join(ROOT, "tests", "golden", "touchstone.npz"))
    fd, sd = g["11SQ39N_f"], g["11SQ39N_s"]
    blk = Q.SBlock.from_arrays(fd, sd[:, 0], sd[:, 1], sd[:, 2], sd[:, 3], 50.0)
    cases = {"line + cfg2 ladder": (Q.Net.from_elements([(Q.TLINE, [75.0, 35.0, 10e6])], 50.0, 50.0).concat(w.net),
                                    [(0, 0, 40, Q.TOL_REL, 0.05), (0, 1, 41, Q.TOL_REL, 0.03)] + [(e + 1, p, v, m, t) for (e, p, v, m, t) in w.tols]),
             "measured two-port + cfg2 ladder": (blk.as_net(True, 50.0, 50.0).concat(w.net), [(e + 1, p, v, m, t) for (e, p, v, m, t) in w.tols])}
    out = {}
    for name, (net, tols) in cases.items():
        row = {}
        for label, env in (("tf", None), ("interpreter", "1")):
            if env:
                os.environ["QO100NET_TF_NO_FRONT"] = env
            else:
                os.environ.pop("QO100NET_TF_NO_FRONT", None)
            plan = Q.Plan(ctx, net, w.f, w.specs, seed=3, tols=tols, **w.hist)
            kname = plan.kernel_name
            plan.close()
            ctx.mc_run(net, w.f, w.specs, 3, 4096, tols, **w.hist)
            best = None
            for rep in range(3):
                r = ctx.mc_run(net, w.f, w.specs, 3, args.samples, tols, sample_offset=rep * args.samples, **w.hist)
                best = r if best is None or r["seconds"] < best["seconds"] else best
            row[label] = {"kernel": kname, "evals_per_s": args.samples * len(w.f) / best["seconds"], "n_pass": int(best["n_pass"])}
        os.environ.pop("QO100NET_TF_NO_FRONT", None)
        out[name] = row
    print(json.dumps(out, indent=1))
    if args.out:
        open(args.out, "w").write(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
