#!/usr/bin/env python
"""Development sweep: ladder-kernel launch shapes vs the opcode interpreter on BASELINE configs 2 and 5.
Checks that every shape reproduces the interpreter's counters bit for bit, then times it with CUDA events.
Needs a library built with `make EXTRA=-DQO_LAD_EXPERIMENT` for variants > 0.

  python tools/lad_sweep.py [--samples 200000] [--variants 0,1,2,3]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=200000)
    ap.add_argument("--variants", default="0,1,2,3,4,5,6")
    ap.add_argument("--workloads", default="cfg2,cfg5")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--s11", action="store_true", help="add an |S11| spec (exercises the second row vector)")
    ap.add_argument("--fp32", action="store_true", help="optional FP32 mode")
    args = ap.parse_args()
    import torch
    import qo100net as Q
    from qo100net import workloads as W
    ctx = Q.Context(device=0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    peak = ctx.measure_dfma_peak()
    print("dfma peak %.2f TFLOP/s" % peak)
    n = args.samples
    for wn in args.workloads.split(","):
        wl = getattr(W, wn)(n, 4096)
        nf = len(wl.f)
        if args.s11:
            fc = 10e6 if wn == "cfg2" else 3e9
            wl.specs = list(wl.specs)[:2] + [(Q.SPEC_S11_MAX_DB, 0.0, 0.8 * fc, -8.0)]
            wl.hist = dict(wl.hist, hist_spec=0)

        def run(kernel, variant):
            os.environ["QO100NET_KERNEL"] = kernel
            os.environ["QO100NET_LAD_VARIANT"] = str(variant)
            plan = Q.Plan(ctx, wl.net, wl.f, wl.specs, seed=wl.seed, tols=wl.tols, precision=32 if args.fp32 else 64, **wl.hist)
            cnt = torch.zeros(plan.num_counters, dtype=torch.int64, device="cuda")
            with torch.cuda.stream(stream):
                plan.launch(0, n, cnt.data_ptr())
                torch.cuda.synchronize()
                ref = cnt.cpu().numpy().copy()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for i in range(args.reps):
                    plan.launch((i + 1) * n, n, cnt.data_ptr())
                b.record(stream)
                torch.cuda.synchronize()
            ms = a.elapsed_time(b) / args.reps
            name, fl = plan.kernel_name, plan.flops_per_eval
            plan.close()
            return ref, ms, name, fl

        base, ms, name, fl = run("interp", 0)
        print("%s interp          %-22s %8.3f ms  %.3e evals/s  frac(ALG-v1 %g) %.3f  pass %d/%d" %
              (wn, name, ms, n * nf / ms * 1e3, fl, fl * n * nf / ms * 1e3 / peak / 1e12, base[0], base[1]))
        for v in [int(x) for x in args.variants.split(",")]:
            try:
                got, ms, name, fl = run("auto", v)
            except Exception as ex:
                print("%s variant %d: %s" % (wn, v, ex))
                continue
            same = bool(np.array_equal(got, base))
            print("%s variant %d       %-22s %8.3f ms  %.3e evals/s  frac(ALG-v1 %g) %.3f  counters==interp %s%s" %
                  (wn, v, name, ms, n * nf / ms * 1e3, fl, fl * n * nf / ms * 1e3 / peak / 1e12, same,
                   "" if same else "  diff at %s" % np.nonzero(got != base)[0][:8]))
    ctx.close()


if __name__ == "__main__":
    main()
