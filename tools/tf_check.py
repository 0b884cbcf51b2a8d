#!/usr/bin/env python
"""Development check of the transfer-function kernel (qo_tf.cuh): integer counters against the opcode
interpreter and the straight-line ladder kernel on BASELINE configs 2 / 5 / 5p and on rf-tools filters with
tanks and traps, then CUDA-event timing of each kernel.

  python tools/tf_check.py [--samples 200000] [--reps 5]
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
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--workloads", default="cfg2,cfg5,cfg5p,ifbpf,gpsdo15,ideal11")
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

    def workload(name):
        if name in ("cfg2", "cfg5", "cfg5p"):
            return getattr(W, name)(n, 4096)
        if name in ("cfg2gd", "cfg2s11"):
            w = W.cfg2(n, 4096)
            fc = 10e6
            if name == "cfg2gd":
                gd = ctx.sweep(w.net, w.f, gd=True)[4]
                band = (w.f >= 0.3 * fc) & (w.f <= 0.9 * fc)
                w.specs = list(w.specs) + [(Q.SPEC_GD_MAX, 0.3 * fc, 0.9 * fc, float(gd[band].max()) * 1.01)]
            else:
                w.specs = list(w.specs) + [(Q.SPEC_S11_MAX_DB, 0.0, 0.8 * fc, -8.0)]
            return w
        if name == "ifbpf":
            net = W.if_bpf_net()
            f = Q.grid_log(300e6 / 3.5, 500e6 * 3.5, 4096)
            db = 20 * np.log10(np.abs(ctx.sweep(net, f)[1]))
            pb = (f >= 3.6e8) & (f <= 4.4e8)
            specs = [(Q.SPEC_S21_MIN_DB, 3.6e8, 4.4e8, float(db[pb].min()) - 0.3), (Q.SPEC_S21_MAX_DB, 9e8, 1e99, float(db[f >= 9e8].max()) + 1.0)]
            return W.Workload("ifbpf", net, f, specs, Q.lc_tolerances(net, 0.05, 0.05),
                              dict(hist_bins=64, hist_spec=0, hist_lo=-6.0, hist_hi=0.0), n, 11)
        if name == "gpsdo15":
            _, net, fc = W.gpsdo_bank()[1]
            f = Q.grid_log(fc / 2.5, fc * 6.25, 4096)
            db = 20 * np.log10(np.abs(ctx.sweep(net, f)[1]))
            pb, sb = f <= 0.8 * fc, f >= 2.0 * fc
            specs = [(Q.SPEC_S21_MIN_DB, 0.0, 0.8 * fc, float(db[pb].min()) - 0.2), (Q.SPEC_S21_MAX_DB, 2.0 * fc, 1e99, float(db[sb].max()) + 3.0)]
            return W.Workload("gpsdo15", net, f, specs, Q.lc_tolerances(net, 0.05, 0.05),
                              dict(hist_bins=64, hist_spec=1, hist_lo=-80.0, hist_hi=-20.0), n, 12)
        if name == "ideal11":
            fc = 10e6
            net = Q.Net.cheby_lpf(11, 0.1, fc, 50.0, True)
            f = Q.grid_log(fc / 2.5, fc * 6.25, 4096)
            specs = [(Q.SPEC_S21_MIN_DB, 0.0, 0.95 * fc, -0.5), (Q.SPEC_S21_MAX_DB, 1.3 * fc, 1e99, -49.0)]
            return W.Workload("ideal11", net, f, specs, Q.lc_tolerances(net, 0.05, 0.02),
                              dict(hist_bins=64, hist_spec=0, hist_lo=-2.0, hist_hi=0.0), n, 13)
        raise SystemExit("unknown workload " + name)

    for wn in args.workloads.split(","):
        wl = workload(wn)
        nf = len(wl.f)

        def run(kernel):
            os.environ["QO100NET_KERNEL"] = kernel
            plan = Q.Plan(ctx, wl.net, wl.f, wl.specs, seed=wl.seed, tols=wl.tols, **wl.hist)
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
            if kernel == "auto":
                print("         tf plan:", plan.tf_info)
            plan.close()
            return ref, ms, name, fl

        base = None
        for kernel in ("interp", "ladder", "auto"):
            got, ms, name, fl = run(kernel)
            if base is None:
                base = got
            same = bool(np.array_equal(got, base))
            print("%-8s %-7s %-22s %8.3f ms  %.3e evals/s  frac(ALG-v1 %g) %.3f  pass %d/%d  counters==interp %s%s" %
                  (wn, kernel, name, ms, n * nf / ms * 1e3, fl, fl * n * nf / ms * 1e3 / peak / 1e12, got[0], got[1], same,
                   "" if same else "  diff %s" % [(int(i), int(got[i]) - int(base[i])) for i in np.nonzero(got != base)[0][:8]]))
    os.environ.pop("QO100NET_KERNEL", None)
    ctx.close()


if __name__ == "__main__":
    main()
