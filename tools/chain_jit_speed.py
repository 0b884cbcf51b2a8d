#!/usr/bin/env python
"""Throughput of the run-time compiled chain kernel (qo_chain_jit.h) against the opcode interpreter on jobs the polynomial
kernels cannot take: a line inside the config-2 ladder with a measured two-port behind it, the config-2 ladder itself kept off
the polynomial kernels (QO100NET_KERNEL=interp), and config 5 written out in full (FULL_S behind the coupler).  Counters of both
kernels are compared.

  python tools/chain_jit_speed.py [--samples 200000] [--fs-samples 16384] [--out gpurun_out/chain_jit.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)


def measure(Q, torch, ctx, stream, samples=200000, fs_samples=16384, modes=("jit", "interp")):
    from qo100net import workloads as W

    class args:
        pass
    args.samples, args.fs_samples = samples, fs_samples
    w = W.cfg2()
    g = np.load(os.path.join(ROOT, "tests", "golden", "touchstone.npz"))
    fd, sd = g["11SQ39N_f"], g["11SQ39N_s"]
    blk = Q.SBlock.from_arrays(fd, sd[:, 0], sd[:, 1], sd[:, 2], sd[:, 3], 50.0)
    el = w.net.elements
    head = Q.Net.from_elements([(k, list(p)) for k, p in el[:5]], 50.0, 50.0)
    rest = Q.Net.from_elements([(k, list(p)) for k, p in el[5:]], 50.0, 50.0)
    mixed = head.concat(Q.Net.from_elements([(Q.TLINE, [75.0, 20.0, 10e6])], 50.0, 50.0)).concat(rest).concat(blk.as_net(True, 50.0, 50.0))
    mtol = [(e if e < 5 else e + 1, p, v, m, t) for (e, p, v, m, t) in w.tols] + [(5, 0, 40, Q.TOL_REL, 0.05), (5, 1, 41, Q.TOL_REL, 0.03)]
    db = 20 * np.log10(np.abs(ctx.sweep(mixed, w.f)[1]))
    mspec = [(Q.SPEC_S21_MIN_DB, 0.0, 9e6, float(db[w.f <= 9e6].min()) - 0.3), (Q.SPEC_S21_MAX_DB, 2e7, 1e99, float(db[w.f >= 2e7].max()) + 1.0)]
    mhist = dict(hist_bins=64, hist_spec=0, hist_lo=float(db[w.f <= 9e6].min()) - 2.0, hist_hi=float(db[w.f <= 9e6].min()) + 0.5)
    out = {}

    def reduce_case(name, net, f, specs, tols, hist, kernel_env):
        row = {}
        for label in modes:
            os.environ["QO100NET_CHAIN"] = label
            if kernel_env:
                os.environ["QO100NET_KERNEL"] = kernel_env
            plan = Q.Plan(ctx, net, f, specs, seed=3, tols=tols, **hist)
            t0 = time.time()
            with torch.cuda.stream(stream):
                plan.launch(10 ** 9, 4096)
                torch.cuda.synchronize()
                first = time.time() - t0
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                best = 1e30
                for rep in range(3):
                    a.record(stream)
                    plan.launch(rep * args.samples, args.samples)
                    b.record(stream)
                    torch.cuda.synchronize()
                    best = min(best, a.elapsed_time(b))
            r = plan.read()
            row[label] = {"kernel": plan.kernel_name, "evals_per_s": args.samples * len(f) / (best * 1e-3), "ms": best,
                          "first_launch_s": first, "n_pass": int(r["n_pass"]), "n_total": int(r["n_total"]),
                          "fail_per_spec": [int(x) for x in r["fail_per_spec"]], "hist_sum": int(r["hist"].sum())}
            plan.close()
        os.environ.pop("QO100NET_CHAIN", None)
        os.environ.pop("QO100NET_KERNEL", None)
        if len(modes) == 2:
            row["counters_equal"] = all(row["jit"][k] == row["interp"][k] for k in ("n_pass", "n_total", "fail_per_spec", "hist_sum"))
            row["speedup"] = row["jit"]["evals_per_s"] / row["interp"]["evals_per_s"]
        out[name] = row

    reduce_case("cfg2 ladder with a line inside and a measured two-port behind", mixed, w.f, mspec, mtol, mhist, None)
    reduce_case("cfg2 ladder kept off the polynomial kernels", w.net, w.f, w.specs, w.tols, w.hist, "interp")
    w5 = W.cfg5(1000)
    reduce_case("cfg5 (coupler + ladder) kept off the polynomial kernels", w5.net, w5.f, w5.specs, w5.tols, w5.hist, "interp")
    # FULL_S behind the coupler
    n = args.fs_samples
    buf = torch.empty((4, n, len(w5.f), 2), dtype=torch.float64, device="cuda")
    row = {}
    keep = {}
    for label in modes:
        os.environ["QO100NET_CHAIN"] = label
        plan = Q.Plan(ctx, w5.net, w5.f, [], seed=1, tols=w5.tols, mode=Q.MODE_FULL_S)
        with torch.cuda.stream(stream):
            plan.launch(0, n, None, buf.data_ptr())
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(3):
                plan.launch(0, n, None, buf.data_ptr())
            b.record(stream)
            torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        keep[label] = buf[:, :64].clone()
        row[label] = {"kernel": plan.kernel_name, "ms": ms, "gb_per_s": n * len(w5.f) * 64 / ms * 1e-6, "evals_per_s": n * len(w5.f) / (ms * 1e-3)}
        plan.close()
    os.environ.pop("QO100NET_CHAIN", None)
    if len(modes) == 2:
        row["max_abs_diff"] = float((keep["jit"] - keep["interp"]).abs().max())
        row["speedup"] = row["interp"]["ms"] / row["jit"]["ms"]
    out["cfg5 FULL_S (%d samples x %d points)" % (n, len(w5.f))] = row
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=200000)
    ap.add_argument("--fs-samples", type=int, default=16384)
    ap.add_argument("--out", default=None)
    ap.add_argument("--jit-only", action="store_true")
    a = ap.parse_args()
    import torch
    import qo100net as Q
    ctx = Q.Context(device=0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    out = measure(Q, torch, ctx, stream, a.samples, a.fs_samples, ("jit",) if a.jit_only else ("jit", "interp"))
    print(json.dumps(out, indent=1))
    if a.out:
        open(a.out, "w").write(json.dumps(out, indent=1))
    ctx.close()


if __name__ == "__main__":
    main()
