#!/usr/bin/env python
"""Differential fuzz of the nodal kernels on random N-port networks (GPU box): the run-time compiled kernel (QO100NET_NODAL=jit)
against the interpreted static plan (=static) and per-point partial pivoting (=dense), on FULL_S planes of a few samples and on
Monte-Carlo counters with specs placed at quantiles of those planes; every --oracle-every-th network also against the oracle.

   python tools/fuzz_nodal.py [--nets 60] [--seed 1] [--out gpurun_out/fuzz_nodal.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)

NB_R, NB_L, NB_C, NB_VCVS, NB_SBLOCK = 1, 2, 3, 4, 5


def random_network(rng):
    nn = int(rng.integers(3, 13))
    fc = 10.0 ** rng.uniform(7.0, 9.3)
    wc = 2 * np.pi * fc
    z0 = 50.0
    br = []

    def branch(a, b):
        k = rng.choice([NB_R, NB_L, NB_C, NB_C, NB_R])
        g = rng.uniform(0.2, 5.0)
        if k == NB_R:
            return (NB_R, [a, b], [z0 * g])
        if k == NB_L:
            L = g * z0 / wc
            par = rng.random() < 0.5
            return (NB_L, [a, b], [L, rng.uniform(0.05, 2.0) if par else 0.0, (1 / (L * (wc * rng.uniform(5, 40)) ** 2)) if par else 0.0])
        C = g / (z0 * wc)
        par = rng.random() < 0.5
        return (NB_C, [a, b], [C, rng.uniform(0.02, 1.0) if par else 0.0, (1 / (C * (wc * rng.uniform(8, 50)) ** 2)) if par else 0.0])
    for node in range(2, nn + 1):                    # spanning tree over nodes 1..nn
        br.append(branch(int(rng.integers(1, node)), node))
    for _ in range(int(rng.integers(1, nn + 2))):    # extra branches, many of them to ground
        a = int(rng.integers(1, nn + 1))
        b = 0 if rng.random() < 0.6 else int(rng.integers(1, nn + 1))
        if a != b:
            br.append(branch(a, b))
    br.append((NB_R, [1, 0], [z0 * rng.uniform(2, 20)]))          # a DC path to ground
    nports = int(rng.integers(1, min(4, nn) + 1))
    pnodes = [int(v) for v in rng.choice(np.arange(1, nn + 1), nports, replace=False)]
    ports = [(p, float(rng.choice([50.0, 50.0, 75.0]))) for p in pnodes]
    if rng.random() < 0.25 and nn >= 4:              # an ideal buffer: in+ at a node, output drives a fresh internal node through a resistor
        a = int(rng.integers(1, nn + 1))
        nn += 1
        br.append((NB_VCVS, [a, nn, 0, 0], [rng.uniform(0.5, 3.0), 0.0 if rng.random() < 0.5 else rng.uniform(0.0, 0.3) / fc]))
        br.append((NB_R, [nn, int(rng.integers(1, nn))], [z0 * rng.uniform(0.5, 3.0)]))
    if rng.random() < 0.3 and fc < 2.5e9:            # the measured inductor of pa-bias-simulation.sch:39 between two nodes (or to ground)
        a = int(rng.integers(1, nn + 1))
        b = int(rng.integers(0, nn + 1))
        if a != b:
            br.append((NB_SBLOCK, [a, b, 0], [0, float(rng.integers(0, 2)), 50.0]))
    tols = []
    for i, (k, _n, _p) in enumerate(br):
        if k in (NB_R, NB_L, NB_C) and rng.random() < 0.8:
            tols.append((i, 0, len(tols), 0, float(rng.choice([0.01, 0.02, 0.05, 0.1]))))
    if not tols:
        tols.append((0, 0, 0, 0, 0.05))
    nf = int(rng.choice([5, 33, 100, 257, 1000]))
    span = rng.uniform(2.0, 8.0)
    return br, nn, ports, tols, fc, nf, span


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--nets", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--samples", type=int, default=2000)
    ap.add_argument("--oracle-every", type=int, default=4)
    ap.add_argument("--out", default=None)
    args = ap.parse_args(argv)
    import qo100net as Q
    from oracle import refbind as R
    ctx = Q.Context(device=0)
    kernels, mism, skipped, oracle_checked, compared = {}, [], 0, 0, 0
    for i in range(args.nets):
        rng = np.random.default_rng([args.seed, i])
        br, nn, ports, tols, fc, nf, span = random_network(rng)
        f = Q.grid_log(fc / span, fc * span, nf) if rng.random() < 0.5 else Q.grid_lin(fc / span, fc * span, nf)
        nd = Q.Nodal(nn)
        if any(k == NB_SBLOCK for k, _n, _p in br):
            g = np.load(os.path.join(ROOT, "tests", "golden", "touchstone.npz"))
            fd, sd = g["11SQ39N_f"], g["11SQ39N_s"]
            idx = nd.add_sblock(Q.SBlock.from_arrays(fd, sd[:, 0], sd[:, 1], sd[:, 2], sd[:, 3], 50.0))
            R.sblock_clear()
            R.sblock_register(0, fd, sd[:, 0], sd[:, 1], sd[:, 2], sd[:, 3], 50.0)
            br = [(k, n_, [idx, p[1], p[2]] if k == NB_SBLOCK else p) for k, n_, p in br]
        for kind, nodes, p in br:
            nd.add_branch(kind, nodes, p)
        for node, z in ports:
            nd.add_port(node, z)
        npn = len(ports)
        planes = {}
        try:
            for mode in ("jit", "static", "dense"):
                os.environ["QO100NET_NODAL"] = mode
                planes[mode] = ctx.nodal_mc_run(nd, f, [], 77 + i, 24, tols, mode=Q.MODE_FULL_S)["s"]
                kernels[ctx.nodal_last_kernel()] = kernels.get(ctx.nodal_last_kernel(), 0) + 1
        except Q.QoError:
            skipped += 1                              # the static plan was refused (or a singular matrix): nothing to compare
            os.environ.pop("QO100NET_NODAL", None)
            nd.close()
            continue
        ref = planes["dense"]
        if not np.all(np.isfinite(ref.view(float))):
            skipped += 1
            nd.close()
            continue
        compared += 1
        bad = None
        tol_abs = 1e-9 * max(1.0, float(np.max(np.abs(ref))))
        for mode in ("jit", "static"):
            if not np.all(np.abs(planes[mode] - ref) <= 1e-8 * np.abs(ref) + tol_abs):
                bad = "FULL_S %s vs dense: max abs diff %.3e" % (mode, float(np.max(np.abs(planes[mode] - ref))))
        if not np.array_equal(planes["jit"], planes["static"]) and bad is None:
            d = float(np.max(np.abs(planes["jit"] - planes["static"])))
            if d > 1e-12 * max(1.0, float(np.max(np.abs(ref)))):
                bad = "FULL_S jit vs static: max abs diff %.3e" % d
        # Monte-Carlo counters: specs on random S entries at quantiles of the sampled planes
        specs = []
        a = 20 * np.log10(np.maximum(np.abs(ref), 1e-300))                 # [samples, nf, np, np]
        for _ in range(int(rng.integers(1, 5))):
            r_, c_ = int(rng.integers(0, npn)), int(rng.integers(0, npn))
            lo, hi = sorted(rng.uniform(f[0], f[-1], 2))
            band = (f >= lo) & (f <= hi)
            if not band.any():
                continue
            kind = int(rng.choice([Q.SPEC_S21_MIN_DB, Q.SPEC_S21_MAX_DB]))
            v = a[:, band, r_, c_].min(axis=1) if kind == Q.SPEC_S21_MIN_DB else a[:, band, r_, c_].max(axis=1)
            if np.quantile(v, 0.9) - np.quantile(v, 0.1) < 0.01 or v.min() < -200.0:
                continue
            specs.append((kind, r_, c_, float(lo), float(hi), float(np.quantile(v, rng.uniform(0.3, 0.7)))))
        if specs and bad is None:
            hist = dict(hist_bins=32, hist_spec=0, hist_lo=specs[0][5] - 3.0, hist_hi=specs[0][5] + 3.0) if rng.random() < 0.5 else {}
            res = {}
            for mode in ("jit", "static", "dense"):
                os.environ["QO100NET_NODAL"] = mode
                res[mode] = ctx.nodal_mc_run(nd, f, specs, 500 + i, args.samples, tols, sample_offset=12345 + i, **hist)
            for mode in ("jit", "static"):
                if not (res[mode]["n_pass"] == res["dense"]["n_pass"] and np.array_equal(res[mode]["fail_per_spec"], res["dense"]["fail_per_spec"])
                        and np.array_equal(res[mode]["hist"], res["dense"]["hist"])):
                    bad = "counters %s vs dense: %d vs %d" % (mode, res[mode]["n_pass"], res["dense"]["n_pass"])
            if bad is None and i % args.oracle_every == 0:
                from oracle import refbind
                o = R.nodal_mc_run([(k, n_ + [0] * (4 - len(n_)), p + [0.0] * (4 - len(p))) for k, n_, p in br], nn, ports, f, specs,
                                   refbind.mc_cfg(500 + i, args.samples, tols, sample_offset=12345 + i, **hist), nthreads=R.max_threads())
                oracle_checked += 1
                if not (o["n_pass"] == res["jit"]["n_pass"] and np.array_equal(o["fail_per_spec"], res["jit"]["fail_per_spec"]) and np.array_equal(o["hist"], res["jit"]["hist"])):
                    bad = "counters jit vs oracle: %d vs %d" % (res["jit"]["n_pass"], o["n_pass"])
        os.environ.pop("QO100NET_NODAL", None)
        if bad:
            mism.append({"net": i, "what": bad, "n_nodes": nn, "ports": ports, "nf": nf, "branches": [(int(k), [int(x) for x in n_], [float(x) for x in p]) for k, n_, p in br],
                         "specs": specs})
        nd.close()
    out = {"networks": args.nets, "skipped": skipped, "compared": compared, "oracle_checked": oracle_checked, "kernels": kernels,
           "mismatches": len(mism), "details": mism[:10], "seed": args.seed}
    print(json.dumps(out, indent=1))
    if args.out:
        open(args.out, "w").write(json.dumps(out, indent=1))
    ctx.close()
    return out


if __name__ == "__main__":
    sys.exit(1 if main()["mismatches"] else 0)
