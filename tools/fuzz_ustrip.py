#!/usr/bin/env python
"""Differential fuzz of the microstrip yield kernels (GPU box): random Qucs-style microstrip cascades (lines, corners, tees with
open stubs, a few lumped parts) on random substrates with board-level tolerances, 1-4 frequencies: thread-per-board kernel
(qo_mc_board_kernel) against the item-per-thread kernel (QO100NET_USTRIP=item) and, every --oracle-every-th network, the oracle.

   python tools/fuzz_ustrip.py [--nets 200] [--seed 1] [--out gpurun_out/fuzz_ustrip.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)


def random_board(Q, rng):
    er, h, t = rng.uniform(2.2, 10.0), rng.uniform(0.2e-3, 1.6e-3), rng.uniform(17e-6, 70e-6)
    items = [(Q.SUBST, [er, h, t, rng.uniform(0.001, 0.04), 1.68e-8, 0.15e-6])]
    w0 = h * rng.uniform(0.8, 3.0)
    widths = [w0, w0 * rng.uniform(1.5, 4.0), w0 * rng.uniform(0.4, 0.8)]
    for _ in range(int(rng.integers(3, 16))):
        k = rng.integers(0, 10)
        if k < 5:
            items.append((Q.MLIN, [float(rng.choice(widths)), rng.uniform(0.3e-3, 12e-3)]))
        elif k < 7:
            items.append((Q.MCORN, [widths[0]]))
        elif k < 9:
            w2 = float(rng.choice(widths[1:]))
            items += [(Q.MTEE, [widths[0], widths[0], w2]), (Q.MLIN, [w2, rng.uniform(0.0, 1e-3)]), (Q.MLIN, [w2, rng.uniform(1e-3, 8e-3)]), (Q.MOPEN, [w2])]
        else:
            items.append((Q.SER_L, [rng.uniform(0.2e-9, 2e-9), 0.1, 0.05e-12]) if rng.random() < 0.5 else (Q.SHUNT_C, [rng.uniform(0.1e-12, 1e-12), 0.1, 0.0]))
    tols = [(0, 0, 0, Q.TOL_ABS, 0.02 * er), (0, 1, 1, Q.TOL_REL, 0.08), (0, 2, 3, Q.TOL_REL, 0.2)]
    for i, (k, _p) in enumerate(items):
        if k in (Q.MLIN, Q.MCORN, Q.MOPEN):
            tols.append((i, 0, 2, Q.TOL_ABS, 0.03e-3))
        elif k == Q.MTEE:
            tols += [(i, 0, 2, Q.TOL_ABS, 0.03e-3), (i, 1, 2, Q.TOL_ABS, 0.03e-3), (i, 2, 2, Q.TOL_ABS, 0.03e-3)]
    nf = int(rng.integers(1, 5))
    f = np.sort(rng.uniform(0.5e9, 9e9, nf))
    return Q.Net.from_elements(items, 50.0, 50.0), f, tols


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--nets", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--samples", type=int, default=4000)
    ap.add_argument("--oracle-every", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args(argv)
    import qo100net as Q
    from oracle import refbind as R
    ctx = Q.Context(device=0)
    mism, skipped, compared, oracle_checked, by_nf = [], 0, 0, 0, {}
    for i in range(args.nets):
        rng = np.random.default_rng([args.seed, i])
        net, f, tols = random_board(Q, rng)
        os.environ.pop("QO100NET_USTRIP", None)
        try:
            fs = ctx.mc_run(net, f, [], 5, 200, tols, mode=Q.MODE_FULL_S)["s"]
        except Q.QoError:
            skipped += 1
            continue
        a21, a11 = 20 * np.log10(np.maximum(np.abs(fs[1]), 1e-300)), 20 * np.log10(np.maximum(np.abs(fs[0]), 1e-300))
        if not (np.all(np.isfinite(a21)) and np.all(np.isfinite(a11))):
            skipped += 1
            continue
        specs = []
        for k in range(len(f)):
            kind = int(rng.choice([Q.SPEC_S21_MIN_DB, Q.SPEC_S21_MAX_DB, Q.SPEC_S11_MAX_DB]))
            v = a11[:, k] if kind == Q.SPEC_S11_MAX_DB else a21[:, k]
            if np.quantile(v, 0.9) - np.quantile(v, 0.1) < 0.01:
                continue
            specs.append((kind, float(f[k]) * 0.999, float(f[k]) * 1.001, float(np.quantile(v, rng.uniform(0.3, 0.7)))))
        if not specs:
            skipped += 1
            continue
        hist = dict(hist_bins=32, hist_spec=0, hist_lo=specs[0][3] - 3.0, hist_hi=specs[0][3] + 3.0) if rng.random() < 0.6 else {}
        dist = Q.DIST_GAUSS3S if rng.random() < 0.3 else Q.DIST_UNIFORM
        res = {}
        for label, env in (("board", None), ("item", "item")):
            if env:
                os.environ["QO100NET_USTRIP"] = env
            else:
                os.environ.pop("QO100NET_USTRIP", None)
            plan = Q.Plan(ctx, net, f, specs, seed=40 + i, tols=tols, dist=dist, **hist)
            assert plan.kernel_name == ("qo_mc_board_kernel" if label == "board" else "qo_mc_generic_kernel")
            plan.launch(7 * i, args.samples)
            res[label] = plan.read()
            plan.close()
        os.environ.pop("QO100NET_USTRIP", None)
        compared += 1
        by_nf[len(f)] = by_nf.get(len(f), 0) + 1
        same = res["board"]["n_pass"] == res["item"]["n_pass"] and np.array_equal(res["board"]["fail_per_spec"], res["item"]["fail_per_spec"]) and \
            np.array_equal(res["board"]["hist"], res["item"]["hist"])
        what = "board vs item"
        if same and i % args.oracle_every == 0:
            o = R.mc_run(R.make_elems(net.elements), 50.0, 50.0, f, specs, R.mc_cfg(40 + i, args.samples, tols, sample_offset=7 * i, dist=dist, **hist), nthreads=R.max_threads())
            oracle_checked += 1
            same = o["n_pass"] == res["board"]["n_pass"] and np.array_equal(o["fail_per_spec"], res["board"]["fail_per_spec"]) and np.array_equal(o["hist"], res["board"]["hist"])
            what = "board vs oracle"
        if not same:
            mism.append({"net": i, "what": what, "nf": len(f), "f": [float(x) for x in f], "elements": [(int(k), [float(x) for x in p]) for k, p in net.elements], "specs": specs,
                         "board": [int(res["board"]["n_pass"])] + [int(v) for v in res["board"]["fail_per_spec"]],
                         "item": [int(res["item"]["n_pass"])] + [int(v) for v in res["item"]["fail_per_spec"]]})
    out = {"networks": args.nets, "skipped": skipped, "compared": compared, "oracle_checked": oracle_checked, "by_frequencies": by_nf,
           "mismatches": len(mism), "details": mism[:10], "seed": args.seed, "samples": args.samples}
    print(json.dumps(out, indent=1))
    if args.out:
        open(args.out, "w").write(json.dumps(out, indent=1))
    ctx.close()
    return out


if __name__ == "__main__":
    sys.exit(1 if main()["mismatches"] else 0)
