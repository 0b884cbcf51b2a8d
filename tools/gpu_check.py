"""Developer smoke/parity/timing script for a GPU box (not a test, not the bench):
   python tools/gpu_check.py [quick]
Compares the CUDA path with the CPU oracle on the BASELINE configs and prints timings."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)
import qo100net as Q            # noqa: E402
from oracle import refbind as R  # noqa: E402


def relerr(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def to_ref(net):
    return R.make_elems(net.elements)


def pa_lpf_net():
    zw, mm = 0.75e-3, 1e-3
    ML = lambda L, W=zw: (Q.MLIN, [W, L * mm])
    CO = (Q.MCORN, [zw])
    items = [(Q.SUBST, [4.5, 0.6e-3, 34.79e-6, 0.045, 1.68e-8, 0.15e-6]),
             ML(1.65), CO, ML(1.40415), CO, ML(1.65), CO, ML(1.40415), CO, ML(0.95),
             (Q.MTEE, [zw, zw, zw]), ML(0.15), ML(4.35, 3e-3), (Q.MOPEN, [3e-3]),
             ML(0.95), CO, ML(2.5), CO, ML(3.45), CO, ML(1.55), CO, ML(0.95),
             (Q.MTEE, [zw, zw, zw]), ML(2.5), ML(3.0, 3e-3), (Q.MOPEN, [3e-3]),
             ML(0.95), CO, ML(0.8), CO, ML(1.6), CO, ML(1.75), CO, ML(5.95)]
    return Q.Net.from_elements(items, 50, 50)


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    ctx = Q.Context(device=0)
    print("version", Q.lib().qo_version().decode())
    print("dfma peak TFLOP/s", ctx.measure_dfma_peak())

    x = np.concatenate([np.random.default_rng(0).uniform(0.5, 2.0, 1 << 20), 10.0 ** np.random.default_rng(1).uniform(-30, 30, 1 << 20)])
    r = ctx.device_rcp(x)
    print("rcp max rel err (ulp of 2^-53)", float(np.max(np.abs(r * x - 1.0))) / 2.0 ** -53)

    # cfg 1: IF BPF
    bpf = Q.Net.from_elements([(Q.SER_LC_SER, [33e-9, 4.7e-12]), (Q.SHUNT_LC_PAR, [4.7e-9, 33e-12]), (Q.SER_LC_SER, [33e-9, 4.7e-12])])
    f1 = Q.grid_log(300e6 / 3.5, 500e6 * 3.5, 1024)
    g = ctx.sweep(bpf, f1, gd=True)
    o = R.sweep(to_ref(bpf), 50, 50, f1, gd=True)
    print("cfg1 S21 rel", relerr(g[1], o[1]), "S11 abs", float(np.max(np.abs(g[0] - o[0]))), "gd rel", relerr(g[4], o[4]))
    g32 = ctx.sweep(bpf, f1, precision=32)
    print("cfg1 fp32 dB err", float(np.max(np.abs(20 * np.log10(np.abs(g32[1])) - 20 * np.log10(np.abs(o[1]))))))

    # cfg 2
    fc = 10e6
    lad = Q.Net.cheby_lpf(11, 0.1, fc, 50.0).add_parasitics(fc)
    f2 = Q.grid_log(fc / 2.5, fc * 6.25, 4096)
    g = ctx.sweep(lad, f2)
    o = R.sweep(to_ref(lad), 50, 50, f2)
    print("cfg2 nominal S21 rel", relerr(g[1], o[1]), "S12", relerr(g[2], o[2]), "S11 abs", float(np.max(np.abs(g[0] - o[0]))), "S22 abs", float(np.max(np.abs(g[3] - o[3]))))
    specs = [(Q.SPEC_S21_MIN_DB, 0, 0.95 * fc, -2.0), (Q.SPEC_S21_MAX_DB, 1.3 * fc, 1e99, -49.0)]
    tols = Q.lc_tolerances(lad, 0.05, 0.02)
    n = 2000
    t = time.time()
    og = R.mc_run(to_ref(lad), 50, 50, f2, specs, R.mc_cfg(0x5EED010000000002, n, tols, hist_bins=256, hist_spec=0, hist_lo=-4, hist_hi=0), nthreads=R.max_threads())
    tcpu = time.time() - t
    gg = ctx.mc_run(lad, f2, specs, 0x5EED010000000002, n, tols, hist_bins=256, hist_spec=0, hist_lo=-4.0, hist_hi=0.0)
    print("cfg2 MC n=%d oracle pass %d fails %s | gpu pass %d fails %s | hist equal %s | cpu %.2fs (%d thr, %.3g evals/s)" % (
        n, og["n_pass"], og["fail_per_spec"], gg["n_pass"], gg["fail_per_spec"], np.array_equal(og["hist"], gg["hist"]), tcpu, R.max_threads(), n * 4096 / tcpu))
    fs_g = ctx.mc_run(lad, f2, [], 7, 64, tols, mode=Q.MODE_FULL_S)["s"]
    fs_o = R.mc_run(to_ref(lad), 50, 50, f2, [], R.mc_cfg(7, 64, tols), full_s=True)["s"]
    print("cfg2 FULL_S S21 rel", relerr(fs_g[1], fs_o[1]), "S11 abs", float(np.max(np.abs(fs_g[0] - fs_o[0]))))
    fac = ctx.device_perturb_factors(1234567, 10, 5000, 11, Q.DIST_UNIFORM, 0.05)
    ref = np.array([[R.lib().ref_perturb_factor(1234567, 10 + s, v, 0, 0.05) for v in range(11)] for s in range(5000)])
    print("perturbation factors bit-exact (uniform):", np.array_equal(fac, ref))
    fac = ctx.device_perturb_factors(1234567, 10, 5000, 11, Q.DIST_GAUSS3S, 0.05)
    ref = np.array([[R.lib().ref_perturb_factor(1234567, 10 + s, v, 1, 0.05) for v in range(11)] for s in range(5000)])
    print("perturbation factors bit-exact (gauss3s):", np.array_equal(fac, ref))

    # timing cfg 2
    for ns in ([100000] if quick else [100000, 1000000]):
        plan = Q.Plan(ctx, lad, f2, specs, seed=0x5EED010000000002, tols=tols, hist_bins=256, hist_spec=0, hist_lo=-4.0, hist_hi=0.0)
        plan.launch(0, 1000)
        plan.read()
        plan.reset()
        t = time.time()
        plan.launch(0, ns)
        res = plan.read()
        dt = time.time() - t
        print("cfg2 GPU n=%d: %.4f s  %.4g evals/s  yield %.4f  flops/eval %.0f -> %.2f TFLOP/s (ALG-v1)" % (
            ns, dt, ns * 4096 / dt, res["n_pass"] / res["n_total"], plan.flops_per_eval, ns * 4096 / dt * plan.flops_per_eval * 1e-12))
        plan.close()

    # cfg 5 network
    cp = Q.Net.from_elements([(Q.CPL_THRU, [55.2771, 45.2267, 95.4225, 95.4225, 2.4e9, 50.0])])
    lad3 = Q.Net.cheby_lpf(11, 0.1, 3e9, 50.0).add_parasitics(3e9)
    net5 = cp.concat(lad3)
    f5 = Q.grid_lin(70e6, 4000e6, 4096)
    g = ctx.sweep(net5, f5)
    o = R.sweep(to_ref(net5), 50, 50, f5)
    print("cfg5 nominal S21 rel", relerr(g[1], o[1]), "S11 abs", float(np.max(np.abs(g[0] - o[0]))))
    tols5 = [(0, 0, 0, Q.TOL_REL, 0.02), (0, 1, 1, Q.TOL_REL, 0.02), (0, 2, 2, Q.TOL_REL, 0.01), (0, 3, 2, Q.TOL_REL, 0.01)]
    tols5 += [(e, p, v + 3, m, tl) for (e, p, v, m, tl) in Q.lc_tolerances(net5, 0.05, 0.02)]
    specs5 = [(Q.SPEC_S21_MIN_DB, 2.3e9, 2.5e9, -1.4), (Q.SPEC_S21_MAX_DB, 3.9e9, 1e99, -48.0)]
    n = 1000
    og = R.mc_run(to_ref(net5), 50, 50, f5, specs5, R.mc_cfg(5, n, tols5, hist_bins=256, hist_spec=0, hist_lo=-3, hist_hi=0), nthreads=R.max_threads())
    gg = ctx.mc_run(net5, f5, specs5, 5, n, tols5, hist_bins=256, hist_spec=0, hist_lo=-3.0, hist_hi=0.0)
    print("cfg5 MC oracle pass %d fails %s | gpu pass %d fails %s | hist equal %s" % (og["n_pass"], og["fail_per_spec"], gg["n_pass"], gg["fail_per_spec"], np.array_equal(og["hist"], gg["hist"])))
    plan = Q.Plan(ctx, net5, f5, specs5, seed=5, tols=tols5, hist_bins=256, hist_spec=0, hist_lo=-3.0, hist_hi=0.0)
    plan.launch(0, 1000); plan.read(); plan.reset()
    ns = 100000
    t = time.time(); plan.launch(0, ns); res = plan.read(); dt = time.time() - t
    print("cfg5 GPU n=%d: %.4f s %.4g evals/s yield %.4f flops/eval %.0f" % (ns, dt, ns * 4096 / dt, res["n_pass"] / res["n_total"], plan.flops_per_eval))
    plan.close()

    # cfg 3 / oracle-pinned microstrip network
    pa = pa_lpf_net()
    f3 = Q.grid_lin(1e7, 1e10, 5000)
    t = time.time(); g = ctx.sweep(pa, f3); dt = time.time() - t
    o = R.sweep(to_ref(pa), 50, 50, f3)
    print("PA LPF sweep %.3fs  S21 rel" % dt, relerr(g[1], o[1]), "S12", relerr(g[2], o[2]), "S11 abs", float(np.max(np.abs(g[0] - o[0]))), "S22 abs", float(np.max(np.abs(g[3] - o[3]))))
    wel = [i for i, (k, p) in enumerate(pa.elements) if k in (Q.MLIN, Q.MCORN, Q.MOPEN)]
    tols3 = [(0, 0, 0, Q.TOL_ABS, 0.2), (0, 1, 1, Q.TOL_REL, 0.10), (0, 2, 3, Q.TOL_REL, 0.20)]
    tols3 += [(i, 0, 2, Q.TOL_ABS, 0.05e-3) for i in wel]
    for i, (k, p) in enumerate(pa.elements):
        if k == Q.MTEE:
            tols3 += [(i, 0, 2, Q.TOL_ABS, 0.05e-3), (i, 1, 2, Q.TOL_ABS, 0.05e-3), (i, 2, 2, Q.TOL_ABS, 0.05e-3)]
    fh = np.array([2.4e9, 4.8e9, 7.2e9])
    specs3 = [(Q.SPEC_S21_MIN_DB, 2.3e9, 2.5e9, -1.0), (Q.SPEC_S21_MAX_DB, 4.7e9, 4.9e9, -22.0), (Q.SPEC_S21_MAX_DB, 7.1e9, 7.3e9, -8.5)]
    n = 2000
    og = R.mc_run(to_ref(pa), 50, 50, fh, specs3, R.mc_cfg(3, n, tols3, hist_bins=64, hist_spec=1, hist_lo=-30, hist_hi=-15), nthreads=R.max_threads())
    gg = ctx.mc_run(pa, fh, specs3, 3, n, tols3, hist_bins=64, hist_spec=1, hist_lo=-30.0, hist_hi=-15.0)
    print("cfg3 MC oracle pass %d fails %s | gpu pass %d fails %s | hist equal %s" % (og["n_pass"], og["fail_per_spec"], gg["n_pass"], gg["fail_per_spec"], np.array_equal(og["hist"], gg["hist"])))
    ns = 200000
    t = time.time(); gg = ctx.mc_run(pa, fh, specs3, 3, ns, tols3); dt = time.time() - t
    print("cfg3 GPU n=%d: %.3f s (kernel %.3f s) %.4g evals/s yield %.4f" % (ns, dt, gg["seconds"], ns * 3 / gg["seconds"], gg["n_pass"] / gg["n_total"]))

    # cfg 4 FULL_S bandwidth
    net4 = Q.Net.from_elements([(Q.SHUNT_C, [430e-12]), (Q.SER_L, [1.3e-6]), (Q.SHUNT_C, [620e-12]), (Q.SER_L, [1.3e-6]),
                                (Q.SHUNT_C, [560e-12]), (Q.SER_L, [1.1e-6]), (Q.SHUNT_C, [240e-12])], 100.0, 50.0)
    f4 = Q.grid_log(4e6, 62.5e6, 4096)
    import torch
    ns = 16384
    buf = torch.empty((4, ns, 4096, 2), dtype=torch.float64, device="cuda")
    plan = Q.Plan(ctx, net4, f4, [], seed=4, tols=Q.lc_tolerances(net4, 0.05, 0.05), mode=Q.MODE_FULL_S)
    plan.launch(0, ns, None, buf.data_ptr()); plan.read()
    t = time.time(); plan.launch(0, ns, None, buf.data_ptr()); plan.read(); dt = time.time() - t
    print("cfg4 FULL_S n=%d: %.4f s %.4g evals/s %.1f GB/s" % (ns, dt, ns * 4096 / dt, ns * 4096 * 64 / dt * 1e-9))
    fs_o = R.mc_run(to_ref(net4), 100, 50, f4, [], R.mc_cfg(4, 8, Q.lc_tolerances(net4, 0.05, 0.05)), full_s=True)["s"]
    got = torch.view_as_complex(buf[:, :8].contiguous()).cpu().numpy()
    print("cfg4 FULL_S vs oracle S21 rel", relerr(got[1], fs_o[1]), "S22 abs", float(np.max(np.abs(got[3] - fs_o[3]))))


if __name__ == "__main__":
    main()
