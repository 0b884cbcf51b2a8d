#!/usr/bin/env python
"""All five BASELINE.json configurations at their named shapes on this box, next to the CPU oracle
(single thread and all host cores, bounded samples).  One JSON document on stdout / --out.

  python tools/run_configs.py [--out gpurun_out/configs.json] [--quick]
  python -m torch.distributed.run --nproc-per-node N ... tools/run_configs.py --only cfg5     # 1e8 samples sharded over N GPUs

Timing: CUDA events on the launching stream around plan launches (tables resident), after warm-up;
N > 1: max over ranks.  The CPU figures are reported baselines (oracle port), not targets.
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


def cpu_rate(R, wl, n, nthreads, full_s=False):
    e = R.make_elems(wl.net.elements)
    rs, rl = wl.net.terminations
    cfg = R.mc_cfg(wl.seed, n, wl.tols, **(wl.hist if wl.specs else {}))
    t0 = time.perf_counter()
    R.mc_run(e, rs, rl, wl.f, wl.specs, cfg, nthreads=nthreads, full_s=full_s)
    return n * len(wl.f) / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="cfg1,cfg2,cfg3,cfg4,cfg5")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    import torch
    import qo100net as Q
    from qo100net import dist as qd
    from qo100net import workloads as W
    from oracle import refbind as R
    rank, world, local = qd.env_rank_world()
    torch.cuda.set_device(local)
    dist = qd.init_process_group("nccl") if world > 1 else None
    ctx = Q.Context(device=local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    peak = ctx.measure_dfma_peak()
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    nthr = R.max_threads()
    only = args.only.split(",")
    out = {"n_gpus": world, "dfma_peak_tflops": peak, "hbm_peak_gbs": hbm, "host_cores": nthr, "configs": {}}

    def timed_plan(wl, n_total, reps, mode=None, buf_samples=0):
        """n_total samples of the GLOBAL job per rep, sharded contiguously over the ranks."""
        kw = dict(mode=Q.MODE_FULL_S) if mode == "full_s" else dict(wl.hist)
        plan = Q.Plan(ctx, wl.net, wl.f, wl.specs, seed=wl.seed, tols=wl.tols, **kw)
        lo, hi = qd.shard_range(n_total, rank, world)
        cnt = torch.zeros(plan.num_counters, dtype=torch.int64, device="cuda")
        buf = torch.empty((4, buf_samples, len(wl.f), 2), dtype=torch.float64, device="cuda") if mode == "full_s" else None
        with torch.cuda.stream(stream):
            def go(rep):
                if mode == "full_s":
                    plan.launch(lo, hi - lo, None, buf.data_ptr())
                else:
                    plan.launch(rep * n_total + lo, hi - lo, cnt.data_ptr())
                    qd.allreduce_counters(cnt)
            go(0)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            cnt.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for r in range(reps):
                go(1 + r)
            b.record(stream)
            torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        c = cnt.cpu().numpy()
        name, fl = plan.kernel_name, plan.flops_per_eval
        plan.close()
        return float(ms.item()), name, fl, c

    q = 10 if args.quick else 1
    if "cfg1" in only and rank == 0:
        wl = W.cfg1()
        ctx.sweep(wl.net, wl.f)
        t0 = time.perf_counter()
        for _ in range(20):
            g = ctx.sweep(wl.net, wl.f)
        gpu_ms = (time.perf_counter() - t0) / 20 * 1e3
        t0 = time.perf_counter()
        for _ in range(20):
            o = R.sweep(R.make_elems(wl.net.elements), 50, 50, wl.f)
        cpu_ms = (time.perf_counter() - t0) / 20 * 1e3
        out["configs"]["cfg1"] = {"workload": wl.name, "nf": len(wl.f), "gpu_sweep_ms_e2e": gpu_ms, "cpu_oracle_sweep_ms": cpu_ms,
                                  "s21_rel_err_vs_oracle": float(np.max(np.abs(g[1] - o[1]) / np.abs(o[1]))),
                                  "note": "1024-point nominal sweep, host buffers, launch+copy latency dominated"}
    if "cfg2" in only:
        n = 1000000 // q
        wl = W.cfg2(n, 4096)
        ms, name, fl, c = timed_plan(wl, n * world, 5)
        ev = n * world * 4096 / ms * 1e3
        d = {"workload": wl.name, "samples": n * world, "nf": 4096, "ms": ms, "evals_per_s": ev, "kernel": name, "alg_v1_flops_per_eval": fl,
             "roofline_frac_alg_v1": fl * ev / world / (peak * 1e12), "yield": float(c[0]) / max(1, int(c[1]))}
        if rank == 0 and not args.no_cpu:
            d["cpu_1_thread_evals_per_s"] = cpu_rate(R, wl, 1500, 1)
            d["cpu_all_cores_evals_per_s"] = cpu_rate(R, wl, 20000, nthr)
        out["configs"]["cfg2"] = d
    if "cfg3" in only and rank == 0:
        n = 10000000 // q
        wl = W.cfg3(n)
        ms, name, fl, c = timed_plan(wl, n, 2)
        d = {"workload": wl.name, "samples": n, "nf": 3, "ms": ms, "evals_per_s": n * 3 / ms * 1e3, "samples_per_s": n / ms * 1e3, "kernel": name,
             "yield": float(c[0]) / max(1, int(c[1])),
             "note": "transcendental/latency bound (Qucs microstrip models); excluded from the FMA roofline claim (SURVEY 8d)"}
        if not args.no_cpu:
            d["cpu_1_thread_evals_per_s"] = cpu_rate(R, wl, 2000, 1)
            d["cpu_all_cores_evals_per_s"] = cpu_rate(R, wl, 40000, nthr)
        out["configs"]["cfg3"] = d
        # the lumped twin BASELINE.json's wording describes: 3 points per sample, per-sample work dominates
        wl = W.cfg3b(n)
        ms, name, fl, c = timed_plan(wl, n, 2)
        d = {"workload": wl.name, "samples": n, "nf": 3, "ms": ms, "evals_per_s": n * 3 / ms * 1e3, "samples_per_s": n / ms * 1e3, "kernel": name,
             "yield": float(c[0]) / max(1, int(c[1]))}
        if not args.no_cpu:
            d["cpu_all_cores_evals_per_s"] = cpu_rate(R, wl, 400000, nthr)
        out["configs"]["cfg3b"] = d
    if "cfg4" in only and rank == 0:
        n = 65536 // q
        rows = {}
        for wl in W.cfg4(n, 4096):
            ms, name, fl, _ = timed_plan(wl, n, 3, mode="full_s", buf_samples=n)
            gbs = n * 4096 * 64 / ms * 1e-6
            rows[wl.name] = {"samples": n, "ms": ms, "evals_per_s": n * 4096 / ms * 1e3, "write_gbs": gbs, "hbm_frac": gbs / hbm, "kernel": name,
                             "bytes_written": n * 4096 * 64}
            if not args.no_cpu:
                rows[wl.name]["cpu_all_cores_evals_per_s"] = cpu_rate(R, wl, 256, nthr, full_s=True)
        out["configs"]["cfg4"] = rows
    if "cfg5" in only:
        n_total = (100000000 if world >= 2 else 12500000) // q
        wl = W.cfg5(n_total, 4096)
        ms, name, fl, c = timed_plan(wl, n_total, 1 if n_total >= 50000000 else 2)
        ev = n_total * 4096 / ms * 1e3
        d = {"workload": wl.name, "samples": n_total, "nf": 4096, "ms": ms, "seconds_for_job": ms * 1e-3, "evals_per_s": ev, "kernel": name,
             "alg_v1_flops_per_eval": fl, "roofline_frac_alg_v1": fl * ev / world / (peak * 1e12), "yield": float(c[0]) / max(1, int(c[1])),
             "counters_head": [int(x) for x in c[:5]]}
        if rank == 0 and not args.no_cpu:
            d["cpu_1_thread_evals_per_s"] = cpu_rate(R, wl, 1000, 1)
            d["cpu_all_cores_evals_per_s"] = cpu_rate(R, wl, 16000, nthr)
        out["configs"]["cfg5"] = d
    if rank == 0:
        txt = json.dumps(out, indent=1)
        print(txt)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            open(args.out, "w").write(txt)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
