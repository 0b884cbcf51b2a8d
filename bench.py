#!/usr/bin/env python
"""bench.py -- Monte-Carlo network evals/s (1 eval = one (sample, frequency) point).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg5]

A "step" is one pass of the hot path over one batch of synthetic input: BASELINE config 2
(pcb/generic-filter 11th-order 0.1 dB Chebyshev LPF with ESR/SRF parasitics, +-5 % L / +-2 % C,
1e6 tolerance samples x 4096 log-spaced frequency points, reduce-only yield + 256-bin histogram)
per GPU.  N > 1 (torchrun, one process per GPU) shards the samples: every rank takes 1e6 samples of
a global N x 1e6 job (weak scaling), and the only collective is the sum all-reduce of the uint64
counters.  Prints ONE JSON line (rank 0).

  value     whole-job evals/s with tables/grid/specs already resident in HBM (qo_plan_launch)
  e2e       the same through the host-buffer C-ABI call qo_mc_run (H2D of grid/specs/tolerances and
            D2H of the counters inside the timed region)
  roofline  achieved ALG-v1 TFLOP/s of the dominant kernel vs the FP64-FMA peak MEASURED in this run
            by a dependency-free DFMA loop (MEASURED_PEAKS.json holds no FP64 figure)
  cpu_baseline / --impl reference
            the CPU oracle (oracle/, the port of the models the reference's external tools apply)
            on the box's host cores, on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)

REAL_STDOUT = sys.stdout
METRIC = "Monte Carlo network evals/s (samples x freq pts)"
UNIT = "evals/s"
SAMPLES_PER_GPU = 1000000
NF = 4096
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel and its EXECUTED FP64 instruction
# counts per eval, both from the committed ncu --set full captures (profiles/executed_fp64.json names the reports)


def executed_profile(plan_kernel, wl_name):
    """{"dfma", "dmul", "dadd" per eval, "dram_bytes_per_launch", "source"} of the kernel on this workload, or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "executed_fp64.json")))
        return d[plan_kernel]["cfg5" if wl_name.startswith("cfg5") else "cfg2"]
    except Exception:
        return None


def time_plan(plan, stream, counters, nspg, steps, warmup, torch):
    """mean kernel ms over `steps` launches of a resident plan (CUDA events on the launching stream)"""
    with torch.cuda.stream(stream):
        for i in range(warmup):
            plan.launch(i * nspg, nspg, counters.data_ptr())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for i in range(steps):
            plan.launch((warmup + i) * nspg, nspg, counters.data_ptr())
        b.record(stream)
        torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def workload(name, n_samples):
    from qo100net import workloads as W
    if name == "cfg5":
        return W.cfg5(n_samples, NF)
    return W.cfg2(n_samples, NF)


def cpu_reference(wl, n_samples, steps, warmup, nthreads):
    """Times the oracle port (OpenMP over samples) on a bounded sample; returns (evals/s, s/step)."""
    from oracle import refbind as R
    e = R.make_elems(wl.net.elements)
    rs, rl = wl.net.terminations
    times = []
    for i in range(warmup + steps):
        cfg = R.mc_cfg(wl.seed, n_samples, wl.tols, sample_offset=i * n_samples, **wl.hist)
        t0 = time.perf_counter()
        R.mc_run(e, rs, rl, wl.f, wl.specs, cfg, nthreads=nthreads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return n_samples * len(wl.f) * len(times) / tot, tot / len(times)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        # the sampler also sees idle gaps; "under load" = samples in the upper half of the power range
        load = [s for s, p in zip(sm, pw) if pw and p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = workload(args.workload, SAMPLES_PER_GPU)
    from oracle import refbind as R
    nthr = R.max_threads()
    n = 20000                                         # bounded sample per step (the workload is 1e6 per GPU)
    v, per_step = cpu_reference(wl, n, args.steps, args.warmup, nthr)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "nf": NF, "samples_per_step": n,
                       "note": "the reference tree holds no evaluator for this path (external GUI tools); this is the "
                               "CPU oracle port of the same models, OpenMP over samples on all host cores"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": nthr, "kind": "port",
                             "sample": "%d samples x %d freq per step, %d steps" % (n, NF, args.steps)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def emit(line):
    """the ONE JSON line, on the process's real stdout"""
    REAL_STDOUT.write(json.dumps(line) + "\n")
    REAL_STDOUT.flush()


def main():
    # native libraries write banners to fd 1 (NCCL: "NCCL version ..."): keep the real stdout for the JSON line only and
    # send everything else to stderr
    global REAL_STDOUT
    sys.stdout.flush()
    REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5"])
    ap.add_argument("--samples", type=int, default=SAMPLES_PER_GPU, help="samples per GPU per step")
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / FULL_S roofline legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import qo100net as Q
    from qo100net import dist as qd
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libqo100net has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        dist = qd.init_process_group("nccl")

    nspg = args.samples
    wl = workload(args.workload, nspg)
    nf = len(wl.f)
    ctx = Q.Context(device=local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    plan = Q.Plan(ctx, wl.net, wl.f, wl.specs, seed=wl.seed, tols=wl.tols, **wl.hist)
    ncnt = plan.num_counters
    counters = torch.zeros(ncnt, dtype=torch.int64, device="cuda")
    evals_per_step = nspg * nf * world

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        """one pass: this rank's 1e6 samples of global step i, then the counter all-reduce"""
        counters.zero_()
        off = (i * world + rank) * nspg
        plan.launch(off, nspg, counters.data_ptr())
        qd.allreduce_counters(counters)

    sampler = ClockSampler(local) if rank == 0 else None
    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            step(i)
        barrier()
        if sampler:
            sampler.start()
        launches0 = plan.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        ev0.record(stream)
        for i in range(args.steps):
            counters.zero_()
            off = ((args.warmup + i) * world + rank) * nspg
            kev[i][0].record(stream)
            plan.launch(off, nspg, counters.data_ptr())
            kev[i][1].record(stream)
            qd.allreduce_counters(counters)
        ev1.record(stream)
        barrier()
        ms_total = ev0.elapsed_time(ev1)
        kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
        launches = plan.launches - launches0
        last = qd.split_counters(counters.cpu(), len(wl.specs), wl.hist["hist_bins"])
        # end to end through the host-buffer C-ABI call (qo_mc_run): per step H2D of the grid, specs and
        # tolerance table, kernel, D2H of the counters; N > 1 adds the all-reduce of the host result
        h2d = plan.h2d_bytes                     # what qo_plan_create copies: frequency tables, spec masks, device program
        d2h = ncnt * 8
        for i in range(2):
            ctx.mc_run(wl.net, wl.f, wl.specs, wl.seed, nspg, wl.tols, sample_offset=rank * nspg, **wl.hist)
        # N > 1: the host results are summed across ranks through one reused pinned buffer (H2D, all-reduce, D2H per step)
        hc_pin = torch.zeros(2 + len(wl.specs) + wl.hist["hist_bins"], dtype=torch.int64).pin_memory() if dist is not None else None
        hc_dev = torch.zeros_like(hc_pin, device="cuda") if dist is not None else None
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            r = ctx.mc_run(wl.net, wl.f, wl.specs, wl.seed, nspg, wl.tols, sample_offset=((100 + i) * world + rank) * nspg, **wl.hist)
            if dist is not None:
                hc_pin[0], hc_pin[1] = r["n_pass"], r["n_total"]
                hc_pin[2:2 + len(wl.specs)] = torch.from_numpy(r["fail_per_spec"].astype(np.int64))
                hc_pin[2 + len(wl.specs):] = torch.from_numpy(r["hist"].astype(np.int64))
                hc_dev.copy_(hc_pin, non_blocking=True)
                qd.allreduce_counters(hc_dev)
                hc_pin.copy_(hc_dev)
                torch.cuda.current_stream().synchronize()
        barrier()
        e2e_s = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None       # sampled over both timed regions (resident steps and end-to-end steps)

    tm = torch.tensor([ms_total, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kernel_ms = [float(x) for x in tm.cpu()]
    value = evals_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = evals_per_step * args.steps / (e2e_ms * 1e-3)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    flops = plan.flops_per_eval
    peak = ctx.measure_dfma_peak()                      # TFLOP/s, measured now on this GPU
    evals_per_launch = nspg * nf
    achieved = flops * evals_per_launch / (kernel_ms * 1e-3) * 1e-12

    def executed(kernel, ms):
        """executed FP64 work of `kernel` (ncu counts per eval) at the measured launch time: pipe utilisation = FP64-pipe
        instructions/s over the measured DFMA issue rate (peak TFLOP/s / 2); executed TFLOP/s counts DFMA as 2"""
        ex = executed_profile(kernel, wl.name)
        if not ex:
            return None
        ipe = ex["dfma"] + ex["dmul"] + ex["dadd"]
        rate = evals_per_launch / (ms * 1e-3)
        return {"fp64_pipe_instr_per_eval": ipe, "dfma_per_eval": ex["dfma"], "pipe_util": ipe * rate / (peak * 0.5e12),
                "tflops": (2 * ex["dfma"] + ex["dmul"] + ex["dadd"]) * rate * 1e-12,
                "frac_of_peak": (2 * ex["dfma"] + ex["dmul"] + ex["dadd"]) * rate * 1e-12 / peak, "source": ex.get("source")}

    ex = executed(plan.kernel_name, kernel_ms)
    prof = executed_profile(plan.kernel_name, wl.name) or {}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name, "samples_per_gpu_per_step": nspg, "nf": nf, "elements": len(wl.net),
                   "flops_per_eval_alg_v1": flops, "parallelism": "samples sharded x%d, u64 counter all-reduce" % world,
                   "l2": "no flush: the reduce-only path reads < 200 KB of tables by design (bytes/eval ~ 0); every "
                         "step draws a fresh global sample range",
                   "kernel_plan": plan.tf_info if plan.kernel_name == "qo_mc_tf_kernel" else None,
                   "yield_last_step": last["n_pass"] / max(1, last["n_total"])},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "qo_mc_run (host buffers)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "fp64_fma", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": prof.get("dram_bytes_per_launch"), "kernel": plan.kernel_name,
                     "kernel_ms": kernel_ms, "executed": ex,
                     "note": "achieved = ALG-v1 algorithmic flops/eval (SURVEY 8d: a 2x2 complex chain step and a complex divide "
                             "per element) x evals / kernel time.  The transfer-function kernel expands the cascade into real "
                             "polynomials once per sample and evaluates them by Horner per point, so it executes ~8x fewer "
                             "operations than ALG-v1 counts and frac exceeds 1; `executed` is what the FP64 pipe really did "
                             "(ncu instruction counts x the measured rate): pipe_util against the measured DFMA issue rate, "
                             "frac_of_peak with DFMA = 2 flops.  `chain_kernel` is the same job on the straight-line ABCD-chain "
                             "kernel (QO100NET_KERNEL=ladder), whose executed work is close to ALG-v1.",
                     "peak_source": "measured in this run: qo_measure_dfma_peak (8 independent DFMA chains/thread, "
                                    "best of 5); MEASURED_PEAKS.json has no FP64 figure (nominal 37.2 TFLOP/s)"},
    }
    if not args.no_extras and plan.kernel_name == "qo_mc_tf_kernel":
        # the same workload on the straight-line ABCD-chain kernel (north_star (c) as worded), for comparison
        os.environ["QO100NET_KERNEL"] = "ladder"
        try:
            p2 = Q.Plan(ctx, wl.net, wl.f, wl.specs, seed=wl.seed, tols=wl.tols, **wl.hist)
            c2 = torch.zeros(ncnt, dtype=torch.int64, device="cuda")
            ms2 = time_plan(p2, stream, c2, nspg, max(3, args.steps // 2), 2, torch)
            line["roofline"]["chain_kernel"] = {"kernel": p2.kernel_name, "kernel_ms": ms2, "evals_per_s_per_gpu": evals_per_launch / (ms2 * 1e-3),
                                                "alg_v1_tflops": flops * evals_per_launch / (ms2 * 1e-3) * 1e-12,
                                                "alg_v1_frac": flops * evals_per_launch / (ms2 * 1e-3) * 1e-12 / peak,
                                                "executed": executed(p2.kernel_name, ms2)}
            p2.close()
        except Exception as ex2:
            line["roofline"]["chain_kernel"] = {"error": str(ex2)}
        finally:
            os.environ.pop("QO100NET_KERNEL", None)
    if not args.no_extras and world == 1:
        # CPU baseline on a bounded sample of the same workload (rank 0, N = 1 only)
        from oracle import refbind as R
        nthr = R.max_threads()
        n = 40000
        v, _ = cpu_reference(wl, n, 2, 1, nthr)
        v1, _ = cpu_reference(wl, 2000, 1, 0, 1)          # the "single-threaded C loop over the same model" of north_star
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": nthr, "kind": "port",
                                "sample": "%d samples x %d freq, 2 timed passes, OpenMP over samples" % (n, nf),
                                "single_thread": {"value": v1, "unit": UNIT, "cores": 1, "sample": "2000 samples x %d freq, 1 pass" % nf}}
        # HBM-bound mode (BASELINE config 4): full S-matrix written out, 64 B/eval
        try:
            from qo100net import workloads as W
            w4 = W.cfg4(32768, NF)[0]
            buf = torch.empty((4, 32768, NF, 2), dtype=torch.float64, device="cuda")
            p4 = Q.Plan(ctx, w4.net, w4.f, [], seed=w4.seed, tols=w4.tols, mode=Q.MODE_FULL_S)
            with torch.cuda.stream(stream):
                for _ in range(3):
                    p4.launch(0, 32768, None, buf.data_ptr())
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record(stream)
                for _ in range(5):
                    p4.launch(0, 32768, None, buf.data_ptr())
                b.record(stream)
                torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            gbs = 32768 * NF * 64 / (ms * 1e-3) * 1e-9
            # what a pure write stream reaches on this GPU, same buffer (the copy figure of MEASURED_PEAKS.json is half reads)
            with torch.cuda.stream(stream):
                buf.zero_()
                torch.cuda.synchronize()
                a.record(stream)
                for _ in range(3):
                    buf.zero_()
                b.record(stream)
                torch.cuda.synchronize()
            fill_gbs = buf.numel() * 8 / (a.elapsed_time(b) / 3 * 1e-3) * 1e-9
            hbm = None
            try:
                hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            except Exception:
                pass
            line["roofline_hbm"] = {"bound": "hbm", "workload": w4.name + " (32768 samples, 8.6 GB written, > L2)",
                                    "achieved": gbs, "peak": hbm or 6650.0, "unit": "GB/s", "frac": gbs / (hbm or 6650.0),
                                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if hbm else "fallback 6.65 TB/s",
                                    "evals_per_s": 32768 * NF / (ms * 1e-3), "traffic": None,
                                    "write_only_fill_gbs": fill_gbs, "frac_of_write_only_fill": gbs / fill_gbs,
                                    "kernel": p4.kernel_name,
                                    "note": "peak is the COPY bandwidth of MEASURED_PEAKS.json (half reads, half writes); this mode only "
                                            "writes, and a plain fill of the same buffer (write_only_fill_gbs, measured here) is the tighter roof"}
            p4.close()
            del buf
        except Exception as ex:          # the secondary leg must never take the headline line down
            line["roofline_hbm"] = {"error": str(ex)}
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
