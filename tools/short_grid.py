import os, sys
sys.path.insert(0, "qo-100-tools_b200/python"); sys.path.insert(0, ".")
import numpy as np, torch, qo100net as Q
from qo100net import workloads as W
ctx = Q.Context(device=0); stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
for wn in ("cfg2", "cfg5"):
    for nf in (16, 51, 101, 201, 401, 1001):
        w = getattr(W, wn)(0, nf); n = 1000000
        plan = Q.Plan(ctx, w.net, w.f, w.specs, seed=w.seed, tols=w.tols, **w.hist)
        cnt = torch.zeros(plan.num_counters, dtype=torch.int64, device="cuda")
        with torch.cuda.stream(stream):
            plan.launch(0, n, cnt.data_ptr()); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(3): plan.launch((i + 1) * n, n, cnt.data_ptr())
            b.record(stream); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(wn, "nf", nf, plan.kernel_name, "%.3f ms  %.3e evals/s  %.3e samples/s" % (ms, n * nf / ms * 1e3, n / ms * 1e3), flush=True)
        plan.close()
