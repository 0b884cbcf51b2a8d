#!/usr/bin/env python
"""Development sweep of transfer-function-kernel launch shapes (library built with `make EXTRA=-DQO_TF_EXPERIMENT`):
   python tools/tf_sweep.py cfg2|cfg5|cfg5p  pp:variant[,pp:variant...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python")); sys.path.insert(0, ROOT)
import torch
import qo100net as Q
from qo100net import workloads as W
ctx = Q.Context(device=0)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
n = 400000
wl = getattr(W, sys.argv[1])(n, 4096); nf = 4096
base = None
for item in sys.argv[2].split(","):
    pp, v = item.split(":")
    os.environ["QO100NET_LAD_VARIANT"] = v
    if pp != "0": os.environ["QO100NET_TF_PP"] = pp
    else: os.environ.pop("QO100NET_TF_PP", None)
    plan = Q.Plan(ctx, wl.net, wl.f, wl.specs, seed=wl.seed, tols=wl.tols, **wl.hist)
    cnt = torch.zeros(plan.num_counters, dtype=torch.int64, device="cuda")
    with torch.cuda.stream(stream):
        plan.launch(0, n, cnt.data_ptr()); torch.cuda.synchronize()
        got = cnt.cpu().numpy().copy()
        best = 1e9
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(3): plan.launch((i + 1) * n, n, cnt.data_ptr())
            b.record(stream); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 3)
    if base is None: base = got
    print("pp %s variant %s %s %.3f ms %.3e evals/s same=%s" % (pp, v, plan.kernel_name, best, n * nf / best * 1e3, np.array_equal(got, base)), flush=True)
    plan.close()
