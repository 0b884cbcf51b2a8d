import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python")); sys.path.insert(0, ROOT)
import torch
import qo100net as Q
from qo100net import workloads as W
ctx = Q.Context(device=0)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
n = 400000
wl = W.cfg2(n, 4096); nf = 4096
base = None
for v in [int(x) for x in sys.argv[1].split(",")]:
    os.environ["QO100NET_LAD_VARIANT"] = str(v)
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
    print("variant %d %s %.3f ms %.3e evals/s same=%s" % (v, plan.kernel_name, best, n * nf / best * 1e3, np.array_equal(got, base)), flush=True)
    plan.close()
