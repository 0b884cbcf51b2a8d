import sys
sys.path.insert(0, "qo-100-tools_b200/python"); sys.path.insert(0, ".")
import numpy as np, torch, qo100net as Q
from qo100net import workloads as W
ctx = Q.Context(device=0); stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
n = 32768
buf = torch.empty((4, n, 4096, 2), dtype=torch.float64, device="cuda")
for name, w in (("cfg2 net", W.cfg2(n, 4096)), ("cfg5 net", W.cfg5(n, 4096)), ("gpsdo 15M", W.cfg4(n, 4096)[1])):
    plan = Q.Plan(ctx, w.net, w.f, [], seed=1, tols=w.tols, mode=Q.MODE_FULL_S)
    with torch.cuda.stream(stream):
        plan.launch(0, n, None, buf.data_ptr()); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(3): plan.launch(0, n, None, buf.data_ptr())
        b.record(stream); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print("%-10s FULL_S %-20s %.3f ms  %.0f GB/s" % (name, plan.kernel_name, ms, n * 4096 * 64 / ms * 1e-6))
    plan.close()
