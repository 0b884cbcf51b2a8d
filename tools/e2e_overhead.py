import sys, os, time
sys.path.insert(0, 'qo-100-tools_b200/python'); sys.path.insert(0, '.')
import qo100net as Q
from qo100net import workloads as W
ctx = Q.Context(device=0)
w = W.cfg2()
for n in (64, 1000000):
    for i in range(5):
        ctx.mc_run(w.net, w.f, w.specs, w.seed, n, w.tols, **w.hist)
    t = time.perf_counter()
    reps = 200 if n < 1000 else 20
    ks = 0.0
    for i in range(reps):
        r = ctx.mc_run(w.net, w.f, w.specs, w.seed, n, w.tols, sample_offset=i * n, **w.hist)
        ks += r["seconds"]
    dt = (time.perf_counter() - t) / reps
    print("n=%d: %.1f us per call wall, kernel %.1f us, overhead %.1f us" % (n, dt * 1e6, ks / reps * 1e6, (dt - ks / reps) * 1e6))
