"""End-to-end latency of the host-buffer entry points for tiny jobs (development probe)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python")); sys.path.insert(0, ROOT)
import qo100net as Q
from qo100net import workloads as W
ctx = Q.Context(device=0)
w1, w2 = W.cfg1(), W.cfg2()
def t(fn, n=30):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
print("sweep cfg1 1024 pts         %.3f ms" % t(lambda: ctx.sweep(w1.net, w1.f)))
print("sweep cfg1 64 pts           %.3f ms" % t(lambda: ctx.sweep(w1.net, w1.f[:64])))
print("sweep cfg1 1024 pts + gd    %.3f ms" % t(lambda: ctx.sweep(w1.net, w1.f, gd=True)))
print("sweep cfg2 4096 pts         %.3f ms" % t(lambda: ctx.sweep(w2.net, w2.f)))
print("mc_run cfg2 64 samples      %.3f ms" % t(lambda: ctx.mc_run(w2.net, w2.f, w2.specs, 1, 64, w2.tols, **w2.hist)))
pa = W.pa_lpf_net()
f = Q.grid_lin(1e7, 1e10, 5000)
print("sweep PA-LPF 5000 pts       %.3f ms" % t(lambda: ctx.sweep(pa, f), 10))
for name, net, fc in W.gpsdo_bank():
    fg = Q.grid_log(fc / 2.5, fc * 6.25, 1024)
    print("sweep gpsdo %-4s 1024 pts     %.3f ms" % (name, t(lambda: ctx.sweep(net, fg))))
r = ctx.mc_run(w1.net, w1.f, [], 1, 1, [], mode=Q.MODE_FULL_S)
print("cfg1 FULL_S n=1 kernel seconds", r["seconds"])
ser = Q.Net.from_elements([(Q.SER_LC_SER, [33e-9, 4.7e-12])], 50.0, 50.0)
print("single SER_LC_SER             %.3f ms" % t(lambda: ctx.sweep(ser, w1.f)))
sh = Q.Net.from_elements([(Q.SHUNT_LC_PAR, [4.7e-9, 33e-12])], 50.0, 50.0)
print("single SHUNT_LC_PAR           %.3f ms" % t(lambda: ctx.sweep(sh, w1.f)))
print("cfg1 at cfg2 grid             %.3f ms" % t(lambda: ctx.sweep(w1.net, w2.f)))
print("cfg2 at cfg1 grid             %.3f ms" % t(lambda: ctx.sweep(w2.net, w1.f)))
