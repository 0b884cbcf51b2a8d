import sys, os
sys.path.insert(0, 'qo-100-tools_b200/python'); sys.path.insert(0, '.')
import qo100net as Q
from qo100net import workloads as W
ctx = Q.Context(device=0)
w = W.cfg3(400000)
for i in range(4):
    r = ctx.mc_run(w.net, w.f, w.specs, w.seed, 400000, w.tols, **w.hist)
print(r["seconds"], 400000*3/r["seconds"])
