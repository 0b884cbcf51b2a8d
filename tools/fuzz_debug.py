#!/usr/bin/env python
"""Re-run one mismatch recorded by tools/fuzz_parity.py on every kernel and on the oracle.
   python tools/fuzz_debug.py gpurun_out/fuzz_r02_seed5.json 119"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)
import qo100net as Q
from oracle import refbind as R

d = json.load(open(sys.argv[1]))
m = next(x for x in d["details"] if x["net"] == int(sys.argv[2]))
net = Q.Net.from_elements([(k, p) for k, p in m["elements"]], *m["terminations"])
f = Q.grid_log(m["f0"], m["f1"], m["nf"]) if m["log_grid"] else Q.grid_lin(m["f0"], m["f1"], m["nf"])
tols = [tuple(t) for t in m["tols"]]
specs = [tuple(s) for s in m["specs"]]
ctx = Q.Context(device=0)
n, off = m["n"], m["offset"]
print(Q.plan_analyze(net, f, specs, tols, dist=m["dist"], **m["hist"]))
for force in (None, "tf", "ladder", "interp"):
    if force:
        os.environ["QO100NET_KERNEL"] = force
    else:
        os.environ.pop("QO100NET_KERNEL", None)
    plan = Q.Plan(ctx, net, f, specs, seed=m["seed"], tols=tols, dist=m["dist"], **m["hist"])
    plan.launch(off, n)
    r = plan.read()
    print("%-8s %-22s n_pass %d fails %s" % (force, plan.kernel_name, r["n_pass"], list(r["fail_per_spec"])))
    plan.close()
os.environ.pop("QO100NET_KERNEL", None)
rs, rl = net.terminations
o = R.mc_run(R.make_elems(net.elements), rs, rl, f, specs, R.mc_cfg(m["seed"], n, tols, sample_offset=off, dist=m["dist"], **m["hist"]), nthreads=R.max_threads())
print("oracle   n_pass %d fails %s" % (o["n_pass"], list(o["fail_per_spec"])))
# how close do the samples sit to the threshold?  (interpreter FULL_S on the first 20000 samples)
ns = min(n, 20000)
fs = ctx.mc_run(net, f, [], m["seed"], ns, tols, sample_offset=off, mode=Q.MODE_FULL_S, dist=m["dist"])["s"]
for s in specs:
    band = (f >= s[1]) & (f <= s[2])
    a = 20 * np.log10(np.abs(fs[0 if s[0] == 3 else 1][:, band]))
    v = a.min(axis=1) if s[0] == 1 else a.max(axis=1)
    dist = np.abs(v - s[3])
    print("spec", s, "values: min %.6f median %.6f max %.6f; |value - limit| < 1e-6 dB: %d, < 1e-3 dB: %d of %d" %
          (v.min(), np.median(v), v.max(), int((dist < 1e-6).sum()), int((dist < 1e-3).sum()), ns))
# value path against sign path: the same spec tracked through the histogram (value of n2 / dd) on both transfer-function kernels
s0 = specs[0]
for force in (None, "tf"):
    if force:
        os.environ["QO100NET_KERNEL"] = force
    else:
        os.environ.pop("QO100NET_KERNEL", None)
    plan = Q.Plan(ctx, net, f, specs, seed=m["seed"], tols=tols, dist=m["dist"], hist_bins=1024, hist_spec=0, hist_lo=s0[3] - 1.5, hist_hi=s0[3] + 1.5)
    plan.launch(off, n)
    r = plan.read()
    h = np.array(r["hist"])
    print("hist run %-4s %-18s n_pass %d fails %s  hist checksum %d  below-limit mass %d" % (force, plan.kernel_name, r["n_pass"], list(r["fail_per_spec"]),
          int((h * np.arange(1024)).sum()), int(h[:512].sum())))
    plan.close()
os.environ.pop("QO100NET_KERNEL", None)
# locate the samples on which the selected kernel and the interpreter disagree (bisection over sample ranges; only meaningful when
# the selected kernel does not depend on the launch size), and how far from the limit the interpreter puts them
def run(force, o, cnt):
    if force:
        os.environ["QO100NET_KERNEL"] = force
    else:
        os.environ.pop("QO100NET_KERNEL", None)
    plan = Q.Plan(ctx, net, f, specs, seed=m["seed"], tols=tols, dist=m["dist"], **m["hist"])
    plan.launch(o, cnt)
    r = plan.read()
    plan.close()
    return [int(r["n_pass"])] + [int(v) for v in r["fail_per_spec"]]
print("hist of the recorded job:", m["hist"])
if m["kernel"] == "qo_mc_tf_kernel":
    found = []
    stack = [(off, n)]
    while stack and len(found) < 4:
        o, c = stack.pop()
        if run("tf", o, c) == run("interp", o, c):
            continue
        if c == 1:
            found.append(o)
            continue
        h = c // 2
        stack += [(o, h), (o + h, c - h)]
    os.environ.pop("QO100NET_KERNEL", None)
    for o in found:
        fs1 = ctx.mc_run(net, f, [], m["seed"], 1, tols, sample_offset=o, mode=Q.MODE_FULL_S, dist=m["dist"])["s"]
        for si, s in enumerate(specs):
            band = (f >= s[1]) & (f <= s[2])
            a = 20 * np.log10(np.abs(fs1[0 if s[0] == 3 else 1][0, band]))
            v = a.min() if s[0] == 1 else a.max()
            print("sample %d spec %d: value %.12f dB, limit %.12f dB, margin %.3e dB; tf %s interp %s" % (o, si, v, s[3], v - s[3], run("tf", o, 1), run("interp", o, 1)))
