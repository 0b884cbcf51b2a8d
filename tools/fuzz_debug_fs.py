#!/usr/bin/env python
"""FULL_S mismatch recorded by tools/fuzz_parity.py: qo_fs_tf_kernel (large launch) vs interpreter (small launch) vs oracle.
   python tools/fuzz_debug_fs.py profiles/fuzz/<file>.json <net>"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python")); sys.path.insert(0, ROOT)
import qo100net as Q
from oracle import refbind as R
d = json.load(open(sys.argv[1]))
m = next(x for x in d["details"] if x["net"] == int(sys.argv[2]))
net = Q.Net.from_elements([(k, p) for k, p in m["elements"]], *m["terminations"])
f = Q.grid_log(m["f0"], m["f1"], m["nf"]) if m["log_grid"] else Q.grid_lin(m["f0"], m["f1"], m["nf"])
tols = [tuple(t) for t in m["tols"]]
ctx = Q.Context(device=0)
a = ctx.mc_run(net, f, [], m["seed"], 1536, tols, sample_offset=m["offset"], mode=Q.MODE_FULL_S, dist=m["dist"])["s"]
b = ctx.mc_run(net, f, [], m["seed"], 64, tols, sample_offset=m["offset"], mode=Q.MODE_FULL_S, dist=m["dist"])["s"]
rs, rl = net.terminations
o = R.mc_run(R.make_elems(net.elements), rs, rl, f, [], R.mc_cfg(m["seed"], 64, tols, sample_offset=m["offset"], dist=m["dist"]), full_s=True)["s"]
for pl in range(4):
    ga, gb, go = np.asarray(a[pl])[:64], np.asarray(b[pl]), np.asarray(o[pl])
    fl = 1e-6 if pl in (1, 2) else 0.02
    for name, x, y in (("fs_tf vs interp", ga, gb), ("fs_tf vs oracle", ga, go), ("interp vs oracle", gb, go)):
        r = np.abs(x - y) / (1e-9 * np.maximum(np.abs(y), fl))
        k = np.unravel_index(np.argmax(r), r.shape)
        print("plane %d %-17s worst ratio %.3g at sample %d f %.6e: %s vs %s (|ref| %.3e)" % (pl, name, r[k], k[0], f[k[1]], x[k], y[k], abs(y[k])))
