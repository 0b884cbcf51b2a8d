#!/usr/bin/env python
"""Config 3 (microstrip PA low-pass, yield at 2.4 / 4.8 / 7.2 GHz) on the thread-per-board kernel and on the item-per-thread kernel
(QO100NET_USTRIP=item): time, evals/s, counters.   python tools/cfg3_run.py [samples]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "qo-100-tools_b200", "python"))
sys.path.insert(0, ROOT)
import numpy as np
import qo100net as Q
from qo100net import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
ctx = Q.Context(device=0)
w = W.cfg3(n)
out = {}
for label, env in (("board", None), ("item", "item")):
    if env:
        os.environ["QO100NET_USTRIP"] = env
    else:
        os.environ.pop("QO100NET_USTRIP", None)
    plan = Q.Plan(ctx, w.net, w.f, w.specs, seed=w.seed, tols=w.tols, **w.hist)
    kname = plan.kernel_name
    plan.close()
    ctx.mc_run(w.net, w.f, w.specs, w.seed, 20000, w.tols, **w.hist)
    best = None
    for i in range(3):
        r = ctx.mc_run(w.net, w.f, w.specs, w.seed, n, w.tols, **w.hist)
        best = r if best is None or r["seconds"] < best["seconds"] else best
    out[label] = {"kernel": kname, "seconds": best["seconds"], "evals_per_s": n * 3 / best["seconds"], "n_pass": int(best["n_pass"]),
                  "fail_per_spec": [int(v) for v in best["fail_per_spec"]], "hist_sum": int(np.sum(best["hist"])), "hist": [int(v) for v in best["hist"]]}
os.environ.pop("QO100NET_USTRIP", None)
out["counters_equal"] = bool(out["board"]["n_pass"] == out["item"]["n_pass"] and out["board"]["hist"] == out["item"]["hist"]
                             and out["board"]["fail_per_spec"] == out["item"]["fail_per_spec"])
for k in ("board", "item"):
    out[k].pop("hist")
print(json.dumps(out, indent=1))
