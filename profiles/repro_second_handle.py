#!/usr/bin/env python
"""Two handles in one process on matrices of different sizes: the bound of the second must equal
that of a fresh process (stale pool memory / counters would show here)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from ccfindr_b200 import synth
from ccfindr_b200.engine import Engine

wl = bench.WORKLOADS["c3"]; n, r = wl["n"], wl["rank"]
dev = torch.device("cuda", 0)
def run(m, host=False):
    colptr, rowidx, values, _ = synth.tenx_like_device(n, m, wl["r_true"], wl["density"], wl["seed"], dev)
    w0, h0 = bench.init_factors(n, m, r, seed=1000 * r + 1)
    if host:
        import scipy.sparse as sp
        csc = sp.csc_matrix((values.double().cpu().numpy(), rowidx.cpu().numpy(), colptr.cpu().numpy()), shape=(n, m))
        eng = Engine(csc, device=0)
    else:
        eng = Engine.from_device_csc(n, m, int(rowidx.numel()), colptr, rowidx, values)
    eng.set_state(w0, h0)
    out = eng.run(bench.HYPER, Itmax=3, Tol=0.0)
    eng.close()
    return [float(v) for v in out["lkh_trace"]]
order = sys.argv[1:] or ["40000", "80000", "40000"]
for o in order:
    host = o.endswith("h"); m = int(o.rstrip("h"))
    print(o, run(m, host), flush=True)
