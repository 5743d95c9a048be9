#!/usr/bin/env python
"""BASELINE config 1 (the reference's own CPU-runnable case): vb_factorize-style runs on
simulate_whx 1,000 genes x 200 cells, rank 3, and on the 1,030 x 450 PBMC fixture, rank 5.
Small problems are launch-bound: the device loop replays batches of 8 iterations as a CUDA graph
(VBNMF_NO_GRAPH=1 launches the kernels one by one).  Prints ms per iteration for both."""
import json, os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, scipy.sparse as sp, torch
from ccfindr_b200 import synth
from ccfindr_b200.engine import Engine
from conftest import load_counts

hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
cases = {"C1 simulate_whx 1000x200 r=3": (sp.csc_matrix(synth.simulate_whx(1000, 200, 3, seed=1)["x"]), 3),
         "PBMC fixture 1030x450 r=5": (load_counts("pbmc"), 5)}
out = {}
for name, (x, r) in cases.items():
    n, m = x.shape
    w0, h0 = synth.random_init(n, m, r, hyper, seed=3)
    res = {}
    with Engine(x) as eng:
        for mode in ("graph", "launches", "graph", "launches"):
            if mode == "launches":
                os.environ["VBNMF_NO_GRAPH"] = "1"
            else:
                os.environ.pop("VBNMF_NO_GRAPH", None)
            eng.set_state(w0, h0)
            torch.cuda.synchronize(); t0 = time.time()
            rr = eng.run(hyper, Itmax=800, Tol=0.0)
            torch.cuda.synchronize(); dt = time.time() - t0
            res.setdefault(mode, []).append(dt / rr["niter"] * 1e3)
            res["lml_" + mode] = rr["lml"]
    out[name] = {"ms_per_iteration_graph": min(res["graph"]), "ms_per_iteration_launches": min(res["launches"]),
                 "nnz": int(x.nnz), "same_result": res["lml_graph"] == res["lml_launches"]}
print(json.dumps(out))
