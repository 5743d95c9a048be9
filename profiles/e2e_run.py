#!/usr/bin/env python
"""Times the host-buffer path of one factorization (VBNMF_TIMING=1 prints the stages)."""
import os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, scipy.sparse as sp, torch
import bench
from ccfindr_b200 import synth
from ccfindr_b200.engine import Engine
wl = bench.WORKLOADS["c2"]; n, r, m = wl["n"], wl["rank"], wl["m_per_gpu"]
dev = torch.device("cuda", 0)
colptr, rowidx, values, _ = synth.tenx_like_device(n, m, wl["r_true"], wl["density"], wl["seed"], dev)
csc = sp.csc_matrix((values.cpu().numpy().astype(np.float64), rowidx.cpu().numpy(), colptr.cpu().numpy()), shape=(n, m))
w0, h0 = bench.init_factors(n, m, r, seed=1000 * r + 1)
for rep in range(2):
    t0 = time.time(); eng = Engine(csc); t1 = time.time()
    eng.set_state(w0, h0); t2 = time.time()
    out = eng.run(bench.HYPER, Itmax=20, Tol=0.0); t3 = time.time()
    st = eng.get_state(("ew", "eh")); t4 = time.time()
    eng.close()
    print("rep %d: Engine() %.3f  set_state %.3f  run(20) %.3f  get_state %.3f  total %.3f s" % (rep, t1-t0, t2-t1, t3-t2, t4-t3, t4-t0), flush=True)
