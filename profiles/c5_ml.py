#!/usr/bin/env python
"""BASELINE config 5: factorize() maximum-likelihood KL multiplicative updates (R/factorize.R:2-27,
189-212) on the C2-shaped matrix (20k genes x 100k cells, ~8% nonzero), rank 15, w, h ~ U(0,1)
seed 5.  mlnmf_run is called through the C ABI with K1 and K2 > K1 iterations
(Tol = 0: the likelihood rule never fires); the per-iteration time is the difference quotient, so
the upload of w0/h0 and the download of w/h drop out.  One iteration = cell-owner sweep + h update
+ gene-owner sweep + w update + the likelihood readback of the loop (host sync).  Prints one JSON
line.   python profiles/c5_ml.py [--precision 0|1] [--rank 15]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from ccfindr_b200 import synth
from ccfindr_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--rank", type=int, default=15); ap.add_argument("--precision", type=int, default=0)
ap.add_argument("--k1", type=int, default=10); ap.add_argument("--k2", type=int, default=40)
ap.add_argument("--cells", type=int, default=100000)
a = ap.parse_args()
wl = bench.WORKLOADS["c2"]; n, m, r = wl["n"], a.cells, a.rank
dev = torch.device("cuda", 0)
colptr, rowidx, values, _ = synth.tenx_like_device(n, m, wl["r_true"], wl["density"], wl["seed"], dev)
nnz = int(rowidx.numel())
w0, h0 = synth.uniform_init(n, m, r, seed=5)
eng = Engine.from_device_csc(n, m, nnz, colptr, rowidx, values)
eng.set_precision(a.precision)
eng.ml_run(w0, h0, Itmax=3, Tol=0.0)                      # warm-up: layout build, allocations
ts = {}
for k in (a.k1, a.k2, a.k1, a.k2):
    torch.cuda.synchronize(); t0 = time.time()
    res = eng.ml_run(w0, h0, Itmax=k, Tol=0.0)
    torch.cuda.synchronize(); ts.setdefault(k, []).append(time.time() - t0)
    assert res["niter"] == k and np.isfinite(res["lik"])
t_iter = (min(ts[a.k2]) - min(ts[a.k1])) / (a.k2 - a.k1)
sp = 4 if a.precision else 8
b_alg = 2 * nnz * 8 + 2 * 8 * (m + 1) + 3 * r * m * sp + 3 * n * r * sp     # SURVEY.md 8(d), ML row
peak, src = bench.peaks()
print(json.dumps({
    "workload": "C5: factorize() ML path, rank %d, 20k genes x %d cells, nnz %d" % (r, m, nnz),
    "metric": "ML-NMF nnz*rank updates/s per iteration", "value": nnz * r / t_iter,
    "ms_per_iteration": t_iter * 1e3, "precision": "fp32-storage" if a.precision else "fp64",
    "iterations_timed": [a.k1, a.k2], "seconds": {str(k): v for k, v in ts.items()},
    "lik_last": res["lik"], "layout": eng.layout_info(),
    "roofline": {"bound": "hbm", "algorithmic_bytes_per_iteration": b_alg,
                 "achieved": b_alg / t_iter / 1e9, "peak": peak, "unit": "GB/s",
                 "frac": b_alg / t_iter / 1e9 / peak, "peak_source": src}}))
