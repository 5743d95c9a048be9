#!/usr/bin/env python
"""BASELINE config 4 run to convergence through the front end: vb_factorize(ranks = 2..30,
nrun = 5) with the reference's defaults (Tol = 1e-5, hyper-parameter updates from iteration 11)
on the C2 matrix, then optimal_rank() on the result (R/bayesian.R:229-301, R/utils2.R:59-95).
Single process: the 145 factorizations run one after the other on one GPU; under torchrun they
are farmed over the GPUs (LPT by rank, every GPU holds the matrix; the role of Rmpi::mpi.applyLB,
R/bayesian.R:263).  Initial factors are drawn on the device.  Prints one JSON line.

    python profiles/c4_converged.py [--rmax 30] [--nrun 5] [--itmax 10000]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, scipy.sparse as sp, torch
import bench
from ccfindr_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--rmin", type=int, default=2); ap.add_argument("--rmax", type=int, default=30)
ap.add_argument("--nrun", type=int, default=5); ap.add_argument("--itmax", type=int, default=10000)
ap.add_argument("--cells", type=int, default=100000); ap.add_argument("--precision", type=int, default=0)
a = ap.parse_args()
rank, local, world = bench.env_rank()
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
wl = bench.WORKLOADS["c2"]; n = wl["n"]; m = a.cells
colptr, rowidx, values, _ = synth.tenx_like_device(n, m, wl["r_true"], wl["density"], wl["seed"], dev)
X = sp.csc_matrix((values.double().cpu().numpy(), rowidx.cpu().numpy(), colptr.cpu().numpy()), shape=(n, m))
X.has_sorted_indices = True
del colptr, rowidx, values
torch.cuda.empty_cache()
ranks = list(range(a.rmin, a.rmax + 1))
s = api.scNMFSet(X)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time()
s = api.vb_factorize(s, ranks=ranks, nrun=a.nrun, verbose=0, Itmax=a.itmax, Tol=1e-5,
                     connectivity=False, device=local, device_init=True, precision=a.precision,
                     parallel=world > 1, unif_stop=False)
torch.cuda.synchronize()
wall = time.time() - t0
if rank == 0:
    its = s.metadata["niter"]                      # {(irun, rank): iterations}
    units = sum(X.nnz * r * it for (_, r), it in its.items())
    try:
        opt = api.optimal_rank(s)
    except Exception as e:
        opt = {"error": str(e)}
    print(json.dumps({
        "workload": "C4 to convergence: vb_factorize(ranks=%d..%d, nrun=%d, Tol=1e-5, Itmax=%d) on 20k x %d, "
                    "nnz %d, device-drawn initial factors" % (a.rmin, a.rmax, a.nrun, a.itmax, m, X.nnz),
        "n_gpus": world, "jobs": len(its), "precision": a.precision, "wall_s": wall,
        "iterations_total": int(sum(its.values())),
        "iterations_by_rank_mean": {str(r): float(np.mean([it for (_, rr), it in its.items() if rr == r]))
                                    for r in ranks},
        "aggregate_updates_per_s": units / wall,
        "lml_by_rank": {str(r): float(v) for r, v in zip(s.measure["rank"], s.measure["lml"])},
        "optimal_rank": {k: (v if isinstance(v, str) else (None if v is None else float(v)))
                         for k, v in opt.items()} if isinstance(opt, dict) else str(opt)}))
if world > 1:
    dist.destroy_process_group()
