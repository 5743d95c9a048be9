#!/usr/bin/env python
"""BASELINE config 4: rank sweep 2..30 x nrun restarts on the C2 matrix (independent jobs, replicas
only).  One process per GPU under torchrun (or a single process): jobs are LPT-scheduled over the
ranks, every rank holds the full matrix, K iterations per job with fixed hypers.  Prints one JSON
line with the aggregate nnz*rank updates/s and the per-rank iteration times."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from ccfindr_b200 import api, synth
from ccfindr_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10); ap.add_argument("--nrun", type=int, default=5)
ap.add_argument("--rmin", type=int, default=2); ap.add_argument("--rmax", type=int, default=30)
ap.add_argument("--cells", type=int, default=100000); ap.add_argument("--precision", type=int, default=0)
ap.add_argument("--device-init", action="store_true", help="draw w0, h0 on the GPU (vbnmf_init_random)")
a = ap.parse_args()
rank, local, world = bench.env_rank()
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
wl = bench.WORKLOADS["c2"]; n = wl["n"]; m = a.cells
colptr, rowidx, values, _ = synth.tenx_like_device(n, m, wl["r_true"], wl["density"], wl["seed"], dev)
nnz = int(rowidx.numel())
jobs = [(r, i) for i in range(1, a.nrun + 1) for r in range(a.rmin, a.rmax + 1)]
mine = api.lpt_schedule([float(r) for r, _ in jobs], world)[rank]
eng = Engine.from_device_csc(n, m, nnz, colptr, rowidx, values, device=local)
eng.set_precision(a.precision)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time(); per_rank = {}; units = 0.0
for j in mine:
    r, i = jobs[j]
    if a.device_init:
        eng.init_random(r, bench.HYPER, 1000 * r + i)
    else:
        w0, h0 = synth.random_init(n, m, r, bench.HYPER, 1000 * r + i)  # seeds 1000*rank + run
        eng.set_state(w0, h0)
    res = eng.bench_iterations(bench.HYPER, a.iters)
    assert np.isfinite(res["lkh"]), (r, i)
    per_rank.setdefault(r, []).append(res["ms_total"] / a.iters)
    units += nnz * r * a.iters
torch.cuda.synchronize()
wall = time.time() - t0
tt = torch.tensor([wall, units], dtype=torch.float64, device=dev)
if world > 1:
    wmax = tt[0:1].clone(); dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
    usum = tt[1:2].clone(); dist.all_reduce(usum)
    wall, units = float(wmax.item()), float(usum.item())
if rank == 0:
    print(json.dumps({"workload": "C4: ranks %d..%d x %d restarts, %d iterations each, 20k x %d, nnz %d"
                      % (a.rmin, a.rmax, a.nrun, a.iters, m, nnz), "n_gpus": world, "jobs": len(jobs),
                      "precision": a.precision, "device_init": bool(a.device_init),
                      "wall_s_incl_init_upload_and_layout_builds": wall,
                      "aggregate_updates_per_s": units / wall,
                      "ms_per_iteration_by_rank_on_rank0": {str(k): round(float(np.mean(v)), 3) for k, v in sorted(per_rank.items())}}))
if world > 1:
    dist.destroy_process_group()
