#!/usr/bin/env python
"""Small driver for ncu: generate a bench workload and run a few iterations in one precision.
   python profiles/prof_run.py --workload c2 --precision 0|1 --iters 3"""
import argparse, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np, torch
import bench
from ccfindr_b200 import synth
from ccfindr_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2"); ap.add_argument("--precision", type=int, default=0)
ap.add_argument("--iters", type=int, default=3); ap.add_argument("--cells", type=int, default=0)
ap.add_argument("--rank", type=int, default=0)
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
n, r = wl["n"], (a.rank or wl["rank"])
m = a.cells or wl.get("m_per_gpu") or wl["m_total"]
dev = torch.device("cuda", 0)
colptr, rowidx, values, _ = synth.tenx_like_device(n, m, wl["r_true"], wl["density"], wl["seed"], dev)
w0, h0 = bench.init_factors(n, m, r, seed=1000 * r + 1)
eng = Engine.from_device_csc(n, m, int(rowidx.numel()), colptr, rowidx, values)
eng.set_precision(a.precision)
eng.set_state(w0, h0)
eng.bench_iterations(bench.HYPER, 3)
res = eng.bench_iterations(bench.HYPER, a.iters)
print({k: (round(v / a.iters, 4) if k.startswith('ms_') else v) for k, v in res.items()}, eng.layout_info())
