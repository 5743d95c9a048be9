import os, sys, time, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, scipy.sparse as sp, torch
from ccfindr_b200 import synth
from ccfindr_b200.engine import Engine
from conftest import load_counts
hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
cases = {"C1": (sp.csc_matrix(synth.simulate_whx(1000, 200, 3, seed=1)["x"]), 3), "PBMC": (load_counts("pbmc"), 5),
         "mid 5000x3000 r=8": (synth.fix_empty(sp.random(5000, 3000, density=0.08, format="csc", random_state=np.random.default_rng(1), data_rvs=lambda k: np.ones(k)), 1), 8)}
for name, (x, r) in cases.items():
    w0, h0 = synth.random_init(*x.shape, r, hyper, seed=3)
    for T in ("0", "64", "128", "256", "512"):
        if T == "0": os.environ.pop("VBNMF_TILE_ROWS", None)
        else: os.environ["VBNMF_TILE_ROWS"] = T
        with Engine(x) as eng:
            eng.set_state(w0, h0); eng.run(hyper, Itmax=100, Tol=0.0)
            best = 1e9
            for rep in range(2):
                eng.set_state(w0, h0)
                torch.cuda.synchronize(); t0 = time.time()
                rr = eng.run(hyper, Itmax=800, Tol=0.0)
                torch.cuda.synchronize(); best = min(best, (time.time() - t0) / rr["niter"] * 1e3)
            print(name, "T", T, "layout", eng.layout_info()["tile_rows"], "ms/iter %.4f" % best, flush=True)
