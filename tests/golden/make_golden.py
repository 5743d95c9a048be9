"""Generates the golden vectors in this directory.  Run ONCE in the build container (it needs
/root/reference and the compiled reference binary oracle/_ref/libccfindr_ref.so):

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

What produces the numbers:
  * every vbnmf_update step is executed by the REFERENCE'S OWN src/vbnmf_update.cpp:16-102,
    compiled in place (oracle/Makefile target `ref`) against the stand-in Eigen/Rcpp/GSL headers
    of oracle/shim/ (none of those libraries is installed here);
  * the loop around it (R/bayesian.R:336-352) and hyper_update (R/bayesian.R:2-53) exist only as R
    source and there is no R interpreter here, so they are driven by the line-by-line Python
    restatement in oracle/oracle_dense.py.
Inputs: the reference's bundled PBMC subset inst/extdata/matrix.mtx (stored here as
pbmc_counts.npz, CSC) and simulate_whx-recipe matrices (R/utils.R:826-846) from NumPy Philox
streams (R's RNG cannot be reproduced without R); initial factors as vb_init 'random'
(R/bayesian.R:111-115) from NumPy Philox streams.  All stored so that nothing is regenerated at
test time.
"""
import os
import sys

import numpy as np
import scipy.io
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from ccfindr_b200 import synth  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from oracle import oracle_dense as od  # noqa: E402

REF = "/root/reference"


def save_csc(path, csc):
    csc = csc.tocsc()
    csc.sort_indices()
    data = csc.data
    assert np.all(data == np.round(data)) and data.max() < 65536
    np.savez_compressed(path, shape=np.array(csc.shape, dtype=np.int64),
                        indptr=csc.indptr.astype(np.int64), indices=csc.indices.astype(np.int32),
                        data=data.astype(np.uint16))


def run_case(name, X, rank, seed, hyper0, **kw):
    Xd = np.asarray(X.todense(), dtype=np.float64)
    n, m = Xd.shape
    w0, h0 = synth.random_init(n, m, rank, hyper0, seed)
    wh, hyper, lk0, it, trace, htrace, reason = od.vb_run_one_rank(
        Xd, w0, h0, hyper0, update=ob.ref_vbnmf_update, **kw)
    out = dict(w0=w0, h0=h0, hyper0=np.array([hyper0[k] for k in ("aw", "bw", "ah", "bh")]),
               lkh_trace=trace, hyper_trace=htrace, lml=lk0, niter=it, stop_reason=reason,
               hyper=np.array([hyper[k] for k in ("aw", "bw", "ah", "bh")]),
               cid=od.cluster_id(wh["eh"]))
    for k in ("lw", "lh", "ew", "eh", "dw", "dh"):
        out[k] = np.asarray(wh[k])
    for k, v in kw.items():
        out["cfg_" + k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "niter", it, "reason", reason, "lml", lk0, "hyper", out["hyper"])


def step_case(name, X, rank, seed, hyper, fudge, warm=0):
    Xd = np.asarray(X.todense(), dtype=np.float64)
    n, m = Xd.shape
    w0, h0 = synth.random_init(n, m, rank, dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0), seed)
    wh = od.vb_init_from(w0, h0)
    for _ in range(warm):  # move away from the init so that eh != lh
        wh = ob.ref_vbnmf_update(Xd, wh, hyper, fudge)
    inp = {("in_" + k): np.asarray(wh[k]) for k in ("lw", "lh", "ew", "eh")}
    res = ob.ref_vbnmf_update(Xd, wh, hyper, fudge)
    out = {("out_" + k): np.asarray(res[k]) for k in ("lw", "lh", "ew", "eh", "dw", "dh")}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), lkh=res["lkh"], fudge=fudge,
                        hyper=np.array([hyper[k] for k in ("aw", "bw", "ah", "bh")]), **inp, **out)
    print(name, "lkh", res["lkh"])


def main():
    assert ob.ref_lib() is not None, "build oracle/_ref first (make -C oracle)"
    print("reference source:", ob.ref_lib().ref_source_path().decode())
    pbmc = sp.csc_matrix(scipy.io.mmread(os.path.join(REF, "inst/extdata/matrix.mtx")))
    save_csc(os.path.join(HERE, "pbmc_counts.npz"), pbmc)
    mats = {"pbmc": pbmc}
    for seed in (1, 2, 3):
        x = synth.simulate_whx(nrow=1000, ncol=200, rank=3, seed=seed)["x"]
        mats["c1s%d" % seed] = sp.csc_matrix(x)
        save_csc(os.path.join(HERE, "c1s%d_counts.npz" % seed), mats["c1s%d" % seed])
    tiny = sp.csc_matrix(synth.simulate_whx(nrow=40, ncol=25, rank=2, aw=0.5, ah=0.5, seed=7)["x"])
    mats["tiny"] = tiny
    save_csc(os.path.join(HERE, "tiny_counts.npz"), tiny)

    h1 = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
    eps = od.EPS
    # single reference steps
    step_case("step_tiny_r2", tiny, 2, 11, h1, eps)
    step_case("step_c1s1_r3", mats["c1s1"], 3, 12, h1, eps, warm=2)
    step_case("step_pbmc_r5_hyp", pbmc, 5, 13, dict(aw=0.5, bw=2.0, ah=0.3, bh=0.7), 1e-10, warm=1)
    # loops: Itmax 40 crosses n0 = 10 so hyper updates are exercised
    kw = dict(Itmax=40, Tol=1e-5, n0=10, dn=1)
    for seed in (1, 2, 3):
        run_case("run_c1s%d_r3" % seed, mats["c1s%d" % seed], 3, seed, h1, **kw)
    run_case("run_c1s1_r2", mats["c1s1"], 2, 21, h1, **kw)
    run_case("run_c1s1_r5", mats["c1s1"], 5, 22, h1, **kw)
    for rank in (2, 3, 5):
        run_case("run_pbmc_r%d" % rank, pbmc, rank, 30 + rank, h1, **kw)
    # a run that stops on the convergence rule, with dn = 2 and a two-element gamma.a style hyper
    run_case("run_c1s2_r3_conv", mats["c1s2"], 3, 41, dict(aw=0.8, bw=1.0, ah=1.2, bh=1.0),
             Itmax=400, Tol=1e-4, n0=5, dn=2)
    # fixed hypers (hyper.update all FALSE)
    run_case("run_tiny_r2_fixed", tiny, 2, 42, h1, Itmax=25, Tol=1e-6, n0=10, dn=1,
             hyper_update_flags=(False,) * 4)


if __name__ == "__main__":
    main()
