"""CPU tests of the checkers themselves (no GPU): the C restatement (oracle/oracle_sparse.c) and the
NumPy restatement (oracle/oracle_dense.py) against the golden vectors that the reference's own
src/vbnmf_update.cpp produced (tests/golden/make_golden.py), and against that binary directly
when it is present (oracle/_ref)."""
import mpmath
import numpy as np
import pytest

from conftest import (RUN_CASES, STEP_CASES, hyper_dict, load_counts, load_golden, relerr,
                      run_kwargs)
from oracle import bindings as ob
from oracle import oracle_dense as od

FACT = ("lw", "lh", "ew", "eh", "dw", "dh")


@pytest.mark.parametrize("case", sorted(STEP_CASES))
def test_step_matches_reference_golden(case):
    g = load_golden(case)
    X = load_counts(STEP_CASES[case])
    wh = {k: g["in_" + k] for k in ("lw", "lh", "ew", "eh")}
    hyper = hyper_dict(g["hyper"])
    fud = float(g["fudge"])
    for name, res in (("dense", od.vbnmf_update(np.asarray(X.todense()), wh, hyper, fud)),
                      ("twin", od.vbnmf_updateR(np.asarray(X.todense()), wh, wh["lw"].shape[1],
                                                hyper, fud)),
                      ("sparse", ob.sparse_vb_step(X, wh, hyper, fud))):
        assert relerr(res["lkh"], g["lkh"]) < 1e-11, name
        for k in FACT:
            assert relerr(res[k], g["out_" + k]) < 1e-12, (name, k)


@pytest.mark.parametrize("case", sorted(RUN_CASES))
def test_sparse_run_matches_reference_golden(case):
    g = load_golden(case)
    X = load_counts(RUN_CASES[case])
    res = ob.sparse_vb_run(X, g["w0"], g["h0"], hyper_dict(g["hyper0"]), **run_kwargs(g))
    assert res["niter"] == int(g["niter"])
    assert res["stop_reason"] == int(g["stop_reason"])
    assert relerr(res["lkh_trace"], g["lkh_trace"]) < 1e-10
    assert relerr(res["hyper_trace"], g["hyper_trace"]) < 1e-10
    assert relerr(res["lml"], g["lml"]) < 1e-10
    for k in FACT:
        assert relerr(res[k], g[k]) < 1e-9, k
    assert np.array_equal(od.cluster_id(res["eh"]), g["cid"])


@pytest.mark.parametrize("case", ["run_tiny_r2_fixed", "run_c1s2_r3_conv"])
def test_dense_run_matches_reference_golden(case):
    g = load_golden(case)
    X = np.asarray(load_counts(RUN_CASES[case]).todense())
    wh, hyper, lk0, it, trace, htrace, reason = od.vb_run_one_rank(
        X, g["w0"], g["h0"], hyper_dict(g["hyper0"]), **run_kwargs(g))
    assert it == int(g["niter"]) and reason == int(g["stop_reason"])
    assert relerr(trace, g["lkh_trace"]) < 1e-10
    assert relerr(htrace, g["hyper_trace"]) < 1e-10
    for k in FACT:
        assert relerr(wh[k], g[k]) < 1e-9, k


def test_against_compiled_reference_live():
    """When oracle/_ref exists (the build container, or shipped to the GPU box), run the
    reference binary itself on a fresh input and compare both restatements."""
    if ob.ref_lib() is None:
        pytest.skip("oracle/_ref/libccfindr_ref.so not available")
    from ccfindr_b200 import synth
    import scipy.sparse as sp
    x = synth.simulate_whx(nrow=120, ncol=70, rank=4, seed=5)["x"]
    n, m = x.shape
    hyper = dict(aw=0.7, bw=1.3, ah=1.1, bh=0.9)
    w0, h0 = synth.random_init(n, m, 4, hyper, 9)
    a = b = c = od.vb_init_from(w0, h0)
    for _ in range(4):
        a = ob.ref_vbnmf_update(x, a, hyper, od.EPS)
        b = od.vbnmf_update(x, b, hyper, od.EPS)
        c = ob.sparse_vb_step(sp.csc_matrix(x), c, hyper, od.EPS)
        assert relerr(b["lkh"], a["lkh"]) < 1e-11 and relerr(c["lkh"], a["lkh"]) < 1e-11
        for k in FACT:
            assert relerr(b[k], a[k]) < 1e-12 and relerr(c[k], a[k]) < 1e-12


def test_special_functions_against_mpmath():
    lib = ob.sparse_lib()
    mpmath.mp.dps = 50
    xs = np.concatenate([np.logspace(-3, 3, 61), [0.5, 1.0, 1.4616321449683623, 2.0, 9.99, 10.0]])
    for x in xs:
        d = float(mpmath.digamma(mpmath.mpf(float(x))))
        t = float(mpmath.polygamma(1, mpmath.mpf(float(x))))
        assert abs(lib.osp_digamma(float(x)) - d) <= 3e-16 * abs(d) + 5e-16
        assert abs(lib.osp_trigamma(float(x)) - t) <= 1e-15 * abs(t)


def test_hyper_update_c_matches_python():
    import ctypes as C
    lib = ob.sparse_lib()
    rng = np.random.default_rng(3)
    for flags in ((1, 1, 1, 1), (1, 0, 0, 1), (0, 1, 1, 0), (0, 0, 0, 0), (0, 1, 0, 1)):
        lw = rng.gamma(0.3, 1.0, size=(30, 3)) + 1e-3
        lh = rng.gamma(0.5, 1.0, size=(3, 20)) + 1e-3
        ew, eh = lw * 1.3, lh * 0.8
        hyper = dict(aw=0.9, bw=1.2, ah=1.1, bh=0.7)
        ref = od.hyper_update(flags, dict(lw=lw, lh=lh, ew=ew, eh=eh), hyper, Niter=100, Tol=1e-3)
        hy = np.array([hyper[k] for k in ("aw", "bw", "ah", "bh")])
        fl = np.array(flags, dtype=np.int32)
        rc = lib.osp_hyper_update(fl.ctypes.data_as(C.POINTER(C.c_int)),
                                  C.c_double(np.mean(np.log(lw))), C.c_double(np.mean(np.log(lh))),
                                  C.c_double(np.mean(ew)), C.c_double(np.mean(eh)),
                                  hy.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(100),
                                  C.c_double(1e-3))
        assert rc == 0
        assert relerr(hy, [ref[k] for k in ("aw", "bw", "ah", "bh")]) < 1e-12


def test_ml_path_sparse_matches_dense():
    from ccfindr_b200 import synth
    X = load_counts("tiny")
    n, m = X.shape
    w0, h0 = synth.uniform_init(n, m, 3, 4)
    w, h, lk, it, trace = od.ml_iterate(np.asarray(X.todense()), w0, h0, Itmax=60, Tol=1e-6)
    res = ob.sparse_ml_run(X, w0, h0, Itmax=60, Tol=1e-6)
    assert res["niter"] == it
    assert relerr(res["lik_trace"], trace) < 1e-11
    assert relerr(res["w"], w) < 1e-9 and relerr(res["h"], h) < 1e-9
