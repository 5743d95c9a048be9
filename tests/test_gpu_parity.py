"""GPU parity tests (run on the B200 box): the CUDA engine, called through the C ABI, against
(1) the golden vectors produced by the reference's own src/vbnmf_update.cpp and (2) the CPU oracle on
fresh seeded inputs.  fp64 mode tolerance: 1e-9 relative on the bound and on every factor entry
(BASELINE.json north_star); cluster assignments bit-exact."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import (RUN_CASES, STEP_CASES, hyper_dict, load_counts, load_golden, relerr,
                      run_kwargs)

pytestmark = pytest.mark.gpu
FACT = ("lw", "lh", "ew", "eh", "dw", "dh")
TOL = 1e-9


@pytest.fixture(scope="module")
def Engine():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from ccfindr_b200.engine import Engine as E
    return E


@pytest.mark.parametrize("case", sorted(STEP_CASES))
def test_step_matches_reference_golden(Engine, case):
    g = load_golden(case)
    with Engine(load_counts(STEP_CASES[case])) as eng:
        eng.set_state(g["in_lw"], g["in_lh"], g["in_ew"], g["in_eh"])
        lkh = eng.step(hyper_dict(g["hyper"]), float(g["fudge"]))
        st = eng.get_state()
    assert relerr(lkh, g["lkh"]) < TOL
    for k in FACT:
        assert relerr(st[k], g["out_" + k]) < TOL, k


@pytest.mark.parametrize("case", sorted(RUN_CASES))
def test_run_matches_reference_golden(Engine, case):
    g = load_golden(case)
    kw = run_kwargs(g)
    flags = kw.pop("hyper_update_flags", (True,) * 4)
    with Engine(load_counts(RUN_CASES[case])) as eng:
        eng.set_state(g["w0"], g["h0"])
        res = eng.run(hyper_dict(g["hyper0"]), hyper_update=flags, **kw)
        st = eng.get_state()
        cid = eng.cluster_id()
    assert res["niter"] == int(g["niter"])
    assert res["stop_reason"] == int(g["stop_reason"])
    assert relerr(res["lkh_trace"], g["lkh_trace"]) < TOL
    assert relerr(res["hyper_trace"], g["hyper_trace"]) < TOL
    assert relerr(res["lml"], g["lml"]) < TOL
    for k in FACT:
        assert relerr(st[k], g[k]) < TOL, k
    assert np.array_equal(cid, g["cid"])  # bit-exact cluster assignments


def _random_problem(n, m, r, density, seed, integer=True):
    rng = np.random.default_rng(seed)
    X = sp.random(n, m, density=density, random_state=rng, format="csc",
                  data_rvs=lambda k: rng.integers(1, 30, size=k).astype(float) if integer
                  else rng.gamma(2.0, 1.5, size=k))
    from ccfindr_b200 import synth
    X = synth.fix_empty(X, seed)
    w0 = rng.gamma(1.0, 1.0, size=(n, r)) + 1e-3
    h0 = rng.gamma(1.0, 1.0, size=(r, m)) + 1e-3
    return X, w0, h0


@pytest.mark.parametrize("n,m,r,density,integer", [
    (300, 200, 1, 0.1, True),      # rank 1 (padded to 2)
    (257, 513, 7, 0.05, True),     # odd rank, ragged sizes
    (64, 2000, 13, 0.3, False),    # non-integer counts -> fp64 value storage
    (3000, 40, 20, 0.02, True),    # few cells, very short rows
    (500, 300, 33, 0.1, True),     # rank > 32 (padded to 40)
    (260, 330, 64, 0.2, True),     # the widest supported rank
    (900, 700, 24, 0.05, True),    # widest two-lane configuration that keeps 512 threads
])
def test_steps_match_oracle(Engine, n, m, r, density, integer):
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    X, w0, h0 = _random_problem(n, m, r, density, seed=n + m + r, integer=integer)
    hyper = dict(aw=0.7, bw=1.3, ah=1.1, bh=0.9)
    ref = od.vb_init_from(w0, h0)
    with Engine(X) as eng:
        eng.set_state(w0, h0)
        for it in range(4):
            ref = ob.sparse_vb_step(X, ref, hyper, od.EPS)
            lkh = eng.step(hyper, od.EPS)
            assert relerr(lkh, ref["lkh"]) < TOL, it
            assert relerr(eng.means(), ref["means"]) < TOL
        st = eng.get_state()
        cid = eng.cluster_id()
    for k in FACT:
        assert relerr(st[k], ref[k]) < TOL, k
    assert np.array_equal(cid, od.cluster_id(ref["eh"]))


def test_int64_colptr_and_single_nonzero_columns(Engine):
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    n, m, r = 50, 30, 3
    rows = np.arange(m) % n
    X = sp.csc_matrix((np.ones(m) * 3.0, (rows, np.arange(m))), shape=(n, m))
    X = (X + sp.csc_matrix((np.ones(n), (np.arange(n), np.arange(n) % m)), shape=(n, m))).tocsc()
    X.indptr = X.indptr.astype(np.int64)
    rng = np.random.default_rng(0)
    w0, h0 = rng.random((n, r)) + 0.1, rng.random((r, m)) + 0.1
    hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
    ref = ob.sparse_vb_step(X, od.vb_init_from(w0, h0), hyper, od.EPS)
    with Engine(X) as eng:
        eng.set_state(w0, h0)
        lkh = eng.step(hyper)
        st = eng.get_state()
    assert relerr(lkh, ref["lkh"]) < TOL
    for k in FACT:
        assert relerr(st[k], ref[k]) < TOL, k


def test_state_before_any_step_is_what_was_loaded(Engine):
    X, w0, h0 = _random_problem(40, 30, 4, 0.3, seed=3)
    with Engine(X) as eng:
        eng.set_state(w0, h0)
        st = eng.get_state()
    assert np.array_equal(st["lw"], w0) and np.array_equal(st["lh"], h0)
    assert np.array_equal(st["ew"], w0) and np.array_equal(st["eh"], h0)
    assert not st["dw"].any() and not st["dh"].any()  # vb_init: dw = dh = 0 (R/bayesian.R:161-162)


def test_sufficient_statistics_sum_rule_large(Engine):
    """Size-independent property: sum_k sw_ik = rowsum_i(X) and sum_k sh_kj = colsum_j(X), because
    sum_k lw_ik lh_kj / p_ij = 1.  With alpha = a + s and alpha = e^2/d from the returned moments."""
    n, m, r = 4000, 30000, 10
    X, w0, h0 = _random_problem(n, m, r, 0.02, seed=11)
    aw, ah = 0.9, 1.2
    with Engine(X) as eng:
        eng.set_state(w0, h0)
        eng.step(dict(aw=aw, bw=1.0, ah=ah, bh=1.0))
        st = eng.get_state(("ew", "eh", "dw", "dh"))
    alw = st["ew"] ** 2 / st["dw"]
    alh = st["eh"] ** 2 / st["dh"]
    rs = np.asarray(X.sum(axis=1)).ravel()
    cs = np.asarray(X.sum(axis=0)).ravel()
    assert relerr((alw - aw).sum(axis=1), rs) < 1e-9
    assert relerr((alh - ah).sum(axis=0), cs) < 1e-9


def test_runs_are_bitwise_reproducible(Engine):
    X, w0, h0 = _random_problem(800, 1200, 6, 0.05, seed=5)
    out = []
    for _ in range(2):
        with Engine(X) as eng:
            eng.set_state(w0, h0)
            res = eng.run(dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0), Itmax=15)
            out.append((res["lkh_trace"], eng.get_state()))
    assert np.array_equal(out[0][0], out[1][0])
    for k in FACT:
        assert np.array_equal(out[0][1][k], out[1][1][k])


@pytest.mark.parametrize("counts,r", [("tiny", 3), ("pbmc", 4)])
def test_ml_path_matches_oracle(Engine, counts, r):
    from ccfindr_b200 import synth
    from oracle import bindings as ob
    X = load_counts(counts)
    n, m = X.shape
    w0, h0 = synth.uniform_init(n, m, r, 4)
    ref = ob.sparse_ml_run(X, w0, h0, Itmax=40, Tol=1e-7)
    with Engine(X) as eng:
        res = eng.ml_run(w0, h0, Itmax=40, Tol=1e-7)
    assert res["niter"] == ref["niter"]
    assert relerr(res["lik_trace"], ref["lik_trace"]) < TOL
    assert relerr(res["w"], ref["w"]) < TOL and relerr(res["h"], ref["h"]) < TOL


@pytest.mark.parametrize("counts,r,ncnn", [("tiny", 2, 3), ("pbmc", 4, 5), ("c1s1", 3, 4)])
def test_ml_connectivity_criterion_matches_literal_reference_loop(Engine, counts, r, ncnn):
    """criterion='connectivity' (R/factorize.R:194-204): stop after ncnn.step iterations without a
    change of the cell co-clustering matrix.  The device loop takes sum(cnn != cnn0) from the r x r
    contingency table of consecutive labelings; the oracle forms the m(m-1)/2 vectors literally."""
    from ccfindr_b200 import synth
    from oracle import oracle_dense as od
    X = load_counts(counts)
    n, m = X.shape
    w0, h0 = synth.uniform_init(n, m, r, 11)
    w, h, lk0, it, trace, nch = od.ml_iterate(np.asarray(X.todense()), w0, h0, Itmax=60,
                                              criterion="connectivity", ncnn_step=ncnn)
    with Engine(X) as eng:
        res = eng.ml_run(w0, h0, Itmax=60, criterion="connectivity", ncnn_step=ncnn)
    assert res["niter"] == it
    assert np.array_equal(res["nchange_trace"], nch)            # exact pair counts
    assert relerr(res["lik_trace"], trace) < TOL
    assert relerr(res["w"], w) < TOL and relerr(res["h"], h) < TOL


@pytest.mark.parametrize("r", [7, 10, 13])
def test_opt_in_four_unit_split_layout_matches_oracle(Engine, monkeypatch, r):
    """VBNMF_SPLIT4=1: ranks 8..14 through the 4-unit split layout (parity classes side by side,
    rotated gathers) -- opt-in because it measured slower, but it must stay correct."""
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    monkeypatch.setenv("VBNMF_SPLIT4", "1")
    X, w0, h0 = _random_problem(700, 900, r, 0.06, seed=50 + r)
    hyper = dict(aw=0.9, bw=1.1, ah=1.2, bh=0.8)
    ref = od.vb_init_from(w0, h0)
    with Engine(X) as eng:
        eng.set_state(w0, h0)
        for it in range(3):
            ref = ob.sparse_vb_step(X, ref, hyper, od.EPS)
            assert relerr(eng.step(hyper, od.EPS), ref["lkh"]) < TOL, it
        st = eng.get_state()
    for k in FACT:
        assert relerr(st[k], ref[k]) < TOL, k


@pytest.mark.parametrize("r", [15, 16, 17, 18, 19, 20])
@pytest.mark.parametrize("variant", ["default", "eight_lane_groups", "owner_order"])
def test_split_layout_variants_match_oracle(Engine, monkeypatch, r, variant):
    """Ranks 15..20 through the split layout with several slabs per side (small tile forced) and
    small windows of the length-sorted segment order: 4-lane groups with the four-class schedule
    (default), 8-lane groups (VBNMF_NO_G4: the four classes of two lanes at ranks 19, 20), and the
    segments left in owner order (VBNMF_SEG_WINDOW=0)."""
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    monkeypatch.setenv("VBNMF_TILE_ROWS", "160")
    monkeypatch.setenv("VBNMF_SEG_WINDOW", "0" if variant == "owner_order" else "64")
    if variant == "eight_lane_groups":
        monkeypatch.setenv("VBNMF_NO_G4", "1")
    X, w0, h0 = _random_problem(700, 900, r, 0.08, seed=80 + r)
    hyper = dict(aw=0.9, bw=1.1, ah=1.2, bh=0.8)
    ref = od.vb_init_from(w0, h0)
    with Engine(X) as eng:
        eng.set_state(w0, h0)
        info = eng.layout_info()
        assert info["tile_rows"] == 160 and info["gene_slabs"] == 5 and info["cell_slabs"] == 6
        assert info["nonzeros_per_group_step"] == (8 if variant == "eight_lane_groups" else 4)
        for it in range(3):
            ref = ob.sparse_vb_step(X, ref, hyper, od.EPS)
            assert relerr(eng.step(hyper, od.EPS), ref["lkh"]) < TOL, it
        st = eng.get_state()
        cid = eng.cluster_id()
    for k in FACT:
        assert relerr(st[k], ref[k]) < TOL, k
    assert np.array_equal(cid, od.cluster_id(ref["eh"]))


def test_errors_are_reported_not_thrown(Engine):
    from ccfindr_b200 import _lib
    X, w0, h0 = _random_problem(40, 30, 4, 0.3, seed=3)
    with Engine(X) as eng:
        with pytest.raises(_lib.VbnmfError):
            eng.step(dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0))  # no state loaded
        with pytest.raises(ValueError):
            eng.set_state(w0[:-1], h0)


# ---- fp32-storage / fp64-accumulate mode (BASELINE.json north_star: within 1e-4) -------------------
TOL32 = 1e-4


def _relerr_floor(a, b, floor):
    """relative error with an absolute floor: entries far below the scale of the matrix carry the
    fp32 rounding of the panels relative to the dominant terms they were summed with"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


@pytest.mark.parametrize("case", ["run_c1s1_r3", "run_pbmc_r5", "run_c1s2_r3_conv", "run_c1s1_r2"])
def test_fp32_storage_mode_within_1e4(Engine, case):
    g = load_golden(case)
    kw = run_kwargs(g)
    flags = kw.pop("hyper_update_flags", (True,) * 4)
    kw["Tol"] = 0.0               # fixed iteration count: compare like with like
    kw["Itmax"] = int(g["niter"])
    with Engine(load_counts(RUN_CASES[case])) as eng:
        eng.set_precision(1)      # VBNMF_FP32_STORAGE
        eng.set_state(g["w0"], g["h0"])
        res = eng.run(hyper_dict(g["hyper0"]), hyper_update=flags, **kw)
        st = eng.get_state()
        cid = eng.cluster_id()
    assert relerr(res["lkh_trace"], g["lkh_trace"]) < TOL32
    assert relerr(res["hyper_trace"], g["hyper_trace"]) < TOL32
    for k in ("ew", "eh", "lw", "lh"):
        scale = float(np.abs(g[k]).max())
        assert _relerr_floor(st[k], g[k], 1e-3 * scale) < TOL32, k
    # cluster labels may only differ where the top two coefficients are within fp32 noise
    diff = np.flatnonzero(cid != g["cid"])
    for j in diff:
        col = np.sort(g["eh"][:, j])[::-1]
        assert (col[0] - col[1]) / col[0] < 1e-3


def test_fp32_mode_step_matches_oracle_random(Engine):
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    X, w0, h0 = _random_problem(700, 900, 20, 0.05, seed=21)
    hyper = dict(aw=0.7, bw=1.3, ah=1.1, bh=0.9)
    ref = od.vb_init_from(w0, h0)
    with Engine(X) as eng:
        eng.set_precision(1)
        eng.set_state(w0, h0)
        for it in range(3):
            ref = ob.sparse_vb_step(X, ref, hyper, od.EPS)
            lkh = eng.step(hyper, od.EPS)
            assert relerr(lkh, ref["lkh"]) < TOL32, it
        st = eng.get_state()
    for k in ("ew", "eh"):
        assert _relerr_floor(st[k], ref[k], 1e-3 * float(np.abs(ref[k]).max())) < TOL32, k


# ---- edge cases ------------------------------------------------------------------------------------
def _check_steps(Engine, X, w0, h0, hyper, nsteps=3, tol=TOL):
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    ref = od.vb_init_from(w0, h0)
    with Engine(X) as eng:
        eng.set_state(w0, h0)
        for it in range(nsteps):
            ref = ob.sparse_vb_step(X, ref, hyper, od.EPS)
            assert relerr(eng.step(hyper, od.EPS), ref["lkh"]) < tol, it
        st = eng.get_state()
    for k in FACT:
        assert relerr(st[k], ref[k]) < tol, k


def test_empty_rows_and_columns_at_the_abi_level(Engine, monkeypatch):
    """The C ABI itself refuses empty rows / columns with VBNMF_ERR_EMPTY, as vb_factorize does
    (R/bayesian.R:244-247: rowSums / colSums == 0, so an explicit zero does not count), and
    negative or non-finite counts.  A SHARD of cells may well contain genes without counts, so the
    kernels must still handle them: VBNMF_ALLOW_EMPTY=1 lifts the gene test."""
    from ccfindr_b200._lib import VbnmfError
    rng = np.random.default_rng(2)
    X = sp.random(60, 45, density=0.2, random_state=rng, format="lil",
                  data_rvs=lambda k: rng.integers(1, 9, size=k).astype(float))
    X[7, :] = 0; X[59, :] = 0
    Xr = sp.csc_matrix(X); Xr.eliminate_zeros()
    Xr = synth_fix_cols(Xr)
    w0, h0 = rng.random((60, 3)) + 0.1, rng.random((3, 45)) + 0.1
    hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
    with Engine(Xr) as eng:                                    # empty genes: refused at set_state
        with pytest.raises(VbnmfError, match="empty rows") as ei:
            eng.set_state(w0, h0)
        assert ei.value.code == 6
    from ccfindr_b200 import synth
    Xc = synth.fix_empty(Xr, 3).tolil(); Xc[:, 11] = 0; Xc = sp.csc_matrix(Xc); Xc.eliminate_zeros()
    assert (np.asarray(Xc.sum(axis=1)).ravel() > 0).all()
    with Engine(Xc) as eng:                                    # empty cells: likewise
        with pytest.raises(VbnmfError, match="empty columns") as ei:
            eng.set_state(w0, h0)
        assert ei.value.code == 6
    Xz = Xr.copy().tolil(); Xz[7, 3] = 1.0; Xz = sp.csc_matrix(Xz)
    Xz.data[Xz.indices == 7] = 0.0                             # explicit zero only: still empty
    with Engine(Xz) as eng:
        with pytest.raises(VbnmfError, match="empty rows"):
            eng.set_state(w0, h0)
    Xn = Xr.copy(); Xn.data[5] = -1.0
    with pytest.raises(VbnmfError, match="non-negative"):
        Engine(Xn)
    Xn.data[5] = np.nan
    with pytest.raises(VbnmfError, match="non-negative"):
        Engine(Xn)
    monkeypatch.setenv("VBNMF_ALLOW_EMPTY", "1")
    _check_steps(Engine, Xr, w0, h0, hyper)


def synth_fix_cols(X):
    """one count into every empty column (rows are left alone)"""
    X = X.tolil()
    for j in np.flatnonzero(np.asarray(X.sum(axis=0)).ravel() == 0):
        X[(3 * j) % X.shape[0] if (3 * j) % X.shape[0] not in (7, 59) else 1, j] = 1.0
    return sp.csc_matrix(X)


def test_two_engines_with_different_tiles_alive_together(Engine):
    """The dynamic shared-memory opt-in is per kernel and device, i.e. shared by all handles: a
    second handle with a smaller tile at the same padded rank must not break the first."""
    from ccfindr_b200 import synth
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    rng = np.random.default_rng(12)
    hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
    big = synth.fix_empty(sp.random(6000, 5000, density=0.05, random_state=rng, format="csc",
                                    data_rvs=lambda k: rng.integers(1, 9, size=k).astype(float)), 1)
    small = synth.fix_empty(sp.random(70, 50, density=0.3, random_state=rng, format="csc",
                                      data_rvs=lambda k: rng.integers(1, 9, size=k).astype(float)), 1)
    r = 4
    wb, hb = rng.random((6000, r)) + 0.1, rng.random((r, 5000)) + 0.1
    ws, hs = rng.random((70, r)) + 0.1, rng.random((r, 50)) + 0.1
    with Engine(big) as e1:
        e1.set_state(wb, hb)
        l1 = e1.step(hyper)
        with Engine(small) as e2:
            e2.set_state(ws, hs)
            assert e2.layout_info()["tile_rows"] < e1.layout_info()["tile_rows"]
            ls = e2.step(hyper)
            l2 = e1.step(hyper)                                # first handle, after the second's opt-in
            ref = ob.sparse_vb_step(small, od.vb_init_from(ws, hs), hyper, od.EPS)
            assert relerr(ls, ref["lkh"]) < 1e-9
    refb = ob.sparse_vb_step(big, od.vb_init_from(wb, hb), hyper, od.EPS)
    assert relerr(l1, refb["lkh"]) < 1e-9
    refb2 = ob.sparse_vb_step(big, refb, hyper, od.EPS)
    assert relerr(l2, refb2["lkh"]) < 1e-9


@pytest.mark.parametrize("n,m,r", [(3, 2, 2), (2, 5, 1), (1, 1, 1), (130, 9, 9)])
def test_tiny_and_degenerate_shapes(Engine, n, m, r):
    rng = np.random.default_rng(n * 100 + m)
    X = sp.csc_matrix(rng.integers(1, 6, size=(n, m)).astype(float))
    w0, h0 = rng.random((n, r)) + 0.1, rng.random((r, m)) + 0.1
    _check_steps(Engine, X, w0, h0, dict(aw=0.5, bw=2.0, ah=2.0, bh=0.5))


def test_large_and_non_fp32_counts(Engine):
    """counts >= 64, >= 2^24 (not fp32-representable -> fp64 count storage) and fractional"""
    rng = np.random.default_rng(9)
    n, m, r = 80, 70, 4
    X = sp.random(n, m, density=0.25, random_state=rng, format="csc",
                  data_rvs=lambda k: rng.integers(1, 2000, size=k).astype(float))
    from ccfindr_b200 import synth
    X = synth.fix_empty(X, 1)
    w0, h0 = rng.random((n, r)) + 0.1, rng.random((r, m)) + 0.1
    hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
    _check_steps(Engine, X, w0, h0, hyper)
    Y = X.copy(); Y.data[::7] = 2.0 ** 24 + 1.0; Y.data[3::11] += 0.37
    _check_steps(Engine, Y, w0, h0, hyper)


def test_nan_bound_stops_the_loop(Engine):
    """fudge = 0 and a gene whose lw row is zero: p = 0 at its nonzeros -> NaN bound ->
    `if(is.na(wh$lkh)) break` (R/bayesian.R:345): one iteration, lml stays at its initial 0."""
    rng = np.random.default_rng(4)
    X = sp.csc_matrix(rng.integers(1, 5, size=(20, 15)).astype(float))
    w0, h0 = rng.random((20, 2)) + 0.1, rng.random((2, 15)) + 0.1
    w0[3, :] = 0.0
    with Engine(X) as eng:
        eng.set_state(w0, h0)
        res = eng.run(dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0), Itmax=50, fudge=0.0)
    assert res["stop_reason"] == 2 and res["niter"] == 1 and res["lml"] == 0.0
    assert np.isnan(res["lkh_trace"][0])


def test_itmax_one_and_rank_switching_on_one_handle(Engine):
    from oracle import bindings as ob
    X = load_counts("tiny")
    n, m = X.shape
    rng = np.random.default_rng(8)
    hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
    with Engine(X) as eng:
        for r in (2, 5, 3, 2):                      # the rank loop of vb_iterate reuses one handle
            w0, h0 = rng.random((n, r)) + 0.1, rng.random((r, m)) + 0.1
            eng.set_state(w0, h0)
            res = eng.run(hyper, Itmax=1)
            ref = ob.sparse_vb_run(X, w0, h0, hyper, Itmax=1)
            assert res["niter"] == 1 and res["stop_reason"] == 0
            assert relerr(res["lml"], ref["lml"]) < TOL
            assert relerr(eng.get_state(("eh",))["eh"], ref["eh"]) < TOL


def test_ml_path_in_fp32_storage_mode(Engine):
    from ccfindr_b200 import synth
    from oracle import bindings as ob
    X = load_counts("pbmc")
    n, m = X.shape
    w0, h0 = synth.uniform_init(n, m, 5, 4)
    ref = ob.sparse_ml_run(X, w0, h0, Itmax=20, Tol=0.0)
    with Engine(X) as eng:
        eng.set_precision(1)
        res = eng.ml_run(w0, h0, Itmax=20, Tol=0.0)
    assert res["niter"] == 20
    assert relerr(res["lik_trace"], ref["lik_trace"]) < TOL32
    assert _relerr_floor(res["h"], ref["h"], 1e-3 * float(np.abs(ref["h"]).max())) < TOL32


def test_external_stream(Engine):
    """vbnmf_set_stream: the engine runs on a caller-provided CUDA stream (a torch stream here)"""
    import torch
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    X, w0, h0 = _random_problem(300, 200, 5, 0.1, seed=12)
    hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
    ref = ob.sparse_vb_step(X, od.vb_init_from(w0, h0), hyper, od.EPS)
    st = torch.cuda.Stream()
    with Engine(X) as eng:
        eng.set_stream(st.cuda_stream)
        eng.set_state(w0, h0)
        assert relerr(eng.step(hyper), ref["lkh"]) < TOL
        st.synchronize()


# ---- storage formats of the nonzeros -----------------------------------
def _check_format_steps(Engine, X, w0, h0, nsteps=3, expect_format=None):
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    hyper = dict(aw=0.9, bw=1.1, ah=1.2, bh=0.8)
    ref = od.vb_init_from(w0, h0)
    with Engine(X) as eng:
        eng.set_state(w0, h0)
        if expect_format is not None:
            assert eng.layout_info()["format"] == expect_format
        for it in range(nsteps):
            ref = ob.sparse_vb_step(X, ref, hyper, od.EPS)
            lkh = eng.step(hyper, od.EPS)
            assert relerr(lkh, ref["lkh"]) < TOL, it
        st = eng.get_state()
        info = eng.layout_info()
    for k in FACT:
        assert relerr(st[k], ref[k]) < TOL, k
    return info


def test_integer_counts_use_the_packed16_layout(Engine):
    X, w0, h0 = _random_problem(700, 900, 10, 0.08, seed=11)
    info = _check_format_steps(Engine, X, w0, h0, expect_format="p16")
    # segments are whole chunks of 4 steps x 8 nonzeros: stored entries >= nonzeros, multiple of 32
    assert info["entries_cols"] >= X.nnz and info["entries_cols"] % 32 == 0
    assert info["entries_rows"] >= X.nnz and info["entries_rows"] % 32 == 0


def test_counts_beyond_16_bits_fall_back_to_float_entries(Engine):
    X, w0, h0 = _random_problem(300, 250, 6, 0.1, seed=12)
    X.data[::7] = 70000.0           # exact in fp32, not in 16 bits
    X.data[1::11] = 65535.0         # the largest packed count
    _check_format_steps(Engine, X, w0, h0, expect_format="f32")
    Y = X.copy()
    Y.data[Y.data > 65535.0] = 65535.0
    _check_format_steps(Engine, Y, w0, h0, expect_format="p16")


def test_packed16_large_counts_take_the_log_slow_path(Engine):
    # counts >= 8 bypass the bit-sliced log-product of the fp64 cell-owner pass
    X, w0, h0 = _random_problem(400, 350, 10, 0.15, seed=13)
    rng = np.random.default_rng(5)
    X.data[:] = rng.choice([1, 2, 3, 7, 8, 9, 15, 16, 255, 4096, 65535], size=X.nnz).astype(float)
    _check_format_steps(Engine, X, w0, h0, expect_format="p16")


@pytest.mark.parametrize("r", [10, 20])   # one / two lanes per nonzero in fp64
def test_packed16_skewed_tile_row_residues(Engine, r):
    """Nonzeros only in genes/cells whose index is a multiple of 8 (plus a few others): after the
    count-sorted renumbering the residue classes of a segment are as unbalanced as they get, so the
    schedule needs pair steps and hole steps everywhere."""
    n, m = 640, 520
    rng = np.random.default_rng(7)
    D = np.zeros((n, m))
    D[::8, :] = rng.integers(0, 4, size=(n // 8, m))
    cols = np.arange(0, m, 8)
    D[:, cols] += rng.integers(0, 3, size=(n, len(cols)))
    D[rng.integers(0, n, 300), rng.integers(0, m, 300)] += 1
    from ccfindr_b200 import synth
    X = synth.fix_empty(sp.csc_matrix(D), 1)
    w0 = rng.gamma(1.0, 1.0, size=(n, r)) + 1e-3
    h0 = rng.gamma(1.0, 1.0, size=(r, m)) + 1e-3
    _check_format_steps(Engine, X, w0, h0, expect_format="p16")


def test_float_entry_layout_still_matches(Engine, monkeypatch):
    """VBNMF_NO_P16=1 keeps the 8-byte {row, float} entries for integer counts (the path that
    non-integer but fp32-exact counts take)."""
    monkeypatch.setenv("VBNMF_NO_P16", "1")
    X, w0, h0 = _random_problem(500, 400, 10, 0.1, seed=14)
    _check_format_steps(Engine, X, w0, h0, expect_format="f32")
    X.data[:] = np.round(X.data * 0.5, 1) + 0.5      # halves: exact in fp32, not integers
    monkeypatch.delenv("VBNMF_NO_P16")
    _check_format_steps(Engine, X, w0, h0, expect_format="f32")


def test_packed16_tiny_matrix_holes_point_at_real_rows(Engine):
    # fewer genes than residue classes: hole words must not read padding rows of the tile
    X = sp.csc_matrix(np.array([[1.0, 0, 2, 0, 1], [0, 3, 0, 1, 0], [2, 0, 0, 0, 9]]))
    rng = np.random.default_rng(1)
    w0, h0 = rng.random((3, 2)) + 0.1, rng.random((2, 5)) + 0.1
    _check_format_steps(Engine, X, w0, h0, expect_format="p16")
