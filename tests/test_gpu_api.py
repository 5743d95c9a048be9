"""GPU tests of the ccfindR-style front end against the oracle's restatement of the R driver
(vb_iterate + best-run selection, R/bayesian.R:265-390; factorize loop, R/factorize.R:187-223)."""
import numpy as np
import pytest

from conftest import load_counts, relerr

pytestmark = pytest.mark.gpu


def test_vb_factorize_matches_reference_driver():
    from ccfindr_b200 import api, synth
    from oracle import bindings as ob
    from oracle import oracle_dense as od
    X = load_counts("c1s3")
    n, m = X.shape
    ranks, nrun = [2, 3, 4], 2
    kw = dict(Itmax=60, Tol=1e-5)
    hyper0 = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
    inits = {(i, r): synth.random_init(n, m, r, hyper0, 100 * i + r)
             for i in range(1, nrun + 1) for r in ranks}
    s = api.vb_factorize(api.scNMFSet(X), ranks=ranks, nrun=nrun, verbose=0, inits=inits, **kw)
    # reference driver: per run, per rank the loop of vb_iterate; then best run per rank
    runs = []
    for i in range(1, nrun + 1):
        out = dict(rdat=[], wdat=[], hdat=[], dwdat=[], dhdat=[], hyperp=[], nunif=[])
        for r in ranks:
            res = ob.sparse_vb_run(X, *inits[(i, r)], hyper0, **kw)
            out["rdat"].append(res["lml"]); out["wdat"].append(res["ew"]); out["hdat"].append(res["eh"])
            out["dwdat"].append(np.sqrt(res["dw"])); out["dhdat"].append(np.sqrt(res["dh"]))
            out["hyperp"].append(res["hyper"]); out["nunif"].append(0)
        runs.append(out)
    ref = od.vb_select(runs, ranks)
    assert list(s.ranks) == ref["ranks"]
    assert relerr(s.measure["lml"], ref["lml"]) < 1e-9
    for key in ("aw", "bw", "ah", "bh"):
        assert relerr(s.measure[key], ref[key]) < 1e-9
    for k in range(len(ranks)):
        assert relerr(s.basis[k], ref["basis"][k]) < 1e-9
        assert relerr(s.coeff[k], ref["coeff"][k]) < 1e-9
        assert relerr(s.dbasis[k], ref["dbasis"][k]) < 1e-9
        assert relerr(s.dcoeff[k], ref["dcoeff"][k]) < 1e-9
        assert np.array_equal(api.cluster_id(s, ranks[k]), od.cluster_id(ref["coeff"][k]))
    out = api.optimal_rank(s)
    assert out["type"] in (1, 2) and out["ropt"] in ranks


def test_factorize_ml_matches_reference_loop():
    from ccfindr_b200 import api, synth
    from oracle import bindings as ob
    X = load_counts("tiny")
    n, m = X.shape
    s = api.factorize(api.scNMFSet(X), ranks=[2, 3], nrun=3, verbose=0, Itmax=200, Tol=1e-6, seed=2)
    for k, r in enumerate([2, 3]):
        best = -np.inf
        for irun in range(1, 4):
            w0, h0 = synth.uniform_init(n, m, r, 2 * 100003 + 1000 * r + 37 + irun)
            res = ob.sparse_ml_run(X, w0, h0, Itmax=200, Tol=1e-6)
            if irun == 1 or res["lik"] > best:
                best, wb, hb = res["lik"], res["w"], res["h"]
        assert relerr(s.measure["likelihood"][k], best) < 1e-9
        assert relerr(s.basis[k], wb) < 1e-8 and relerr(s.coeff[k], hb) < 1e-8
    assert np.isfinite(s.measure["dispersion"]).all()


def test_device_random_init_matches_its_cpu_restatement():
    import scipy.sparse as sp
    from ccfindr_b200 import synth
    from ccfindr_b200.engine import Engine
    rng = np.random.default_rng(3)
    X = synth.fix_empty(sp.random(90, 70, density=0.2, random_state=rng, format="csc",
                                  data_rvs=lambda k: rng.integers(1, 9, size=k).astype(float)), 1)
    for hyper, r in ((dict(aw=0.4, bw=1.5, ah=2.5, bh=0.8), 5), (dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0), 3)):
        w, h = synth.device_random_init_reference(90, 70, r, hyper, seed=11)
        with Engine(X) as eng:
            eng.init_random(r, hyper, seed=11)
            st = eng.get_state()
            assert np.allclose(st["lw"], w, rtol=1e-12, atol=0) and np.allclose(st["lh"], h, rtol=1e-12, atol=0)
            assert np.array_equal(st["ew"], st["lw"]) and np.array_equal(st["eh"], st["lh"])
            assert not st["dw"].any() and not st["dh"].any()
            a = eng.run(hyper, Itmax=15)
            eng.set_state(st["lw"], st["lh"])
            b = eng.run(hyper, Itmax=15)
        # same state either way (rowSums(eh) is summed in a different order: not bitwise)
        assert a["niter"] == b["niter"] and relerr(a["lkh_trace"], b["lkh_trace"]) < 1e-12
    # a shard of cells draws the same columns (cell_offset = first global cell of the shard)
    with Engine(X[:, 30:].tocsc()) as eng:
        eng.init_random(3, hyper, seed=11, cell_offset=30)
        hs = eng.get_state(("lh",))["lh"]
    assert np.allclose(hs, h[:, 30:], rtol=1e-12, atol=0)


def test_vb_factorize_with_device_init():
    from ccfindr_b200 import api, synth
    import scipy.sparse as sp
    x = sp.csc_matrix(synth.simulate_whx(nrow=200, ncol=80, rank=3, seed=4)["x"])
    a = api.vb_factorize(api.scNMFSet(count=x), ranks=[2, 3], nrun=2, verbose=0, Itmax=40,
                         device_init=True, connectivity=False)
    b = api.vb_factorize(api.scNMFSet(count=x), ranks=[2, 3], nrun=2, verbose=0, Itmax=40,
                         device_init=True, connectivity=False)
    assert a.ranks == [2, 3] and np.isfinite(a.measure["lml"]).all()
    assert np.array_equal(a.measure["lml"], b.measure["lml"])      # same seeds, same draws
    assert np.array_equal(a.basis[1], b.basis[1])


def test_graph_replay_and_plain_launches_give_the_same_run(monkeypatch):
    """Small problems replay batches of iterations as a CUDA graph; VBNMF_NO_GRAPH=1 launches the
    kernels one by one.  Same kernels, same arguments: bitwise identical traces."""
    import scipy.sparse as sp
    from ccfindr_b200 import synth
    from ccfindr_b200.engine import Engine
    x = sp.csc_matrix(synth.simulate_whx(nrow=300, ncol=120, rank=3, seed=9)["x"])
    hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
    w0, h0 = synth.random_init(*x.shape, 4, hyper, seed=5)
    out = []
    for no_graph in (False, True):
        if no_graph:
            monkeypatch.setenv("VBNMF_NO_GRAPH", "1")
        with Engine(x) as eng:
            eng.set_state(w0, h0)
            out.append(eng.run(hyper, Itmax=70, Tol=1e-7))     # 70 is not a multiple of the batch
    assert out[0]["niter"] == out[1]["niter"] and out[0]["stop_reason"] == out[1]["stop_reason"]
    assert np.array_equal(out[0]["lkh_trace"], out[1]["lkh_trace"])
    assert np.array_equal(out[0]["hyper_trace"], out[1]["hyper_trace"])


@pytest.mark.parametrize("counts,rank", [("pbmc", 5), ("c1s1", 3), ("pbmc", 2)])
def test_device_svd2_initializer_matches_the_host_one(counts, rank):
    """vb_init(initializer='svd2') (R/bayesian.R:150-159) computed on the GPU (randomized truncated
    SVD of the resident CSC matrix) against the host restatement on scipy's ARPACK SVD: the
    singular subspaces agree, abs() removes the sign freedom."""
    from ccfindr_b200 import api
    from ccfindr_b200.engine import Engine
    X = load_counts(counts)
    n, m = X.shape
    hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=0.7)
    w_ref, h_ref = api.vb_init(n, m, X, rank, hyper, "svd2", seed=1)
    with Engine(X) as eng:
        eng.init_svd2(rank, hyper, seed=5)
        st = eng.get_state(("lw", "lh", "ew", "eh", "dw"))
    for got, ref in ((st["lw"], w_ref), (st["lh"], h_ref), (st["ew"], w_ref), (st["eh"], h_ref)):
        assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-6
    assert abs(st["lh"].mean() - hyper["bh"]) < 1e-12 * hyper["bh"] * 10   # scale = bh / mean(h)
    assert not st["dw"].any()                                   # dw = dh = 0 at initialisation (:161-162)


def test_vb_factorize_with_device_svd2_equals_host_svd2():
    from ccfindr_b200 import api
    X = load_counts("pbmc")
    kw = dict(ranks=[3, 4], nrun=1, verbose=0, Itmax=30, initializer="svd2")
    a = api.vb_factorize(api.scNMFSet(X), **kw)
    b = api.vb_factorize(api.scNMFSet(X), device_init=True, **kw)
    assert list(a.ranks) == list(b.ranks)
    assert np.max(np.abs(a.measure["lml"] - b.measure["lml"]) / np.abs(a.measure["lml"])) < 1e-6


def test_matrix_market_parsed_on_the_device_equals_host_reader(tmp_path):
    """read_10x's Matrix::readMM + as(., 'dgCMatrix') (R/utils.R:34) done on the GPU: the bundled
    PBMC matrix.mtx (integer field), a real-valued file with exponents and unsorted entries, and
    the error paths."""
    import scipy.io
    import scipy.sparse as sp
    from ccfindr_b200 import _lib
    from ccfindr_b200.engine import Engine
    rng = np.random.default_rng(0)
    # integer counts, as 10x writes them
    X = load_counts("pbmc")
    p = tmp_path / "matrix.mtx"
    scipy.io.mmwrite(str(p), sp.coo_matrix(X), field="integer")
    with Engine.from_mtx(str(p)) as eng:
        got = eng.csc()
    assert got.shape == X.shape and (got != X).nnz == 0
    assert np.array_equal(got.indptr, X.indptr) and np.array_equal(got.indices, X.indices)
    # real values (normalized counts), shuffled order, exponents
    Y = sp.random(300, 200, density=0.05, random_state=rng, format="coo")
    Y.data = np.round(Y.data * 10.0 ** rng.integers(-3, 6, size=Y.nnz), 6) + 1e-3
    perm = rng.permutation(Y.nnz)
    p2 = tmp_path / "real.mtx"
    with open(p2, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% a comment\n")
        f.write("%d %d %d\n" % (300, 200, Y.nnz))
        for t in perm:
            f.write("%d %d %.10e\n" % (Y.row[t] + 1, Y.col[t] + 1, Y.data[t]))
    ref = sp.csc_matrix(scipy.io.mmread(str(p2)))
    ref.sort_indices()
    with Engine.from_mtx(str(p2)) as eng:                      # (empty rows/cells are refused at
        got = eng.csc()                                        # set_state, not at creation)
    assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
    assert np.max(np.abs(got.data - ref.data) / ref.data) < 4e-16
    # errors: wrong entry count, index out of range, unsupported banner
    bad = tmp_path / "bad.mtx"
    bad.write_text("%%MatrixMarket matrix coordinate integer general\n3 3 3\n1 1 1\n2 2 2\n")
    with pytest.raises(_lib.VbnmfError, match="number of entry lines"):
        Engine.from_mtx(str(bad))
    bad.write_text("%%MatrixMarket matrix coordinate integer general\n3 3 2\n1 1 1\n4 2 2\n")
    with pytest.raises(_lib.VbnmfError, match="out of range"):
        Engine.from_mtx(str(bad))
    bad.write_text("%%MatrixMarket matrix array real general\n3 3\n")
    with pytest.raises(_lib.VbnmfError, match="supported"):
        Engine.from_mtx(str(bad))


def test_read_10x_on_the_device_then_factorize(tmp_path):
    from ccfindr_b200 import api
    X = load_counts("pbmc")
    s = api.scNMFSet(X)
    api.write_10x(s, str(tmp_path))
    a = api.read_10x(str(tmp_path))
    b = api.read_10x(str(tmp_path), device=0)
    assert (a.counts != b.counts).nnz == 0 and a.rowData == b.rowData and a.colData == b.colData


def test_job_farm_front_end_in_a_single_process_group():
    """vb_factorize(parallel=True) -- the (run, rank) job farm that replaces Rmpi::mpi.applyLB
    (R/bayesian.R:263) -- inside a one-rank process group equals the serial front end bit for bit
    (the multi-rank version of this check lives in tests/mgpu_check.py and needs 2 GPUs)."""
    import os
    import torch.distributed as dist
    from ccfindr_b200 import api
    X = load_counts("c1s2")
    kw = dict(ranks=[2, 3], nrun=2, verbose=0, Itmax=30, seed=4)
    ser = api.vb_factorize(api.scNMFSet(X), **kw)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29741")
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        par = api.vb_factorize(api.scNMFSet(X), parallel=True, **kw)
    finally:
        dist.destroy_process_group()
    assert list(par.ranks) == list(ser.ranks) and par.metadata["niter"] == ser.metadata["niter"]
    for key in ser.measure:
        assert np.array_equal(par.measure[key], ser.measure[key]), key
    for k in range(len(ser.ranks)):
        assert np.array_equal(par.basis[k], ser.basis[k]) and np.array_equal(par.coeff[k], ser.coeff[k])
