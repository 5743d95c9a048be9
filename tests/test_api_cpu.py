"""CPU tests of the host-side front end (no GPU): argument validation, the label-tuple form of the
dispersion measure against the reference's pairwise definition (R/factorize.R:51-67), and
optimal_rank (R/utils2.R:59-111)."""
import numpy as np
import pytest
import scipy.sparse as sp

from ccfindr_b200 import api


def _dispersion_pairwise(label_runs):
    """connectivity() + dispersion() exactly as R/factorize.R:51-67 (O(m^2))."""
    nc = len(label_runs[0])
    conav = np.zeros(nc * (nc - 1) // 2)
    iu = np.triu_indices(nc, 1)
    for cid in label_runs:
        cid = np.asarray(cid)
        conav += (cid[:, None] == cid[None, :])[iu]
    cnn = conav / len(label_runs)
    return 1.0 / nc + 8.0 * np.sum((cnn - 0.5) ** 2) / nc ** 2


def test_dispersion_from_labels_equals_pairwise_definition():
    rng = np.random.default_rng(0)
    for nrun in (1, 2, 5):
        runs = [rng.integers(1, 5, size=60) for _ in range(nrun)]
        assert abs(api.dispersion_from_labels(runs) - _dispersion_pairwise(runs)) < 1e-12
    same = [np.repeat([1, 2, 3], 20)] * 4
    assert abs(api.dispersion_from_labels(same) - _dispersion_pairwise(same)) < 1e-12


def test_empty_rows_and_columns_are_rejected():
    x = np.ones((4, 5)); x[2] = 0
    with pytest.raises(ValueError, match="empty rows"):
        api._check_no_empty(sp.csc_matrix(x))
    x = np.ones((4, 5)); x[:, 1] = 0
    with pytest.raises(ValueError, match="empty columns"):
        api._check_no_empty(sp.csc_matrix(x))


def test_svd_initializer_with_several_runs_is_an_error():
    s = api.scNMFSet(np.ones((6, 5)))
    with pytest.raises(ValueError, match="SVD initializer"):
        api.vb_factorize(s, ranks=2, nrun=2, initializer="svd2")


def test_vb_init_shapes_and_positivity():
    rng = np.random.default_rng(1)
    mat = sp.csc_matrix(rng.poisson(2.0, size=(40, 30)).astype(float))
    hyper = dict(aw=1.0, bw=1.0, ah=1.0, bh=1.0)
    for init in ("random", "svd", "svd2"):
        w, h = api.vb_init(40, 30, mat, 3, hyper, init, seed=3)
        assert w.shape == (40, 3) and h.shape == (3, 30)
        assert np.isfinite(w).all() and np.isfinite(h).all() and (w >= 0).all() and (h >= 0).all()
    w, h = api.vb_init(40, 30, mat, 3, hyper, "svd2", seed=3)
    assert abs(h.mean() - hyper["bh"]) < 1e-12   # scale <- bh/mean(h), R/bayesian.R:156-158


def test_smooth_spline_limits():
    x = np.arange(2, 12, dtype=float)
    y = np.sin(x / 3.0) + 0.05 * np.cos(7 * x)
    assert np.allclose(api._smooth_spline(x, y, len(x)), y)          # df = n interpolates
    lin = api._smooth_spline(x, y, 2.0)                               # df = 2 is the LS line
    a, b = np.polyfit(x, y, 1)
    assert np.allclose(lin, a * x + b, atol=1e-6)
    mid = api._smooth_spline(x, y, 5.0)
    assert np.sum((mid - y) ** 2) < np.sum((lin - y) ** 2)


def test_optimal_rank_types():
    ranks = np.arange(2, 11)
    peak = -0.7 - 0.01 * (ranks - 5.0) ** 2                          # clear maximum at 5 -> type 1
    out = api.optimal_rank(dict(rank=ranks, lml=peak), m=1000)
    assert out["type"] == 1 and out["ropt"] == 5.0
    plateau = -0.7 - 0.05 * np.exp(-(ranks - 2.0))                   # saturating -> type 2
    out = api.optimal_rank(dict(rank=ranks, lml=plateau), m=1000)
    assert out["type"] == 2 and 3.0 <= out["ropt"] <= 10.0
    with pytest.raises(ValueError):
        api.optimal_rank(dict(rank=ranks, lml=peak))


def test_lpt_schedule_balances_rank_sweep():
    ranks = list(range(2, 31)) * 5                       # BASELINE config 4: ranks 2..30 x 5 restarts
    plan = api.lpt_schedule([float(r) for r in ranks], 8)
    assert sorted(i for w in plan for i in w) == list(range(len(ranks)))
    loads = [sum(ranks[i] for i in w) for w in plan]
    assert max(loads) - min(loads) <= max(ranks)
    assert api.lpt_schedule([3.0, 1.0], 4) == [[0], [1], [], []]


def test_read_write_10x_roundtrip(tmp_path):
    from conftest import load_counts
    X = load_counts("tiny").tolil()
    X[5, :] = 0                                            # an empty gene: dropped on read
    s = api.scNMFSet(sp.csc_matrix(X), rowData=[("g%d" % i, "sym%d" % i) for i in range(X.shape[0])],
                     colData=["cell%d" % j for j in range(X.shape[1])])
    api.write_10x(s, str(tmp_path))
    back = api.read_10x(str(tmp_path), remove_zeros_=False)
    assert (back.counts != s.counts).nnz == 0 and back.rowData[3] == ("g3", "sym3")
    clean = api.read_10x(str(tmp_path))
    assert clean.nrow() == s.nrow() - 1 and ("g5", "sym5") not in clean.rowData
    assert clean.counts.format == "csc" and clean.counts.dtype == np.float64
    with pytest.raises(FileNotFoundError):
        api.read_10x(str(tmp_path / "nope"))


def test_device_random_init_reference_is_a_gamma_sampler():
    """CPU restatement of the on-device initialiser (synth.device_random_init_reference): keyed by
    (seed, position), so a shard of cells reproduces the same columns; moments of Gamma(a, scale b/a)."""
    from ccfindr_b200 import synth
    hyper = dict(aw=0.5, bw=2.0, ah=3.0, bh=0.7)
    w, h = synth.device_random_init_reference(300, 40, 4, hyper, seed=7)
    w2, h2 = synth.device_random_init_reference(300, 40, 4, hyper, seed=7)
    assert np.array_equal(w, w2) and np.array_equal(h, h2)
    _, hs = synth.device_random_init_reference(1, 15, 4, hyper, seed=7, cell_offset=25)
    assert np.array_equal(hs, h[:, 25:40])
    w3, _ = synth.device_random_init_reference(300, 1, 4, hyper, seed=8)
    assert not np.array_equal(w, w3)
    assert (w > 0).all() and (h > 0).all()
    # mean b, variance b^2 / a (1200 and 160 draws: loose bounds)
    assert abs(w.mean() - 2.0) < 0.35 and abs(w.var() - 8.0) < 3.0
    assert abs(h.mean() - 0.7) < 0.12


def _label_runs(rng, m, nrun, rank, noise):
    """labelings that mostly agree with a planted partition (like converged NMF runs)"""
    truth = rng.integers(0, rank, size=m)
    runs = []
    for _ in range(nrun):
        lab = truth.copy()
        flip = rng.random(m) < noise
        lab[flip] = rng.integers(0, rank, size=int(flip.sum()))
        runs.append(rng.permutation(rank)[lab] + 1)            # labels are arbitrary per run
    return runs


@pytest.mark.parametrize("m,nrun,rank,noise", [(40, 4, 3, 0.2), (150, 6, 4, 0.1), (300, 10, 5, 0.3)])
def test_consensus_measures_from_label_groups_match_the_literal_ones(m, nrun, rank, noise):
    """dispersion and the cophenetic correlation computed from groups of cells with identical
    label tuples equal the reference's definitions on the full m(m-1)/2 connectivity vector
    (R/factorize.R:51-78, restated literally in oracle/oracle_dense.py)."""
    from oracle import oracle_dense as od
    rng = np.random.default_rng(m + nrun)
    runs = _label_runs(rng, m, nrun, rank, noise)
    conav = np.zeros(m * (m - 1) // 2)
    for lab in runs:
        h = np.zeros((rank, m)); h[lab - 1, np.arange(m)] = 1.0
        conav += od.connectivity(h)
    conav /= nrun
    assert abs(api.dispersion_from_labels(runs) - od.dispersion(conav, m)) < 1e-12
    for method in ("average", "single"):
        got = api.cophenet_from_labels(runs, method)
        ref = od.cophenet(conav, m, method)
        # connectivity distances are multiples of 1/nrun, so equal distances abound and the two
        # agglomerations break them in a different order: 'single' does not depend on that order,
        # 'average' (the reference's default) barely; 'complete' trees are not unique under ties
        # (checked tie-free in the next test)
        assert abs(got - ref) < (1e-9 if method == "single" else 5e-3), (method, got, ref)


def test_cophenetic_scales_with_the_number_of_label_tuples_not_cells():
    rng = np.random.default_rng(5)
    runs = _label_runs(rng, 200000, 8, 4, 0.02)
    c = api.cophenet_from_labels(runs)
    d = api.dispersion_from_labels(runs)
    assert 0.9 < c <= 1.0 and 0.5 < d <= 1.0


@pytest.mark.parametrize("method", ["average", "single", "complete", "mcquitty", "ward.D"])
def test_weighted_linkage_equals_linkage_of_the_expanded_point_set(method):
    """Tie-free check of the agglomeration on groups of identical points: random distances between
    G groups of random sizes against scipy on the expanded set (zero distance inside a group)."""
    from scipy.cluster.hierarchy import cophenet, linkage
    from scipy.spatial.distance import squareform
    rng = np.random.default_rng(3)
    G = 9
    sizes = rng.integers(1, 5, size=G)
    D = rng.random((G, G)) + 0.5
    D = (D + D.T) / 2
    np.fill_diagonal(D, 0.0)
    owner = np.repeat(np.arange(G), sizes)
    full = D[np.ix_(owner, owner)]
    z = linkage(squareform(full, checks=False),
                method={"mcquitty": "weighted", "ward.D": "ward"}.get(method, method))
    if method == "ward.D":
        pytest.skip("scipy's ward works on Euclidean distances (ward.D2); no tie-free twin here")
    cf = squareform(cophenet(z))
    got = api._linkage_weighted(D, sizes.astype(float), method)
    first = np.array([np.flatnonzero(owner == g)[0] for g in range(G)])
    ref = cf[np.ix_(first, first)]
    assert np.allclose(got + np.diag(np.diag(ref)), ref, atol=1e-12)


def test_job_farm_exchanges_only_the_matrices_that_can_be_chosen():
    """The best run per rank index by strict > (first of equals, NaN never wins); everything when a
    job raised the uniform-column flag."""
    nan = float("nan")
    sc = {(1, 0): dict(rdat=-2.0, unif=False), (2, 0): dict(rdat=-1.0, unif=False),
          (3, 0): dict(rdat=-1.0, unif=False),
          (1, 1): dict(rdat=nan, unif=False), (2, 1): dict(rdat=-5.0, unif=False),
          (3, 1): dict(rdat=-7.0, unif=False),
          (1, 2): dict(rdat=nan, unif=False), (2, 2): dict(rdat=nan, unif=False),
          (3, 2): dict(rdat=nan, unif=False)}
    assert api.jobs_to_keep(sc, 3, 3) == {(2, 0), (2, 1)}
    sc[(3, 1)]["unif"] = True
    assert api.jobs_to_keep(sc, 3, 3) == set(sc)
