"""Synthetic count matrices for tests and benchmarks.

simulate_whx() follows the recipe of the reference's simulate_whx (R/utils.R:826-846):
w ~ Gamma(shape=aw, scale=bw/aw), h ~ Gamma(shape=ah, scale=bh/ah), x ~ Poisson(w.h), empty
rows/columns dropped.  R's RNG stream cannot be reproduced without R, so the draws come from a
NumPy Philox generator keyed by `seed`.

tenx_like() is the 10x-shaped generator of SURVEY.md section 8(d): W, H ~ Gamma(0.3, mean 1),
per-cell depth ~ LogNormal(0, 0.5), x ~ Poisson(s * d_j * (W H)_ij) with the scalar s bisected to
hit a target nonzero fraction; sampling runs on the GPU in column chunks with one Philox stream
per chunk, so that a shard of cells generated on one GPU equals the same columns of the
single-GPU matrix.
"""
import numpy as np
import scipy.sparse as sp


def _rng(seed, stream=0):
    return np.random.Generator(np.random.Philox(key=[int(seed), int(stream)]))


def simulate_whx(nrow, ncol, rank, aw=0.1, bw=1.0, ah=0.1, bh=1.0, seed=1):
    """R/utils.R:826-846.  Returns dict(w, h, x) with x a dense float64 count matrix."""
    g = _rng(seed)
    # R fills matrix(rgamma(n*rank), nrow, rank) column-major; only the distribution matters here
    w = g.gamma(shape=aw, scale=bw / aw, size=(rank, nrow)).T.copy()
    h = g.gamma(shape=ah, scale=bh / ah, size=(ncol, rank)).T.copy()
    x = g.poisson(w @ h).astype(np.float64)
    i = x.sum(axis=1) > 0
    j = x.sum(axis=0) > 0
    return dict(w=w[i], h=h[:, j], x=x[i][:, j])


def random_init(nrow, ncol, rank, hyper, seed):
    """vb_init(initializer='random') (R/bayesian.R:111-115): w ~ Gamma(aw, scale bw/aw),
    h ~ Gamma(ah, scale bh/ah), from a NumPy Philox stream."""
    g = _rng(seed, stream=1)
    w = g.gamma(shape=hyper["aw"], scale=hyper["bw"] / hyper["aw"], size=(rank, nrow)).T.copy()
    h = g.gamma(shape=hyper["ah"], scale=hyper["bh"] / hyper["ah"], size=(ncol, rank)).T.copy()
    return w, h


def uniform_init(nrow, ncol, rank, seed):
    """init() of the ML path (R/factorize.R:30-38): w, h ~ U(0,1)."""
    g = _rng(seed, stream=2)
    return g.random(size=(rank, nrow)).T.copy(), g.random(size=(ncol, rank)).T.copy()


def fix_empty(csc, seed=0):
    """Give every empty row/column one count at a random position (the API rejects empties,
    R/bayesian.R:244-247)."""
    csc = csc.tocsc()
    n, m = csc.shape
    g = _rng(seed, stream=3)
    er = np.flatnonzero(np.asarray(csc.sum(axis=1)).ravel() == 0)
    ec = np.flatnonzero(np.asarray(csc.sum(axis=0)).ravel() == 0)
    if len(er) == 0 and len(ec) == 0:
        return csc
    rows = np.concatenate([er, g.integers(0, n, size=len(ec))])
    cols = np.concatenate([g.integers(0, m, size=len(er)), ec])
    add = sp.csc_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, m))
    out = (csc + add).tocsc()
    out.sort_indices()
    return out
