"""CPU tests (gloo, world_size 2) of the host-side logic of the sharded path: nnz-balanced
partition of the cells, the packed exchange buffer, and the claim the design rests on -- summing
the per-shard W-side statistics and scalars over ranks reproduces the unsharded iteration.  The
per-shard arithmetic here is the CPU oracle (this is a test; the product never routes through it)."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, load_counts, relerr
from ccfindr_b200 import sharding


def test_balanced_bounds_cover_and_balance():
    X = load_counts("pbmc")
    for nr in (1, 2, 3, 8):
        b = sharding.balanced_bounds(X.indptr, nr)
        assert b[0] == 0 and b[-1] == X.shape[1] and len(b) == nr + 1
        assert all(b[i] <= b[i + 1] for i in range(nr))
        nnz = [X.indptr[b[i + 1]] - X.indptr[b[i]] for i in range(nr)]
        assert max(nnz) - min(nnz) <= 2 * np.diff(X.indptr).max()
    b = sharding.balanced_bounds(X.indptr, 4, align=50)
    assert all(v % 50 == 0 for v in b[1:-1])


def test_shard_csc_roundtrip():
    X = load_counts("c1s1")
    b = sharding.balanced_bounds(X.indptr, 3)
    parts = [sharding.shard_csc(X, b[i], b[i + 1]) for i in range(3)]
    assert (sp.hstack(parts).tocsc() != X).nnz == 0


def test_pack_unpack_exchange():
    rng = np.random.default_rng(0)
    Sw, eh, sc = rng.random((7, 3)), rng.random(3), rng.random(5)
    buf = sharding.pack_exchange(Sw, eh, sc)
    assert len(buf) == sharding.exchange_len(7, 3)
    a, b, c = sharding.unpack_exchange(buf, 7, 3)
    assert np.array_equal(a, Sw) and np.array_equal(b, eh) and np.array_equal(c[:5], sc)


WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch, torch.distributed as dist
from conftest import load_counts
from ccfindr_b200 import sharding, synth
from oracle import oracle_dense as od

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
X = load_counts("tiny"); n, m = X.shape; r = 2
hyper = dict(aw=0.9, bw=1.1, ah=1.2, bh=0.8)
w0, h0 = synth.random_init(n, m, r, hyper, 5)
b = sharding.balanced_bounds(X.indptr, world)
c0, c1 = b[rank], b[rank + 1]
Xs = np.asarray(sharding.shard_csc(X, c0, c1).todense())
lw, lh, ehsum_loc = w0.copy(), h0[:, c0:c1].copy(), h0[:, c0:c1].sum(axis=1)
t = torch.from_numpy(ehsum_loc.copy()); dist.all_reduce(t); ehsum = t.numpy()
out = []
for it in range(3):
    # shard-local statistics at (lw, lh): exactly what one rank's sweep produces
    q = Xs / (lw @ lh)
    Sw_loc = q @ lh.T                       # partial over this shard's cells
    Sh = lw.T @ q                           # complete: the sum over genes is shard-local
    buf = torch.from_numpy(sharding.pack_exchange(Sw_loc, np.zeros(r), []))
    dist.all_reduce(buf)                    # the one collective of the iteration
    Sw, _, _ = sharding.unpack_exchange(buf.numpy(), n, r)
    alw = hyper["aw"] + lw * Sw
    bew = hyper["aw"] / hyper["bw"] + ehsum
    ew = alw / bew
    alh = hyper["ah"] + lh * Sh
    beh = hyper["ah"] / hyper["bh"] + ew.sum(axis=0)
    eh = alh / beh[:, None]
    from scipy.special import digamma
    lw = np.maximum(np.exp(digamma(alw)) / bew, od.EPS)
    lh = np.maximum(np.exp(digamma(alh)) / beh[:, None], od.EPS)
    t = torch.from_numpy(eh.sum(axis=1).copy()); dist.all_reduce(t); ehsum = t.numpy()
    out.append((ew.copy(), eh.copy()))
np.savez(sys.argv[2] + ".%d.npz" % rank, ew=out[-1][0], eh=out[-1][1], c0=c0, c1=c1)
dist.destroy_process_group()
'''


def test_two_rank_gloo_iteration_equals_unsharded(tmp_path):
    """world_size 2 over gloo: all-reducing the packed W-side statistics reproduces the unsharded
    reference iteration (oracle_dense.vbnmf_update) exactly up to summation order."""
    from ccfindr_b200 import synth
    from oracle import oracle_dense as od
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    prefix = str(tmp_path / "out")
    import subprocess
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29741", str(script), ROOT, prefix]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    X = load_counts("tiny")
    n, m = X.shape
    hyper = dict(aw=0.9, bw=1.1, ah=1.2, bh=0.8)
    w0, h0 = synth.random_init(n, m, 2, hyper, 5)
    wh = od.vb_init_from(w0, h0)
    for _ in range(3):
        wh = od.vbnmf_update(np.asarray(X.todense()), wh, hyper, od.EPS)
    for rk in range(2):
        z = np.load(prefix + ".%d.npz" % rk)
        assert relerr(z["ew"], wh["ew"]) < 1e-12
        assert relerr(z["eh"], wh["eh"][:, int(z["c0"]):int(z["c1"])]) < 1e-12
