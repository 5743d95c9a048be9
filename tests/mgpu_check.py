"""Multi-GPU parity check, launched under torchrun (one rank per GPU) by tests/test_gpu_multi.py:
every rank holds a contiguous, nnz-balanced shard of the cells of a golden case, the engine runs
the reference loop with one NCCL all-reduce per iteration, and the gathered result must match the
golden vectors produced by the reference's own vbnmf_update (1e-9) -- i.e. sharding changes nothing.
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from conftest import RUN_CASES, hyper_dict, load_counts, load_golden, relerr, run_kwargs  # noqa: E402
from ccfindr_b200 import sharding  # noqa: E402
from ccfindr_b200.engine import Comm, Engine  # noqa: E402


def main():
    rank, local, world = (int(os.environ[k]) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.tensor(list(Engine.nccl_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(uid, 0)
    comm = Comm(world, rank, bytes(uid.cpu().tolist()), device=local)
    worst = 0.0
    for case in ("run_pbmc_r3", "run_c1s2_r3_conv", "run_c1s1_r5"):
        g = load_golden(case)
        X = load_counts(RUN_CASES[case])
        kw = run_kwargs(g)
        flags = kw.pop("hyper_update_flags", (True,) * 4)
        b = sharding.balanced_bounds(X.indptr, world)
        c0, c1 = b[rank], b[rank + 1]
        eng = Engine(sharding.shard_csc(X, c0, c1), device=local)
        eng.attach_comm(comm)
        eng.set_state(g["w0"], g["h0"][:, c0:c1])
        res = eng.run(hyper_dict(g["hyper0"]), hyper_update=flags, **kw)
        st = eng.get_state()
        cid = eng.cluster_id()
        assert res["niter"] == int(g["niter"]) and res["stop_reason"] == int(g["stop_reason"])
        errs = [relerr(res["lkh_trace"], g["lkh_trace"]), relerr(res["hyper_trace"], g["hyper_trace"]),
                relerr(res["lml"], g["lml"])]
        for k in ("lw", "ew", "dw"):
            errs.append(relerr(st[k], g[k]))
        for k in ("lh", "eh", "dh"):
            errs.append(relerr(st[k], g[k][:, c0:c1]))
        assert max(errs) < 1e-9, (case, rank, errs)
        assert np.array_equal(cid, g["cid"][c0:c1]), (case, rank)
        worst = max(worst, max(errs))
        eng.close()
    from ccfindr_b200 import synth
    from oracle import bindings as ob
    # mixed storage formats: only the LAST shard holds a non-integer count.  The packed 16-bit
    # layout decides the tile height at rank 20 (split layout), and the gene panels are all-reduced
    # in device order: every rank must fall back to the 8-byte entries together.
    import scipy.sparse as sp
    rng = np.random.default_rng(11)
    n, m, r = 300, 480, 20
    D = rng.poisson(0.25, size=(n, m)).astype(np.float64)
    D[np.arange(n), rng.integers(0, m, n)] += 1.0        # no empty genes
    D[rng.integers(0, n, m), np.arange(m)] += 1.0        # no empty cells
    D[7, m - 1] = 2.5
    Xm = sp.csc_matrix(D)
    w0, h0 = rng.random((n, r)) + 0.1, rng.random((r, m)) + 0.1
    hyp = dict(aw=0.9, bw=1.1, ah=1.2, bh=0.8)
    ref = ob.sparse_vb_run(Xm, w0, h0, hyp, Itmax=3, Tol=0.0)
    b = sharding.balanced_bounds(Xm.indptr, world)
    c0, c1 = b[rank], b[rank + 1]
    eng = Engine(sharding.shard_csc(Xm, c0, c1), device=local)
    eng.attach_comm(comm)
    eng.set_state(w0, h0[:, c0:c1])
    assert eng.layout_info()["format"] != "p16"
    out = eng.run(hyp, Itmax=3, Tol=0.0)
    st = eng.get_state(("ew", "eh"))
    e = max(relerr(out["lkh_trace"], ref["lkh_trace"]), relerr(st["ew"], ref["ew"]),
            relerr(st["eh"], ref["eh"][:, c0:c1]))
    assert e < 1e-9, ("mixed formats", rank, e)
    worst = max(worst, e)
    eng.close()
    # ML path, sharded, against the CPU oracle
    X = load_counts("pbmc")
    n, m = X.shape
    w0, h0 = synth.uniform_init(n, m, 4, 4)
    ref = ob.sparse_ml_run(X, w0, h0, Itmax=25, Tol=1e-7)
    b = sharding.balanced_bounds(X.indptr, world)
    c0, c1 = b[rank], b[rank + 1]
    eng = Engine(sharding.shard_csc(X, c0, c1), device=local)
    eng.attach_comm(comm)
    out = eng.ml_run(w0, h0[:, c0:c1], Itmax=25, Tol=1e-7)
    assert out["niter"] == ref["niter"]
    e = max(relerr(out["lik_trace"], ref["lik_trace"]), relerr(out["w"], ref["w"]),
            relerr(out["h"], ref["h"][:, c0:c1]))
    assert e < 1e-9, ("ml", rank, e)
    eng.close()
    # independent (run, rank) jobs farmed over the ranks == the serial front end, bit for bit
    from ccfindr_b200 import api
    Xc = load_counts("c1s2")
    kw = dict(ranks=[2, 3, 4], nrun=2, verbose=0, Itmax=40, seed=3, device=local)
    par = api.vb_factorize(api.scNMFSet(Xc), parallel=True, **kw)
    ser = api.vb_factorize(api.scNMFSet(Xc), parallel=False, **kw)
    assert list(par.ranks) == list(ser.ranks)
    for key in ser.measure:
        assert np.array_equal(par.measure[key], ser.measure[key]), key
    for k in range(len(ser.ranks)):
        assert np.array_equal(par.basis[k], ser.basis[k]) and np.array_equal(par.coeff[k], ser.coeff[k])
    # one factorization at a time with the cells sharded over the GPUs, through the front end:
    # same measure / basis / coeff as the single-GPU front end (1e-9: the all-reduce changes the
    # order of the sums over cells), same cluster labels
    for extra in (dict(), dict(device_init=True)):
        shd = api.vb_factorize(api.scNMFSet(Xc), shard_cells=True, **kw, **extra)
        one = api.vb_factorize(api.scNMFSet(Xc), **kw, **extra)
        assert list(shd.ranks) == list(one.ranks)
        for key in one.measure:
            assert relerr(shd.measure[key], one.measure[key]) < 1e-9, key
        for k in range(len(one.ranks)):
            assert shd.coeff[k].shape == one.coeff[k].shape
            e = max(e, relerr(shd.basis[k], one.basis[k]), relerr(shd.coeff[k], one.coeff[k]),
                    relerr(shd.dcoeff[k], one.dcoeff[k]))
            assert np.array_equal(api.cluster_id(shd, rank=one.ranks[k]),
                                  api.cluster_id(one, rank=one.ranks[k]))
        assert e < 1e-9, ("shard_cells", extra, e)
    t = torch.tensor([max(worst, e)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("MGPU_OK world=%d worst_rel_err=%.3e" % (world, t.item()))
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
