"""Parity at BASELINE size (run on the B200 box): the CUDA path against the CPU oracle on the
whole C2 matrix (20,000 genes x 100,000 cells, ~8 % nonzero, rank 10) and on a C3-shaped matrix
(same genes, 100,000 cells of the C3 generator, rank 20) -- the problems whose tiles have the full
height (T = 2880 / 1312 rows), which the small parity cases never reach.  fp64: bound of every
iteration, ew, eh within 1e-9 relative and cluster ids bit-exact; fp32-storage mode: 1e-4 and the
number of cells whose cluster id differs is reported (allowed only where the two largest entries
of the oracle's eh column are within 1e-4 relative of each other)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,cells", [("c2", 100000), ("c3", 100000)])
def test_bench_size_matrix_matches_oracle(name, cells):
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import bench
    from ccfindr_b200 import synth
    from ccfindr_b200.engine import Engine
    wl = bench.WORKLOADS[name]
    n, r = wl["n"], wl["rank"]
    dev = torch.device("cuda", 0)
    colptr, rowidx, values, _ = synth.tenx_like_device(n, cells, wl["r_true"], wl["density"],
                                                       wl["seed"], dev, 0, cells)
    w0, h0 = bench.init_factors(n, cells, r, seed=1000 * r + 1)
    h0 = np.asfortranarray(h0)
    with Engine.from_device_csc(n, cells, int(rowidx.numel()), colptr, rowidx, values) as eng:
        g64 = bench.gpu_parity_run(eng, w0, h0, 3, 0)
        lay = eng.layout_info()
        g32 = bench.gpu_parity_run(eng, w0, h0, 3, 1)
    assert lay["format"] == "p16" and lay["tile_rows"] >= 1280     # full-height tiles
    blk, cpu = bench.oracle_parity(n, cells, colptr.cpu().numpy(), rowidx.cpu().numpy(),
                                   values.double().cpu().numpy(), w0, h0, 3, g64, g32)
    print(name, {k: blk[k] for k in ("fp64", "fp32_storage")})
    f64, f32 = blk["fp64"], blk["fp32_storage"]
    assert max(f64["lkh_rel_err"]) < 1e-9
    assert f64["ew_max_rel_err"] < 1e-9 and f64["eh_max_rel_err"] < 1e-9
    assert f64["cid_mismatches"] == 0                               # bit-exact cluster ids
    assert max(f32["lkh_rel_err"]) < 1e-4
    assert f32["ew_max_rel_err"] < 1e-4 and f32["eh_max_rel_err"] < 1e-4
    assert f32["cid_mismatches"] <= cells // 1000                   # near-ties only


def test_ml_path_at_bench_size_matches_oracle():
    """BASELINE config 5: factorize() maximum-likelihood updates on the 20,000 x 100,000 matrix,
    rank 15 -- three iterations of the device loop against oracle_sparse.c (1e-9)."""
    import torch
    import bench
    from ccfindr_b200 import synth
    from ccfindr_b200.engine import Engine
    from oracle import bindings as ob
    wl = bench.WORKLOADS["c2"]
    n, cells, r = wl["n"], 100000, 15
    dev = torch.device("cuda", 0)
    colptr, rowidx, values, _ = synth.tenx_like_device(n, cells, wl["r_true"], wl["density"],
                                                       wl["seed"], dev, 0, cells)
    w0, h0 = synth.uniform_init(n, cells, r, 5)
    bench.restore_omp_threads()
    ref = ob.sparse_ml_run((n, cells, colptr.cpu().numpy(), rowidx.cpu().numpy(),
                            values.double().cpu().numpy()), w0, h0, Itmax=3, Tol=0.0)
    with Engine.from_device_csc(n, cells, int(rowidx.numel()), colptr, rowidx, values) as eng:
        res = eng.ml_run(w0, h0, Itmax=3, Tol=0.0)
        lay = eng.layout_info()
    assert lay["tile_rows"] >= 1280 and res["niter"] == ref["niter"] == 3
    err = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    print("ml c5", err(res["lik_trace"], ref["lik_trace"]), err(res["w"], ref["w"]), err(res["h"], ref["h"]))
    assert err(res["lik_trace"], ref["lik_trace"]) < 1e-9
    assert err(res["w"], ref["w"]) < 1e-9 and err(res["h"], ref["h"]) < 1e-9
