"""TEST INFRASTRUCTURE ONLY (oracle/).  ctypes loaders for the two CPU checkers:

  sparse : oracle/_build/liboracle_sparse.so  (oracle_sparse.c, the C restatement)
  ref    : oracle/_ref/libccfindr_ref.so      (the reference's own src/vbnmf_update.cpp compiled
                                               in place against oracle/shim/; may be absent)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SPARSE_SO = os.path.join(HERE, "_build", "liboracle_sparse.so")
REF_SO = os.path.join(HERE, "_ref", "libccfindr_ref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)


def build(ref_dir="/root/reference", quiet=True):
    """Compile the checkers (oracle/Makefile).  The reference binary is only rebuilt where the
    reference sources exist (the build container); the GPU box uses the prebuilt files."""
    cmd = ["make", "-C", HERE, f"REF={ref_dir}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    if not quiet:
        print(res.stdout)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_sparse = None
_ref = None


def sparse_lib():
    global _sparse
    if _sparse is None:
        if not os.path.exists(SPARSE_SO):
            build()
        lib = C.CDLL(SPARSE_SO)
        lib.osp_digamma.restype = C.c_double
        lib.osp_digamma.argtypes = [C.c_double]
        lib.osp_trigamma.restype = C.c_double
        lib.osp_trigamma.argtypes = [C.c_double]
        lib.osp_lgx_sum.restype = C.c_double
        lib.osp_time_iterations.restype = C.c_double
        lib.osp_num_threads.restype = C.c_int
        _sparse = lib
    return _sparse


def ref_lib():
    """The compiled reference source, or None when it is not available."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            try:
                build()
            except Exception:
                return None
        if not os.path.exists(REF_SO):
            return None
        _ref = C.CDLL(REF_SO)
        _ref.ref_source_path.restype = C.c_char_p
    return _ref


def ref_vbnmf_update(X, wh, hyper, fudge):
    """Call the REFERENCE's vbnmf_update (src/vbnmf_update.cpp:16-102) on dense X (n x m).
    wh: dict with lw, ew (n x r), lh, eh (r x m); hyper dict aw, bw, ah, bh."""
    lib = ref_lib()
    if lib is None:
        raise RuntimeError("oracle/_ref/libccfindr_ref.so is not available")
    X = np.asfortranarray(X, dtype=np.float64)
    n, m = X.shape
    lw = np.asfortranarray(wh["lw"], dtype=np.float64)
    lh = np.asfortranarray(wh["lh"], dtype=np.float64)
    ew = np.asfortranarray(wh["ew"], dtype=np.float64)
    eh = np.asfortranarray(wh["eh"], dtype=np.float64)
    r = lw.shape[1]
    hy = np.array([hyper["aw"], hyper["bw"], hyper["ah"], hyper["bh"]], dtype=np.float64)
    out = {k: np.zeros((n, r), order="F") for k in ("lw", "ew", "dw")}
    out.update({k: np.zeros((r, m), order="F") for k in ("lh", "eh", "dh")})
    lkh = C.c_double(0.0)
    rc = lib.ref_vbnmf_update(C.c_int(n), C.c_int(m), C.c_int(r), _p(X, _dp), _p(lw, _dp),
                              _p(lh, _dp), _p(ew, _dp), _p(eh, _dp), _p(hy, _dp),
                              C.c_double(float(fudge)), _p(out["lw"], _dp), _p(out["lh"], _dp),
                              _p(out["ew"], _dp), _p(out["eh"], _dp), _p(out["dw"], _dp),
                              _p(out["dh"], _dp), C.byref(lkh))
    if rc != 0:
        raise RuntimeError("ref_vbnmf_update failed")
    out["lkh"] = lkh.value
    out["w"], out["h"] = out["ew"], out["eh"]
    return out


def _csc_args(csc):
    """csc: scipy.sparse CSC matrix (any dtype).  Returns (n, m, colptr64, rowidx32, val64)."""
    csc = csc.tocsc()
    csc.sort_indices()
    n, m = csc.shape
    colptr = np.ascontiguousarray(csc.indptr, dtype=np.int64)
    rowidx = np.ascontiguousarray(csc.indices, dtype=np.int32)
    val = np.ascontiguousarray(csc.data, dtype=np.float64)
    return n, m, colptr, rowidx, val


def sparse_vb_step(csc, wh, hyper, fudge):
    """oracle_sparse.c osp_vb_step: one reference step on CSC input.  Same dict in/out as
    oracle_dense.vbnmf_update; extra key 'means' = inputs of hyper_update."""
    lib = sparse_lib()
    n, m, colptr, rowidx, val = _csc_args(csc)
    lw = np.array(wh["lw"], dtype=np.float64, order="F")
    lh = np.array(wh["lh"], dtype=np.float64, order="F")
    eh_in = np.asfortranarray(wh["eh"], dtype=np.float64)
    r = lw.shape[1]
    ew = np.zeros((n, r), order="F"); dw = np.zeros((n, r), order="F")
    eh = np.zeros((r, m), order="F"); dh = np.zeros((r, m), order="F")
    hy = np.array([hyper["aw"], hyper["bw"], hyper["ah"], hyper["bh"]], dtype=np.float64)
    lkh = C.c_double(0.0)
    means = np.zeros(4)
    rc = lib.osp_vb_step(C.c_int64(n), C.c_int64(m), C.c_int(r), _p(colptr, _i64p),
                         _p(rowidx, _i32p), _p(val, _dp), _p(lw, _dp), _p(lh, _dp), _p(eh_in, _dp),
                         _p(ew, _dp), _p(eh, _dp), _p(dw, _dp), _p(dh, _dp), _p(hy, _dp),
                         C.c_double(float(fudge)), C.byref(lkh), _p(means, _dp))
    if rc != 0:
        raise RuntimeError("osp_vb_step failed rc=%d" % rc)
    return dict(w=ew, h=eh, lw=lw, lh=lh, ew=ew, eh=eh, dw=dw, dh=dh, lkh=lkh.value, means=means)


def set_omp_threads(nthreads):
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms of bench.py restore the
    host's thread count (process-wide libgomp setting)."""
    for name in ("libgomp.so.1", "libgomp.so"):
        try:
            C.CDLL(name).omp_set_num_threads(int(nthreads))
            return True
        except OSError:
            continue
    return False


def sparse_vb_run(csc, w0, h0, hyper, *, Itmax=10000, hyper_update_flags=(True,) * 4, Tol=1e-5,
                  n0=10, dn=1, fudge=np.finfo(np.float64).eps):
    """oracle_sparse.c osp_vb_run: the loop of vb_iterate for one rank (R/bayesian.R:336-352).
    csc: a scipy.sparse matrix, or the tuple (n, m, colptr int64, rowidx int32, val float64) of
    sorted CSC arrays (no copies are made: matrices of 2e9 nonzeros)."""
    lib = sparse_lib()
    if isinstance(csc, tuple):
        n, m, colptr, rowidx, val = csc
        assert colptr.dtype == np.int64 and rowidx.dtype == np.int32 and val.dtype == np.float64
    else:
        n, m, colptr, rowidx, val = _csc_args(csc)
    lw = np.array(w0, dtype=np.float64, order="F")
    lh = np.array(h0, dtype=np.float64, order="F")
    r = lw.shape[1]
    ew = np.zeros((n, r), order="F"); dw = np.zeros((n, r), order="F")
    eh = np.zeros((r, m), order="F"); dh = np.zeros((r, m), order="F")
    hy = np.array([hyper["aw"], hyper["bw"], hyper["ah"], hyper["bh"]], dtype=np.float64)
    cfg_i = np.array([Itmax, n0, dn] + [int(bool(f)) for f in hyper_update_flags], dtype=np.int32)
    cfg_d = np.array([Tol, fudge], dtype=np.float64)
    trace = np.full(Itmax, np.nan)
    htrace = np.full((Itmax, 4), np.nan)
    niter = C.c_int(0); reason = C.c_int(0); lml = C.c_double(0.0)
    rc = lib.osp_vb_run(C.c_int64(n), C.c_int64(m), C.c_int(r), _p(colptr, _i64p),
                        _p(rowidx, _i32p), _p(val, _dp), _p(lw, _dp), _p(lh, _dp), _p(ew, _dp),
                        _p(eh, _dp), _p(dw, _dp), _p(dh, _dp), _p(cfg_i, _ip), _p(cfg_d, _dp),
                        _p(hy, _dp), _p(trace, _dp), _p(htrace, _dp), C.byref(niter),
                        C.byref(lml), C.byref(reason))
    if rc == 2:
        from .oracle_dense import HyperUpdateError
        raise HyperUpdateError("Hyper-parameter update failed to converge")
    if rc != 0:
        raise RuntimeError("osp_vb_run failed rc=%d" % rc)
    it = niter.value
    return dict(lw=lw, lh=lh, ew=ew, eh=eh, dw=dw, dh=dh,
                hyper=dict(aw=hy[0], bw=hy[1], ah=hy[2], bh=hy[3]), lml=lml.value, niter=it,
                lkh_trace=trace[:it].copy(), hyper_trace=htrace[:it].copy(),
                stop_reason=reason.value)


def sparse_ml_run(csc, w0, h0, *, Itmax=10000, Tol=1e-5):
    """oracle_sparse.c osp_ml_run (R/factorize.R:189-212, criterion='likelihood').  csc: a scipy
    matrix or the tuple (n, m, colptr int64, rowidx int32, val float64)."""
    lib = sparse_lib()
    n, m, colptr, rowidx, val = csc if isinstance(csc, tuple) else _csc_args(csc)
    w = np.array(w0, dtype=np.float64, order="F")
    h = np.array(h0, dtype=np.float64, order="F")
    r = w.shape[1]
    trace = np.full(Itmax, np.nan)
    niter = C.c_int(0)
    rc = lib.osp_ml_run(C.c_int64(n), C.c_int64(m), C.c_int(r), _p(colptr, _i64p),
                        _p(rowidx, _i32p), _p(val, _dp), _p(w, _dp), _p(h, _dp), C.c_int(Itmax),
                        C.c_double(Tol), _p(trace, _dp), C.byref(niter))
    if rc != 0:
        raise RuntimeError("osp_ml_run failed rc=%d" % rc)
    it = niter.value
    return dict(w=w, h=h, niter=it, lik_trace=trace[:it].copy(), lik=trace[it - 1])


def sparse_time_iterations(csc_arrays, lw, lh, hyper4, fudge, iters):
    """Seconds per steady-state VB iteration of the C oracle (bench.py cpu_baseline).
    csc_arrays = (n, m, colptr int64, rowidx int32, val float64)."""
    lib = sparse_lib()
    n, m, colptr, rowidx, val = csc_arrays
    lw = np.asfortranarray(lw, dtype=np.float64)
    lh = np.asfortranarray(lh, dtype=np.float64)
    r = lw.shape[1]
    hy = np.asarray(hyper4, dtype=np.float64)
    lkh = C.c_double(0.0)
    sec = lib.osp_time_iterations(C.c_int64(n), C.c_int64(m), C.c_int(r), _p(colptr, _i64p),
                                  _p(rowidx, _i32p), _p(val, _dp), _p(lw, _dp), _p(lh, _dp),
                                  _p(hy, _dp), C.c_double(float(fudge)), C.c_int(iters),
                                  C.byref(lkh))
    return sec, lkh.value, lib.osp_num_threads()
