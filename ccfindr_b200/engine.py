"""Thin object wrapper over the C ABI: one Engine = one vbnmf_handle = one GPU holding the CSC
count matrix (or one shard of its cells).  Host code only; all arithmetic runs in libvbnmf.so."""
import ctypes as C

import numpy as np

from . import _lib

EPS = float(np.finfo(np.float64).eps)  # .Machine$double.eps (R/bayesian.R:238)
HKEYS = ("aw", "bw", "ah", "bh")


def _dp(a):
    return None if a is None else a.ctypes.data_as(_lib.c_dp)


def hyper_vec(hyper):
    if isinstance(hyper, dict):
        return np.array([float(hyper[k]) for k in HKEYS], dtype=np.float64)
    return np.array(hyper, dtype=np.float64)


def hyper_dict(v):
    return {k: float(x) for k, x in zip(HKEYS, v)}


class Engine:
    """Device-resident VB-NMF state for one count matrix.

    counts: scipy.sparse matrix (genes x cells) or anything scipy can convert to CSC -- the
    `counts(object)` of an scNMFSet (R/bayesian.R:239).
    """

    def __init__(self, counts=None, device=0, _device_csc=None):
        self.lib = _lib.load()
        self.handle = C.c_void_p()
        self._keep = None
        if _device_csc is not None:
            n, m, nnz, d_colptr, d_rowidx, d_val, keep = _device_csc
            self._keep = keep  # borrowed device arrays must outlive the handle
            rc = self.lib.vbnmf_create_from_device(C.byref(self.handle), n, m, nnz,
                                                   C.c_void_p(d_colptr), C.c_void_p(d_rowidx),
                                                   C.c_void_p(d_val), int(device))
        else:
            import scipy.sparse as sp
            csc = counts if sp.isspmatrix_csc(counts) else sp.csc_matrix(counts)
            if not csc.has_sorted_indices:
                csc = csc.copy()
                csc.sort_indices()
            n, m = csc.shape
            nnz = csc.nnz
            rowidx = np.ascontiguousarray(csc.indices, dtype=np.int32)
            values = np.ascontiguousarray(csc.data, dtype=np.float64)
            p32 = p64 = None
            if csc.indptr.dtype == np.int32:  # dgCMatrix @p is int32
                ptr = np.ascontiguousarray(csc.indptr)
                p32 = ptr.ctypes.data_as(_lib.c_i32p)
            else:
                ptr = np.ascontiguousarray(csc.indptr, dtype=np.int64)
                p64 = ptr.ctypes.data_as(_lib.c_i64p)
            rc = self.lib.vbnmf_create(C.byref(self.handle), n, m, nnz, p32, p64,
                                       rowidx.ctypes.data_as(_lib.c_i32p), _dp(values), int(device))
        if rc != 0:
            msg = self.lib.vbnmf_last_error(None)
            self.handle = C.c_void_p()
            raise _lib.VbnmfError(rc, msg.decode() if msg else "unknown")
        self.n, self.m, self.nnz = int(n), int(m), int(nnz)
        self.r = 0

    @classmethod
    def from_device_csc(cls, n, m, nnz, colptr_i64, rowidx_i32, values_f32, device=0):
        """colptr/rowidx/values: torch CUDA tensors (int64, int32, float32) already on `device`.
        The library reads them on its own stream: whatever produced them (the caller's current
        torch stream) is waited for first -- without this the scan of the counts can overtake a
        still running torch.cat and miss the tail of the matrix (seen as a constant offset of the
        bound: the sum of lgamma(x + 1) came out short while the factors were right)."""
        import torch
        torch.cuda.current_stream(torch.device("cuda", int(device))).synchronize()
        keep = (colptr_i64, rowidx_i32, values_f32)
        return cls(device=device, _device_csc=(int(n), int(m), int(nnz), colptr_i64.data_ptr(),
                                               rowidx_i32.data_ptr(), values_f32.data_ptr(), keep))

    @classmethod
    def from_mtx(cls, path, device=0):
        """MatrixMarket coordinate file (10x `matrix.mtx`) parsed and sorted into CSC on the GPU
        (readMM + as(., 'dgCMatrix') of read_10x, R/utils.R:34)."""
        self = cls.__new__(cls)
        self.lib = _lib.load()
        self.handle = C.c_void_p()
        self._keep = None
        dims = (C.c_int64 * 3)()
        rc = self.lib.vbnmf_create_from_mtx(C.byref(self.handle), str(path).encode(), int(device), dims)
        if rc != 0:
            msg = self.lib.vbnmf_last_error(None)
            self.handle = C.c_void_p()
            raise _lib.VbnmfError(rc, msg.decode() if msg else "unknown")
        self.n, self.m, self.nnz = int(dims[0]), int(dims[1]), int(dims[2])
        self.r = 0
        return self

    def csc(self):
        """The count matrix the handle holds, as a scipy CSC matrix."""
        import scipy.sparse as sp
        colptr = np.zeros(self.m + 1, dtype=np.int64)
        rowidx = np.zeros(self.nnz, dtype=np.int32)
        values = np.zeros(self.nnz, dtype=np.float64)
        self._ck(self.lib.vbnmf_get_csc(self.handle, colptr.ctypes.data_as(_lib.c_i64p),
                                        rowidx.ctypes.data_as(_lib.c_i32p), _dp(values)))
        return sp.csc_matrix((values, rowidx, colptr), shape=(self.n, self.m))

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.vbnmf_destroy(self.handle)
            self.handle = C.c_void_p()
        self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        _lib.check(rc, self.handle)

    # -- configuration -------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.vbnmf_set_stream(self.handle, C.c_void_p(int(cuda_stream_ptr))))

    def set_precision(self, precision):
        self._ck(self.lib.vbnmf_set_precision(self.handle, int(precision)))

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        rc = _lib.load().vbnmf_nccl_unique_id(buf)
        if rc != 0:
            _lib.check(rc, None)
        return bytes(buf)

    def attach_comm(self, comm):
        """comm: a Comm (one NCCL communicator per process); cells are sharded over its ranks."""
        self._ck(self.lib.vbnmf_attach_comm(self.handle, comm.ptr))
        self._comm = comm

    def info(self):
        v = (C.c_int64 * 8)()
        self._ck(self.lib.vbnmf_info(self.handle, v))
        return dict(zip(("n", "m", "nnz", "r", "rs", "precision", "nranks", "m_global"), v))

    def layout_info(self):
        """Tiled device layout in use (after set_state): storage format of the nonzeros
        ('f32', 'f64' or 'p16'), tile rows, slab counts, stored entries per pass, bytes."""
        v = (C.c_int64 * 8)()
        self._ck(self.lib.vbnmf_layout_info(self.handle, v))
        d = dict(zip(("format", "tile_rows", "gene_slabs", "cell_slabs", "entries_cols",
                      "entries_rows", "nonzeros_per_group_step", "bytes"), v))
        d["format"] = ("f32", "f64", "p16")[d["format"]]
        return d

    # -- state ---------------------------------------------------------------------------------
    def set_state(self, lw, lh, ew=None, eh=None):
        """wh list of vbnmf_update (src/vbnmf_update.cpp:22-25): lw, ew n x r; lh, eh r x m."""
        lw = np.asfortranarray(lw, dtype=np.float64)
        lh = np.asfortranarray(lh, dtype=np.float64)
        r = lw.shape[1]
        if lw.shape != (self.n, r) or lh.shape != (r, self.m):
            raise ValueError("lw must be n x r and lh r x m")
        ew = None if ew is None else np.asfortranarray(ew, dtype=np.float64)
        eh = None if eh is None else np.asfortranarray(eh, dtype=np.float64)
        self._ck(self.lib.vbnmf_set_state(self.handle, r, _dp(lw), _dp(lh), _dp(ew), _dp(eh)))
        self.r = r

    def init_random(self, rank, hyper, seed, cell_offset=0):
        """vb_init(initializer='random') (R/bayesian.R:111-115) drawn on the device: no host
        matrices, no upload.  The draw is keyed by (seed, gene / global cell index, k); see
        synth.device_random_init_reference for the CPU restatement."""
        hy = hyper_vec(hyper)
        self._ck(self.lib.vbnmf_init_random(self.handle, int(rank), _dp(hy), int(seed),
                                            int(cell_offset)))
        self.r = int(rank)

    def init_svd2(self, rank, hyper, seed=1, cell_offset=0):
        """vb_init(initializer='svd2') (R/bayesian.R:150-159) computed on the device from the
        resident count matrix (randomized truncated SVD): no host SVD, no upload."""
        hy = hyper_vec(hyper)
        self._ck(self.lib.vbnmf_init_svd2(self.handle, int(rank), _dp(hy), int(seed),
                                          int(cell_offset)))
        self.r = int(rank)

    def get_state(self, which=("lw", "lh", "ew", "eh", "dw", "dh")):
        r = self.r
        out = {}
        for k in which:
            out[k] = np.zeros((self.n, r) if k[1] == "w" else (r, self.m), order="F")
        args = [_dp(out.get(k)) for k in ("lw", "lh", "ew", "eh", "dw", "dh")]
        self._ck(self.lib.vbnmf_get_state(self.handle, *args))
        return out

    def step(self, hyper, fudge=EPS):
        """One vbnmf_update (src/vbnmf_update.cpp:16-102).  Returns lkh."""
        hy = hyper_vec(hyper)
        lkh = C.c_double(0.0)
        self._ck(self.lib.vbnmf_step(self.handle, _dp(hy), float(fudge), C.byref(lkh)))
        return lkh.value

    def means(self):
        v = np.zeros(4)
        self._ck(self.lib.vbnmf_get_means(self.handle, _dp(v)))
        return v

    def run(self, hyper, Itmax=10000, Tol=1e-5, hyper_update=(True,) * 4, n0=10, dn=1, fudge=EPS):
        """The it-loop of vb_iterate (R/bayesian.R:336-352).  Returns a dict with the final hyper,
        lml (= lk0 of :379), niter, stop_reason and the per-iteration traces."""
        cfg = _lib.VbnmfCfg()
        cfg.itmax = int(Itmax)
        cfg.tol = float(Tol)
        for i in range(4):
            cfg.hyper_update[i] = int(bool(hyper_update[i]))
        cfg.n0, cfg.dn, cfg.fudge = int(n0), int(dn), float(fudge)
        hy = hyper_vec(hyper)
        trace = np.full(cfg.itmax, np.nan)
        htrace = np.full((cfg.itmax, 4), np.nan)
        niter, reason, lml = C.c_int(0), C.c_int(0), C.c_double(0.0)
        self._ck(self.lib.vbnmf_run(self.handle, C.byref(cfg), _dp(hy), _dp(trace), _dp(htrace),
                                    C.byref(niter), C.byref(lml), C.byref(reason)))
        it = niter.value
        return dict(hyper=hyper_dict(hy), lml=lml.value, niter=it, stop_reason=reason.value,
                    lkh_trace=trace[:it].copy(), hyper_trace=htrace[:it].copy())

    def cluster_id(self):
        cid = np.zeros(self.m, dtype=np.int32)
        self._ck(self.lib.vbnmf_cluster_id(self.handle, cid.ctypes.data_as(_lib.c_i32p)))
        return cid

    def uniform_columns(self, tol):
        fl = np.zeros(self.r, dtype=np.int32)
        self._ck(self.lib.vbnmf_uniform_columns(self.handle, float(tol),
                                                fl.ctypes.data_as(_lib.c_i32p)))
        return fl.astype(bool)

    # -- ML path -------------------------------------------------------------------------------
    def ml_run(self, w0, h0, Itmax=10000, Tol=1e-5, criterion="likelihood", ncnn_step=40):
        """it-loop of factorize() (R/factorize.R:189-212) with either stopping criterion:
        'likelihood' (:207) or 'connectivity' (:194-204, unchanged cell co-clustering for
        ncnn_step iterations)."""
        if criterion not in ("likelihood", "connectivity"):
            raise ValueError("Unknown stopping criterion.")
        w0 = np.asfortranarray(w0, dtype=np.float64)
        h0 = np.asfortranarray(h0, dtype=np.float64)
        r = w0.shape[1]
        w = np.zeros((self.n, r), order="F")
        h = np.zeros((r, self.m), order="F")
        trace = np.full(int(Itmax), np.nan)
        nch = np.full(int(Itmax), np.nan)
        niter = C.c_int(0)
        self._ck(self.lib.mlnmf_run2(self.handle, r, _dp(w0), _dp(h0), int(Itmax), float(Tol),
                                     int(criterion == "connectivity"), int(ncnn_step), _dp(w),
                                     _dp(h), _dp(trace), _dp(nch), C.byref(niter)))
        self.r = r
        it = niter.value
        return dict(w=w, h=h, niter=it, lik_trace=trace[:it].copy(), lik=float(trace[it - 1]),
                    nchange_trace=nch[:it].copy())

    # -- measurement ---------------------------------------------------------------------------
    def bench_iterations(self, hyper, iters, fudge=EPS, hyper_on=False):
        """`iters` iterations of the device-controlled product loop, CUDA-event timed.  hyper_on:
        hyper_update after every iteration (the loop's steady state past hyper.update.n0); the
        updated hyper-parameters come back in the result."""
        hy = hyper_vec(hyper)
        ms = np.zeros(4)
        launches = C.c_int64(0)
        lkh = C.c_double(0.0)
        self._ck(self.lib.vbnmf_bench_iterations(self.handle, _dp(hy), float(fudge), int(iters),
                                                 int(bool(hyper_on)), _dp(ms), C.byref(launches),
                                                 C.byref(lkh)))
        return dict(ms_total=ms[0], ms_cols=ms[1], ms_rows=ms[2], ms_other=ms[3],
                    launches=launches.value, lkh=lkh.value, hyper=hyper_dict(hy))


def set_host_threads(n):
    """Threads that stage host -> device uploads (process-wide); pass cores / processes when one
    process drives each GPU."""
    _lib.check(_lib.load().vbnmf_set_host_threads(int(n)))


def trim_pool(device=0):
    _lib.check(_lib.load().vbnmf_trim_pool(int(device)))


class Comm:
    """Process-level NCCL communicator for cell-sharded factorizations (vbnmf_comm)."""

    def __init__(self, nranks, rank, uid, device=0):
        self.lib = _lib.load()
        self.ptr = C.c_void_p()
        buf = (C.c_char * 128).from_buffer_copy(uid)
        rc = self.lib.vbnmf_comm_create(C.byref(self.ptr), int(nranks), int(rank), buf, int(device))
        if rc != 0:
            msg = self.lib.vbnmf_last_error(None)
            raise _lib.VbnmfError(rc, msg.decode() if msg else "unknown")
        self.nranks, self.rank = int(nranks), int(rank)

    def close(self):
        if self.ptr and self.ptr.value:
            self.lib.vbnmf_comm_destroy(self.ptr)
            self.ptr = C.c_void_p()
