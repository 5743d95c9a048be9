/* libvbnmf — C ABI of the B200-native variational-Bayes Poisson-NMF engine.
 *
 * This is the drop-in boundary for ONE path of the R package ccfindR: the per-iteration VB update
 * and the loop that drives it.  Citations are file:line in the reference tree.
 *
 *   reference interface replaced                                     entry point here
 *   ---------------------------------------------------------------  -------------------------
 *   .Call(`_ccfindR_vbnmf_update`, X, wh, hyper, fudge)               vbnmf_create + vbnmf_set_state
 *     R/RcppExports.R:4-6, src/RcppExports.cpp:11-22                  + vbnmf_step + vbnmf_get_state
 *   vbnmf_update(): one VB iteration, src/vbnmf_update.cpp:16-102     vbnmf_step
 *   return list w,h,lw,lh,ew,eh,lkh,dw,dh, src/vbnmf_update.cpp:92-100 vbnmf_get_state (+ *lkh)
 *   for(it in seq_len(Itmax)) loop of vb_iterate, R/bayesian.R:336-352 vbnmf_run
 *   hyper_update(), R/bayesian.R:2-53                                  (inside vbnmf_run)
 *   uniform-column test, R/bayesian.R:368-369                          vbnmf_uniform_columns
 *   cluster_id(): apply(h,2,which.max), R/utils.R:903-909              vbnmf_cluster_id
 *   nmf_updateR() + likelihood() loop, R/factorize.R:189-212           mlnmf_run, mlnmf_run2 (both criteria)
 *   vb_init(): 'random', 'svd2', R/bayesian.R:111-115,150-159          vbnmf_init_random, vbnmf_init_svd2
 *   read_10x(): readMM + as(., 'dgCMatrix'), R/utils.R:34               vbnmf_create_from_mtx
 *   Rmpi task farm over restarts, R/bayesian.R:263                     one handle per process/GPU;
 *                                                                      cells sharded: vbnmf_attach_comm
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; vbnmf_last_error() gives the text.
 *     No C++ exception crosses this boundary.
 *   - all matrices are column-major doubles exactly as R stores them: lw/ew/dw are n x r,
 *     lh/eh/dh are r x m.  hyper = {aw, bw, ah, bh}.
 *   - the count matrix is CSC as in Matrix::dgCMatrix: colptr = @p (length m+1), rowidx = @i
 *     (0-based), values = @x.  Column pointers may be 32- or 64-bit; on the device they are 64-bit.
 *   - a handle owns one CUDA device and is not thread-safe.  There is no CPU fallback: creating a
 *     handle without a usable CUDA device fails.
 *   - when cells are sharded over several GPUs (one process and one handle per GPU), m is the
 *     LOCAL number of cells, lh/eh/dh are the local columns, lw/ew/dw are replicated, and every
 *     rank must make the same sequence of calls (set_state/step/run are collective).
 */
#ifndef VBNMF_H
#define VBNMF_H
#include <stdint.h>

#if defined(__GNUC__)
#define VBNMF_API __attribute__((visibility("default")))
#else
#define VBNMF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vbnmf_handle vbnmf_handle;

/* mirrors the `bundle` of vb_factorize (R/bayesian.R:252-259) for the fields the loop reads */
typedef struct {
    int itmax;            /* Itmax */
    double tol;           /* Tol */
    int hyper_update[4];  /* hyper.update: aw, bw, ah, bh */
    int n0;               /* hyper.update.n0 */
    int dn;               /* hyper.update.dn */
    double fudge;         /* fudge (R passes .Machine$double.eps when NULL, R/bayesian.R:238) */
} vbnmf_cfg;

enum { VBNMF_FP64 = 0, VBNMF_FP32_STORAGE = 1 };
enum { VBNMF_STOP_ITMAX = 0, VBNMF_STOP_CONVERGED = 1, VBNMF_STOP_NAN = 2 };
enum {
    VBNMF_OK = 0,
    VBNMF_ERR_ARG = 1,
    VBNMF_ERR_HYPER = 2, /* 'Hyper-parameter update failed to converge', R/bayesian.R:43 */
    VBNMF_ERR_CUDA = 3,
    VBNMF_ERR_STATE = 4,
    VBNMF_ERR_NCCL = 5,
    VBNMF_ERR_EMPTY = 6  /* 'Input matrix contains empty rows/columns', R/bayesian.R:244-247 */
};

/* Upload a CSC count matrix from HOST memory and build the device-side layouts.
 * Exactly one of colptr32 / colptr64 is non-NULL.  device = CUDA ordinal. */
VBNMF_API int vbnmf_create(vbnmf_handle **out, int64_t n, int64_t m, int64_t nnz, const int32_t *colptr32,
                 const int64_t *colptr64, const int32_t *rowidx, const double *values, int device);

/* Same, from arrays already resident on `device` (int64 colptr, int32 rowidx, fp32 counts).
 * The arrays are borrowed: they must outlive the handle.  They are read on the handle's own
 * stream, starting inside this call: the work that produces them must have COMPLETED (synchronise
 * the producing stream first) -- the library cannot know that stream. */
VBNMF_API int vbnmf_create_from_device(vbnmf_handle **out, int64_t n, int64_t m, int64_t nnz,
                             const int64_t *d_colptr, const int32_t *d_rowidx, const float *d_values,
                             int device);

/* Host threads that stage uploads (copy / fp64 -> fp32 conversion into pinned buffers).  Default
 * min(16, hardware threads) or VBNMF_HOST_THREADS; with one process per GPU pass cores / processes.
 * Process-wide.  (The reference holds X in R memory; this replaces as.matrix(), R/bayesian.R:339.) */
VBNMF_API int vbnmf_set_host_threads(int nthreads);

/* Return the device memory the engine's stream-ordered pool keeps cached to the driver. */
VBNMF_API int vbnmf_trim_pool(int device);

/* Read a MatrixMarket coordinate file (the `matrix.mtx` of a 10x directory, R/utils.R:28-54:
 * Matrix::readMM + as(., 'dgCMatrix')) straight into a device-resident handle: the host reads the
 * bytes and the header, the text is parsed, sorted into CSC and validated on the GPU.  Supports
 * `matrix coordinate real|integer general`.  dims (may be NULL) receives n, m, nnz. */
VBNMF_API int vbnmf_create_from_mtx(vbnmf_handle **out, const char *path, int device, int64_t dims[3]);

/* The CSC arrays of a handle copied to the host: colptr (m+1), rowidx (nnz), values (nnz); any may
 * be NULL.  (After vbnmf_create_from_mtx the front end needs them for the row/column filters.) */
VBNMF_API int vbnmf_get_csc(vbnmf_handle *h, int64_t *colptr, int32_t *rowidx, double *values);

VBNMF_API void vbnmf_destroy(vbnmf_handle *h);
VBNMF_API const char *vbnmf_last_error(const vbnmf_handle *h); /* h may be NULL: error of the last create */

/* VBNMF_FP64 (default): all arithmetic IEEE double, as the reference.
 * VBNMF_FP32_STORAGE: lw/lh panels held in fp32, per-nonzero arithmetic fp32, all sums fp64. */
VBNMF_API int vbnmf_set_precision(vbnmf_handle *h, int precision);

/* Run the engine's kernels on a caller-provided CUDA stream (cudaStream_t cast to void*). */
VBNMF_API int vbnmf_set_stream(vbnmf_handle *h, void *cuda_stream);

/* Multi-GPU: an NCCL communicator of `nranks` processes (one per GPU) that hold disjoint cell
 * ranges of one matrix.  The communicator belongs to the process, not to a handle: create it once
 * (uid = 128-byte ncclUniqueId from vbnmf_nccl_unique_id() on rank 0, distributed by the host, e.g.
 * over MPI where the reference uses Rmpi, R/bayesian.R:263) and attach it to each handle. */
typedef struct vbnmf_comm vbnmf_comm;
VBNMF_API int vbnmf_nccl_unique_id(void *uid128);
VBNMF_API int vbnmf_comm_create(vbnmf_comm **out, int nranks, int rank, const void *uid128, int device);
VBNMF_API void vbnmf_comm_destroy(vbnmf_comm *c);
VBNMF_API int vbnmf_attach_comm(vbnmf_handle *h, vbnmf_comm *c);

/* Load the state list `wh` (R/bayesian.R:170: lw, lh, ew, eh).  ew may be NULL (it is overwritten
 * before use, src/vbnmf_update.cpp:44); eh NULL means eh = lh (vb_init). */
VBNMF_API int vbnmf_set_state(vbnmf_handle *h, int r, const double *lw, const double *lh, const double *ew,
                    const double *eh);

/* vb_init(initializer = 'random') (R/bayesian.R:111-115,170) drawn on the device instead of
 * vbnmf_set_state from host matrices: w_ik ~ Gamma(shape aw, scale bw/aw), h_kj ~ Gamma(shape ah,
 * scale bh/ah), lw = ew = w, lh = eh = h.  R's RNG stream cannot be reproduced; the draw is defined
 * by a counter RNG keyed by (seed, gene i or GLOBAL cell index cell_offset + j, k) (csrc/
 * kernels_common.cuh, restated in ccfindr_b200/synth.py:device_random_init_reference), so it does
 * not depend on the device layout or on how the cells are sharded.  hyper = {aw, bw, ah, bh}. */
VBNMF_API int vbnmf_init_random(vbnmf_handle *h, int r, const double hyper[4], uint64_t seed,
                                int64_t cell_offset);

/* vb_init(initializer = 'svd2') (R/bayesian.R:150-159) on the device: w = |u| / scale,
 * h = |diag(d) v^T| * scale, scale = bh / mean(h), from a rank-r truncated SVD of the count matrix
 * computed where it lives (randomized subspace iteration with oversampling 20, run until the r
 * leading Ritz values stop moving, in place of irlba; k x k factorizations on the host).  Singular
 * vectors are defined up to sign, which abs() removes; they match an exact truncated SVD to ~1e-7.
 * seed keys the Gaussian test matrix by GLOBAL cell index (cell_offset + j): the result does not
 * depend on the sharding.  hyper = {aw, bw, ah, bh} (only bh is used, as in the reference). */
VBNMF_API int vbnmf_init_svd2(vbnmf_handle *h, int r, const double hyper[4], uint64_t seed,
                              int64_t cell_offset);

/* One call of vbnmf_update (src/vbnmf_update.cpp:16-102): state <- update(state); *lkh = bound. */
VBNMF_API int vbnmf_step(vbnmf_handle *h, const double hyper[4], double fudge, double *lkh);

/* The loop of vb_iterate for one rank (R/bayesian.R:336-352) on the loaded state.
 * hyper: in = initial {aw,bw,ah,bh}, out = final.  lkh_trace (itmax doubles, may be NULL):
 * lkh of every executed iteration.  hyper_trace (4*itmax, may be NULL): hyper after each iteration.
 * *lml = lk0 as stored at R/bayesian.R:379 (NOT updated by the iteration that breaks). */
VBNMF_API int vbnmf_run(vbnmf_handle *h, const vbnmf_cfg *cfg, double hyper[4], double *lkh_trace,
              double *hyper_trace, int *niter, double *lml, int *stop_reason);

/* Copy the state out; any pointer may be NULL.  dw, dh are variances (the R driver stores their
 * square roots, R/bayesian.R:382-383). */
VBNMF_API int vbnmf_get_state(vbnmf_handle *h, double *lw, double *lh, double *ew, double *eh, double *dw,
                    double *dh);

/* inputs of hyper_update for the current state: mean(log lw), mean(log lh), mean(ew), mean(eh) */
VBNMF_API int vbnmf_get_means(vbnmf_handle *h, double means[4]);

/* cid[j] = 1-based index of the first maximum of column j of eh (local cells) */
VBNMF_API int vbnmf_cluster_id(vbnmf_handle *h, int32_t *cid);

/* flags[k] = 1 when column k of ew has |max - min| < tol (R/bayesian.R:368-369) */
VBNMF_API int vbnmf_uniform_columns(vbnmf_handle *h, double tol, int32_t *flags);

/* Maximum-likelihood path: the it-loop of factorize() with criterion='likelihood'
 * (R/factorize.R:189-212) from initial w0 (n x r), h0 (r x m).  lik_trace may be NULL. */
VBNMF_API int mlnmf_run(vbnmf_handle *h, int r, const double *w0, const double *h0, int itmax, double tol,
              double *w, double *h_out, double *lik_trace, int *niter);

/* Same loop with the stopping rule of factorize() selectable (R/factorize.R:194-212):
 * criterion = VBNMF_ML_LIKELIHOOD: |lkold - lk0| < tol |lkold| (:207, what mlnmf_run does);
 * criterion = VBNMF_ML_CONNECTIVITY: stop when the connectivity matrix of the cells
 * (outer(cid, cid, '=='), cid = which.max of each column of h, :51-60) has not changed for
 * ncnn_step consecutive iterations (:195-203).  The m(m-1)/2 connectivity vector is never formed:
 * sum(cnn != cnn0) is taken from the r x r contingency table of consecutive labelings.
 * nchange_trace (itmax doubles, may be NULL): that count per iteration (npair at iteration 1). */
enum { VBNMF_ML_LIKELIHOOD = 0, VBNMF_ML_CONNECTIVITY = 1 };
VBNMF_API int mlnmf_run2(vbnmf_handle *h, int r, const double *w0, const double *h0, int itmax,
               double tol, int criterion, int ncnn_step, double *w, double *h_out,
               double *lik_trace, double *nchange_trace, int *niter);

/* Measurement hooks (bench.py).  Runs `iters` steady-state VB iterations (posterior update +
 * nonzero sweep [+ all-reduce]) with fixed hypers and reports CUDA-event times in ms:
 * ms[0] = whole timed region, ms[1] = sum over the column-sweep kernel, ms[2] = row-sweep kernel,
 * ms[3] = everything else.  launches = kernels launched inside the timed region.
 * hyper_on = 1: hyper_update (R/bayesian.R:2-53) after every iteration as in the reference loop past
 * hyper.update.n0, hyper[] in/out; 0: hyper-parameters held fixed. */
VBNMF_API int vbnmf_bench_iterations(vbnmf_handle *h, double hyper[4], double fudge, int iters,
                           int hyper_on, double ms[4], int64_t *launches, double *lkh_last);

/* shape / layout facts for the host side */
VBNMF_API int vbnmf_info(const vbnmf_handle *h, int64_t info[8]); /* n, m, nnz, r, rs, precision, nranks, m_global */
/* facts about the tiled device layout in use (valid after vbnmf_set_state / mlnmf_run):
 * info[0] storage format of the nonzeros: 0 = {int32 row, float count} (8 bytes), 1 = int32 row +
 *         double count (12 bytes; counts not exact in fp32), 2 = packed {count << 16 | row}
 *         (4 bytes; integer counts below 2^16);
 * info[1] tile rows T, info[2] gene slabs, info[3] cell slabs,
 * info[4], info[5] entries stored for the cell-owner / gene-owner pass (nonzeros + schedule holes),
 * info[6] nonzeros per 8-lane group step, info[7] bytes of both passes' entry and pointer arrays */
VBNMF_API int vbnmf_layout_info(const vbnmf_handle *h, int64_t info[8]);

#ifdef __cplusplus
}
#endif
#endif
