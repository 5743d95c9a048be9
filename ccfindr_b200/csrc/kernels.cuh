// CUDA kernels of the VB-NMF engine (sm_100a).  See DESIGN.md for the data layout and the
// roofline of each kernel.  Reference maths: src/vbnmf_update.cpp:33-90 (file:line cited per kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "special.cuh"

namespace vb {

constexpr int kBlock = 256;          // threads per CTA for all kernels here
constexpr int kWarpsPerBlock = kBlock / 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// deterministic block sum; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *sm /*kWarpsPerBlock*/) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kWarpsPerBlock; i++) t += sm[i];
    }
    return t;
}

// "last CTA finishes" reduction of per-CTA partial rows: part[b*W + c], b < gridDim.x.
// The last CTA to arrive sums every column c in fixed CTA order into out[c].  counter wraps
// back to 0 by itself (atomicInc), so the same counter serves every launch.
__device__ __forceinline__ void last_block_reduce(const double *part, int W, double *out,
                                                  unsigned *counter, double *sm) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicInc(counter, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int nb = gridDim.x;
    for (int c = 0; c < W; c++) {
        double a = 0.0;
        for (int b = threadIdx.x; b < nb; b += kBlock) a += __ldcg(part + (size_t)b * W + c);
        const double s = block_sum(a, sm);
        if (threadIdx.x == 0) out[c] = s;
    }
}

template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) {
    return __ldcs(p);
}

// load one panel row (RP entries, 16-byte aligned) into registers
template <int RP>
__device__ __forceinline__ void load_row_d(const double *__restrict__ base, int64_t row,
                                           double (&out)[RP]) {
    const double2 *p = reinterpret_cast<const double2 *>(base + row * RP);
#pragma unroll
    for (int k = 0; k < RP / 2; k++) {
        const double2 v = __ldg(p + k);
        out[2 * k] = v.x;
        out[2 * k + 1] = v.y;
    }
}

// ------------------------------------------------------------------------------------------
// Column sweep (cell-owner pass).  One warp per cell column j, lanes over its nonzeros.
//   p_ij = sum_k lw_ik lh_kj, q_ij = x_ij / p_ij               src/vbnmf_update.cpp:33-34
//   ShRaw[j][k] = sum_i lw_ik q_ij   (sh = lh o ShRaw)          src/vbnmf_update.cpp:36
//   col_xlogp[j] = sum_i x_ij log p_ij                          data term of :73-77
//   col_enth[j]  = sum_k log(lh_kj) lh_kj ShRaw[j][k]           B-term of :71-77 (entropy collapse)
// lh_j and the Sh accumulators stay in registers; lw rows are gathered; no atomics on outputs.
template <int RP, typename VT>
__global__ void __launch_bounds__(kBlock)
sweep_cols_kernel(int64_t m, int r, const int64_t *__restrict__ colptr,
                  const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                  const double *__restrict__ lw, const double *__restrict__ lh,
                  double *__restrict__ ShRaw, double *__restrict__ col_xlogp,
                  double *__restrict__ col_enth, unsigned long long *work_counter) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long jj = 0;
        if (lane == 0) jj = atomicAdd(work_counter, 1ull);
        jj = __shfl_sync(kFull, jj, 0);
        if (jj >= (unsigned long long)m) break;
        const int64_t j = (int64_t)jj;
        double lhj[RP], acc[RP];
        load_row_d<RP>(lh, j, lhj);
#pragma unroll
        for (int k = 0; k < RP; k++) acc[k] = 0.0;
        double xl = 0.0;
        const int64_t beg = __ldg(colptr + j), end = __ldg(colptr + j + 1);
        for (int64_t t = beg + lane; t < end; t += 32) {
            const int32_t i = ld_stream(rowidx + t);
            const double x = (double)ld_stream(val + t);
            double lwi[RP];
            load_row_d<RP>(lw, i, lwi);
            double p = 0.0;
#pragma unroll
            for (int k = 0; k < RP; k++) p = fma(lwi[k], lhj[k], p);
            const double q = x / p;
            xl = fma(x, log(p), xl);
#pragma unroll
            for (int k = 0; k < RP; k++) acc[k] = fma(lwi[k], q, acc[k]);
        }
        constexpr int NH = (RP + 31) / 32;
        double mine[NH], mylh[NH];
#pragma unroll
        for (int q = 0; q < NH; q++) { mine[q] = 0.0; mylh[q] = 1.0; }
#pragma unroll
        for (int k = 0; k < RP; k++) {
            const double s = warp_sum(acc[k]);
            if ((k & 31) == lane) { mine[k >> 5] = s; mylh[k >> 5] = lhj[k]; }
        }
        xl = warp_sum(xl);
        double e = 0.0;
#pragma unroll
        for (int q = 0; q < NH; q++) {
            const int kk = lane + 32 * q;
            if (kk < RP) ShRaw[j * RP + kk] = mine[q];
            if (kk < r) e += log(mylh[q]) * mylh[q] * mine[q];
        }
        e = warp_sum(e);
        if (lane == 0) {
            col_xlogp[j] = xl;
            col_enth[j] = e;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Row sweep (gene-owner pass) over the CSR mirror.  One warp per work item = (gene row, chunk of
// its nonzeros); recomputes p, q at the same lw, lh and accumulates
//   SwPart[item][k] = sum_{j in chunk} q_ij lh_kj   (sw = lw o SwRaw) src/vbnmf_update.cpp:35
// Partials of one row are summed in item order by combine_rows_kernel: no atomics.
template <int RP, typename VT>
__global__ void __launch_bounds__(kBlock)
sweep_rows_kernel(int64_t n_items, const int32_t *__restrict__ item_row,
                  const int64_t *__restrict__ item_beg, const int32_t *__restrict__ item_len,
                  const int32_t *__restrict__ colidx, const VT *__restrict__ val,
                  const double *__restrict__ lw, const double *__restrict__ lh,
                  double *__restrict__ SwPart, unsigned long long *work_counter) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(work_counter, 1ull);
        it = __shfl_sync(kFull, it, 0);
        if (it >= (unsigned long long)n_items) break;
        const int64_t i = __ldg(item_row + it);
        const int64_t beg = __ldg(item_beg + it), end = beg + __ldg(item_len + it);
        double lwi[RP], acc[RP];
        load_row_d<RP>(lw, i, lwi);
#pragma unroll
        for (int k = 0; k < RP; k++) acc[k] = 0.0;
        for (int64_t t = beg + lane; t < end; t += 32) {
            const int32_t j = ld_stream(colidx + t);
            const double x = (double)ld_stream(val + t);
            double lhj[RP];
            load_row_d<RP>(lh, j, lhj);
            double p = 0.0;
#pragma unroll
            for (int k = 0; k < RP; k++) p = fma(lwi[k], lhj[k], p);
            const double q = x / p;
#pragma unroll
            for (int k = 0; k < RP; k++) acc[k] = fma(lhj[k], q, acc[k]);
        }
        constexpr int NH = (RP + 31) / 32;
        double mine[NH];
#pragma unroll
        for (int q = 0; q < NH; q++) mine[q] = 0.0;
#pragma unroll
        for (int k = 0; k < RP; k++) {
            const double s = warp_sum(acc[k]);
            if ((k & 31) == lane) mine[k >> 5] = s;
        }
#pragma unroll
        for (int q = 0; q < NH; q++) {
            const int kk = lane + 32 * q;
            if (kk < RP) SwPart[(int64_t)it * RP + kk] = mine[q];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Posterior update of one side (W: rows = genes, H: rows = cells); identical algebra:
//   al = a + l o SRaw                                     src/vbnmf_update.cpp:38-39 / 48-49
//   be_k = a/b + osum_k   (osum = rowSums(eh_old) for W, colSums(ew_new) for H)   :40-43 / 50-53
//   e = al / be_k                                                                :44 / 54
//   l_new = max(exp(psi(al)) / be_k, fud)                                        :58-65
// and the reductions the bound and hyper_update need:
//   out[0..RP)   : sum over rows of e  (colSums(ew) / rowSums(eh))
//   out[RP+0]    : sum [ -(a/b) e + al (1 - log be_k) + lgamma(al) ]   (:84-86 / :88-89; the
//                  constant lga term is added on the host)
//   out[RP+1]    : sum log l_new                                       (R/bayesian.R:8-9)
//   out[RP+2]    : sum e                                               (R/bayesian.R:10-11)
// One thread per row.  al is kept for ew/dw (eh/dh) export.
template <int RP>
__global__ void __launch_bounds__(kBlock)
posterior_kernel(int64_t rows, int r, double a, double b, double fud,
                 const double *__restrict__ osum, const double *__restrict__ SRaw,
                 double *__restrict__ l, double *__restrict__ al_out, double *__restrict__ part,
                 double *__restrict__ out, unsigned *counter) {
    __shared__ double sm[kWarpsPerBlock];
    __shared__ double be[RP], lbe[RP];
    if (threadIdx.x < RP) {
        const double v = (threadIdx.x < r) ? a / b + osum[threadIdx.x] : 1.0;
        be[threadIdx.x] = v;
        lbe[threadIdx.x] = log(v);
    }
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    double es[RP];
    double prior = 0.0, sll = 0.0, se = 0.0;
#pragma unroll
    for (int k = 0; k < RP; k++) es[k] = 0.0;
    if (row < rows) {
        double lv[RP], sv[RP];
        load_row_d<RP>(l, row, lv);
        load_row_d<RP>(SRaw, row, sv);
        const double aob = a / b;
#pragma unroll
        for (int k = 0; k < RP; k++) {
            if (k < r) {
                const double al = a + lv[k] * sv[k];
                const double e = al / be[k];
                const double tmp = exp(vb_digamma(al)) / be[k];
                const double ln = tmp > fud ? tmp : fud;
                lv[k] = ln;
                sv[k] = al;
                es[k] = e;
                se += e;
                sll += log(ln);
                prior += -aob * e + al * (1.0 - lbe[k]) + lgamma(al);
            } else {
                lv[k] = 0.0;
                sv[k] = 0.0;
            }
        }
        double2 *lp = reinterpret_cast<double2 *>(l + row * RP);
        double2 *ap = reinterpret_cast<double2 *>(al_out + row * RP);
#pragma unroll
        for (int k = 0; k < RP / 2; k++) {
            lp[k] = make_double2(lv[2 * k], lv[2 * k + 1]);
            ap[k] = make_double2(sv[2 * k], sv[2 * k + 1]);
        }
    }
    constexpr int W = RP + 3;
    double *mypart = part + (size_t)blockIdx.x * W;
#pragma unroll
    for (int k = 0; k < RP; k++) {
        const double s = block_sum(es[k], sm);
        if (threadIdx.x == 0) mypart[k] = s;
    }
    prior = block_sum(prior, sm);
    sll = block_sum(sll, sm);
    se = block_sum(se, sm);
    if (threadIdx.x == 0) {
        mypart[RP + 0] = prior;
        mypart[RP + 1] = sll;
        mypart[RP + 2] = se;
    }
    last_block_reduce(part, W, out, counter, sm);
}

// ---- maximum-likelihood multiplicative updates (R/factorize.R:8-15 for h, :17-24 for w) ----
//   v_new = max(v o SRaw / osum_k, eps);  out[0..RP) = sum over rows of v_new
template <int RP>
__global__ void __launch_bounds__(kBlock)
ml_update_kernel(int64_t rows, int r, double eps, const double *__restrict__ osum,
                 const double *__restrict__ SRaw, double *__restrict__ v,
                 double *__restrict__ part, double *__restrict__ out, unsigned *counter) {
    __shared__ double sm[kWarpsPerBlock];
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    double es[RP];
#pragma unroll
    for (int k = 0; k < RP; k++) es[k] = 0.0;
    if (row < rows) {
        double lv[RP], sv[RP];
        load_row_d<RP>(v, row, lv);
        load_row_d<RP>(SRaw, row, sv);
#pragma unroll
        for (int k = 0; k < RP; k++) {
            if (k < r) {
                double x = lv[k] * sv[k] / osum[k];
                if (x < eps) x = eps;
                lv[k] = x;
                es[k] = x;
            } else {
                lv[k] = 0.0;
            }
        }
        double2 *lp = reinterpret_cast<double2 *>(v + row * RP);
#pragma unroll
        for (int k = 0; k < RP / 2; k++) lp[k] = make_double2(lv[2 * k], lv[2 * k + 1]);
    }
    double *mypart = part + (size_t)blockIdx.x * RP;
#pragma unroll
    for (int k = 0; k < RP; k++) {
        const double s = block_sum(es[k], sm);
        if (threadIdx.x == 0) mypart[k] = s;
    }
    last_block_reduce(part, RP, out, counter, sm);
}

// column sums of a rows x RP panel: out[k] = sum_row v[row][k]
template <int RP>
__global__ void __launch_bounds__(kBlock)
panel_colsum_kernel(int64_t rows, const double *__restrict__ v, double *__restrict__ part,
                    double *__restrict__ out, unsigned *counter) {
    __shared__ double sm[kWarpsPerBlock];
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    double lv[RP];
#pragma unroll
    for (int k = 0; k < RP; k++) lv[k] = 0.0;
    if (row < rows) load_row_d<RP>(v, row, lv);
    double *mypart = part + (size_t)blockIdx.x * RP;
#pragma unroll
    for (int k = 0; k < RP; k++) {
        const double s = block_sum(lv[k], sm);
        if (threadIdx.x == 0) mypart[k] = s;
    }
    last_block_reduce(part, RP, out, counter, sm);
}

}  // namespace vb
