// Kernels that do not depend on the padded rank; included only by vbnmf.cu (single definition).
#pragma once
#include "kernels.cuh"

namespace vb {

// SwRaw[i][k] = sum over the items of row i, in item order
__global__ void __launch_bounds__(kBlock)
combine_rows_kernel(int64_t n, int RP, const int64_t *__restrict__ row_item_ptr,
                    const double *__restrict__ SwPart, double *__restrict__ SwRaw) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e >= n * RP) return;
    const int64_t i = e / RP;
    const int k = (int)(e - i * RP);
    double a = 0.0;
    for (int64_t it = row_item_ptr[i]; it < row_item_ptr[i + 1]; it++) a += SwPart[it * RP + k];
    SwRaw[e] = a;
}

// sum the per-column scalars of the column sweep: out[0] = sum xlogp, out[1] = sum enth
__global__ void __launch_bounds__(kBlock)
reduce_cols_kernel(int64_t m, const double *__restrict__ col_xlogp,
                   const double *__restrict__ col_enth, double *__restrict__ part,
                   double *__restrict__ out, unsigned *counter) {
    __shared__ double sm[kWarpsPerBlock];
    double a = 0.0, b = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x; j < m;
         j += (int64_t)gridDim.x * kBlock) {
        a += col_xlogp[j];
        b += col_enth[j];
    }
    a = block_sum(a, sm);
    b = block_sum(b, sm);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 2 + 0] = a;
        part[blockIdx.x * 2 + 1] = b;
    }
    last_block_reduce(part, 2, out, counter, sm);
}

// out[0] = sum_ik log(lw_ik) lw_ik SwRaw_ik   (A-term of src/vbnmf_update.cpp:69-77)
__global__ void __launch_bounds__(kBlock)
entropy_w_kernel(int64_t n, int RP, int r, const double *__restrict__ lw,
                 const double *__restrict__ SwRaw, double *__restrict__ part,
                 double *__restrict__ out, unsigned *counter) {
    __shared__ double sm[kWarpsPerBlock];
    double a = 0.0;
    const int64_t tot = n * RP;
    for (int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x; e < tot;
         e += (int64_t)gridDim.x * kBlock) {
        const int k = (int)(e % RP);
        if (k < r) {
            const double v = lw[e];
            a += log(v) * v * SwRaw[e];
        }
    }
    a = block_sum(a, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
    last_block_reduce(part, 1, out, counter, sm);
}

// ---- one-time helpers ----------------------------------------------------------------------
// sum over nonzeros of lgamma(x+1) (src/vbnmf_update.cpp:80-81; zeros contribute 0) and of
// -x log x + x (R/factorize.R:45-46); out[0], out[1]
template <typename VT>
__global__ void __launch_bounds__(kBlock)
count_constants_kernel(int64_t nnz, const VT *__restrict__ val, double *__restrict__ part,
                       double *__restrict__ out, unsigned *counter) {
    __shared__ double sm[kWarpsPerBlock];
    double a = 0.0, b = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < nnz;
         t += (int64_t)gridDim.x * kBlock) {
        const double x = (double)val[t];
        a += lgamma(x + 1.0);
        if (x > 0) b += -x * log(x) + x;
    }
    a = block_sum(a, sm);
    b = block_sum(b, sm);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 2 + 0] = a;
        part[blockIdx.x * 2 + 1] = b;
    }
    last_block_reduce(part, 2, out, counter, sm);
}

// expand CSC column pointers into a per-nonzero column index; count nonzeros per row
__global__ void __launch_bounds__(kBlock)
expand_cols_kernel(int64_t m, const int64_t *__restrict__ colptr,
                   const int32_t *__restrict__ rowidx, int32_t *__restrict__ colof,
                   unsigned long long *__restrict__ row_count) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kBlock) >> 5;
    for (int64_t j = warp; j < m; j += nwarps) {
        const int64_t beg = colptr[j], end = colptr[j + 1];
        for (int64_t t = beg + lane; t < end; t += 32) {
            colof[t] = (int32_t)j;
            atomicAdd(row_count + rowidx[t], 1ull);
        }
    }
}

// gather the CSR mirror through the stable row sort permutation
template <typename VT>
__global__ void __launch_bounds__(kBlock)
gather_csr_kernel(int64_t nnz, const uint32_t *__restrict__ perm, const int32_t *__restrict__ colof,
                  const VT *__restrict__ val, int32_t *__restrict__ colidx_out,
                  VT *__restrict__ val_out) {
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < nnz;
         t += (int64_t)gridDim.x * kBlock) {
        const uint32_t s = perm[t];
        colidx_out[t] = colof[s];
        val_out[t] = val[s];
    }
}

__global__ void __launch_bounds__(kBlock)
iota_kernel(int64_t nnz, uint32_t *__restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < nnz;
         t += (int64_t)gridDim.x * kBlock)
        out[t] = (uint32_t)t;
}

// cid[j] = 1 + index of the first maximum over k of alh[j][k] / beh[k]   (R/utils.R:906)
__global__ void __launch_bounds__(kBlock)
cluster_id_kernel(int64_t m, int RP, int r, const double *__restrict__ alh,
                  const double *__restrict__ beh, int32_t *__restrict__ cid) {
    const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (j >= m) return;
    int best = 0;
    double bv = alh[j * RP] / beh[0];
    for (int k = 1; k < r; k++) {
        const double v = alh[j * RP + k] / beh[k];
        if (v > bv) { bv = v; best = k; }
    }
    cid[j] = best + 1;
}

}  // namespace vb
