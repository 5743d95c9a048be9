// Kernels that do not depend on the padded rank; included only by vbnmf.cu (single definition).
#pragma once
#include "kernels.cuh"

namespace vb {

// ---- one-time helpers ----------------------------------------------------------------------
// sum over nonzeros of lgamma(x+1) (src/vbnmf_update.cpp:80-81; zeros contribute 0) and of
// -x log x + x (R/factorize.R:45-46); out[0], out[1].  out[2] = number of values that are not
// integers in [0, 65535] (0 -> the packed 16-bit layout of the sweep applies).
template <typename VT>
__global__ void __launch_bounds__(kBlock)
count_constants_kernel(int64_t nnz, const VT *__restrict__ val, double *__restrict__ part,
                       double *__restrict__ out, unsigned *counter) {
    __shared__ double sm[kWarpsPerBlock];
    double a = 0.0, b = 0.0, c = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < nnz;
         t += (int64_t)gridDim.x * kBlock) {
        const double x = (double)val[t];
        a += lgamma(x + 1.0);
        if (x > 0) b += -x * log(x) + x;
        if (!(x >= 0.0 && x <= 65535.0 && x == floor(x))) c += 1.0;
    }
    a = block_sum(a, sm);
    b = block_sum(b, sm);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 3 + 0] = a;
        part[blockIdx.x * 3 + 1] = b;
        part[blockIdx.x * 3 + 2] = c;
    }
    last_block_reduce(part, 3, out, counter, sm);
}

// expand CSC column pointers into a per-nonzero column index; count nonzeros per row and column
__global__ void __launch_bounds__(kBlock)
expand_cols_kernel(int64_t m, const int64_t *__restrict__ colptr,
                   const int32_t *__restrict__ rowidx, int32_t *__restrict__ colof,
                   unsigned long long *__restrict__ row_count,
                   unsigned long long *__restrict__ col_count) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kBlock) >> 5;
    for (int64_t j = warp; j < m; j += nwarps) {
        const int64_t beg = colptr[j], end = colptr[j + 1];
        if (lane == 0) col_count[j] = (unsigned long long)(end - beg);
        for (int64_t t = beg + lane; t < end; t += 32) {
            colof[t] = (int32_t)j;
            atomicAdd(row_count + rowidx[t], 1ull);
        }
    }
}

// sort key of every nonzero for one pass of the tiled layout:
//   key = slab(tile side) * NO + owner(device row of the owner side),  payload = nonzero index
__global__ void __launch_bounds__(kBlock)
make_keys_kernel(int64_t nnz, const int32_t *__restrict__ rowidx, const int32_t *__restrict__ colof,
                 const int32_t *__restrict__ gene_dev, const int32_t *__restrict__ cell_dev, int T,
                 int64_t NO, bool cols_pass, uint32_t *__restrict__ key,
                 uint32_t *__restrict__ payload) {
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < nnz;
         t += (int64_t)gridDim.x * kBlock) {
        const int64_t gd = gene_dev[rowidx[t]], cd = cell_dev[colof[t]];
        const int64_t k = cols_pass ? (gd / T) * NO + cd : (cd / T) * NO + gd;
        key[t] = (uint32_t)k;
        payload[t] = (uint32_t)t;
    }
}

// ptr[e] = first sorted position whose key is >= e, e in [0, E]
__global__ void __launch_bounds__(kBlock)
segment_ptr_kernel(int64_t E, int64_t nnz, const uint32_t *__restrict__ sorted_key,
                   int64_t *__restrict__ ptr) {
    for (int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x; e <= E;
         e += (int64_t)gridDim.x * kBlock) {
        int64_t lo = 0, hi = nnz;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)sorted_key[mid] < e) lo = mid + 1; else hi = mid;
        }
        ptr[e] = lo;
    }
}

// Gather one pass of the tiled layout through the sort permutation and order each segment for
// conflict-free shared-memory gathers: the nonzeros of a segment are bucketed by (tile row mod 8)
// and emitted round-robin, one from every non-empty bucket per round in the residue order
// 0,2,4,6,1,3,5,7: 8 consecutive nonzeros touch tile rows that differ mod 8, and 4 consecutive
// ones rows of equal parity (what the 2-lanes-per-nonzero variant of the sweep needs), until the
// smaller buckets run dry.
// One thread per segment; the order inside a bucket is the sorted (stable) order.
template <typename VT>
__global__ void __launch_bounds__(kBlock)
build_segments_kernel(int64_t E, const int64_t *__restrict__ ptr,
                      const uint32_t *__restrict__ perm, const int32_t *__restrict__ rowidx,
                      const int32_t *__restrict__ colof, const int32_t *__restrict__ gene_dev,
                      const int32_t *__restrict__ cell_dev, const VT *__restrict__ val, int T,
                      bool cols_pass, int32_t *__restrict__ idx_out, VT *__restrict__ val_out,
                      int2 *__restrict__ ent_out) {
    for (int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x; e < E;
         e += (int64_t)gridDim.x * kBlock) {
        const int64_t beg = ptr[e], end = ptr[e + 1];
        if (beg == end) continue;
        int cnt[8];
#pragma unroll
        for (int b = 0; b < 8; b++) cnt[b] = 0;
        for (int64_t t = beg; t < end; t++) {
            const uint32_t s = perm[t];
            const int32_t d = cols_pass ? gene_dev[rowidx[s]] : cell_dev[colof[s]];
            const int rr = (d % T) & 7;
            cnt[(rr >> 1) | ((rr & 1) << 2)]++;  // bucket order 0,2,4,6,1,3,5,7
        }
        int seen[8];
#pragma unroll
        for (int b = 0; b < 8; b++) seen[b] = 0;
        for (int64_t t = beg; t < end; t++) {
            const uint32_t s = perm[t];
            const int32_t d = cols_pass ? gene_dev[rowidx[s]] : cell_dev[colof[s]];
            const int local = d % T, rr = local & 7, b = (rr >> 1) | ((rr & 1) << 2);
            const int round = seen[b]++;
            // position = elements of all buckets in earlier rounds + earlier buckets in this round
            int64_t pos = 0;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                pos += min(cnt[c], round);
                if (c < b && cnt[c] > round) pos++;
            }
            if (ent_out) {  // float counts: packed {tile row, count bits}
                ent_out[beg + pos] = make_int2(local, __float_as_int((float)val[s]));
            } else {
                idx_out[beg + pos] = local;
                val_out[beg + pos] = val[s];
            }
        }
    }
}

// ---- packed-16 layout (sweep_p16_kernel) ----------------------------------------------------
// quads (4 entries = 16 bytes) per segment, segments padded to a whole number of quads
__global__ void __launch_bounds__(kBlock)
quad_len_kernel(int64_t E, const int64_t *__restrict__ ptr, uint32_t *__restrict__ len4) {
    for (int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x; e <= E;
         e += (int64_t)gridDim.x * kBlock)
        len4[e] = e < E ? (uint32_t)((ptr[e + 1] - ptr[e] + 3) >> 2) : 0u;
}

// position of item p (round-robin residue order) of a segment of L4 entries (multiple of 4): in
// blocks of B = 4*NPG entries, item p of a block with q quads -> quad p mod q, word p div q
__device__ __forceinline__ int64_t p16_position(int64_t p, int64_t L4, int B) {
    const int64_t blk = p / B;
    const int pin = (int)(p - blk * B);
    const int64_t rem = L4 - blk * B;
    const int q = (int)(rem < B ? rem : B) >> 2;
    return blk * B + 4 * (pin % q) + pin / q;
}

// One thread per segment, same bucketed round-robin order as build_segments_kernel, entries
// written as {count << 16 | tile row} at their p16_position; the <= 3 padding items are zero words.
template <typename VT>
__global__ void __launch_bounds__(kBlock)
build_segments_p16_kernel(int64_t E, const int64_t *__restrict__ ptr,
                          const uint32_t *__restrict__ ptr4, const uint32_t *__restrict__ perm,
                          const int32_t *__restrict__ rowidx, const int32_t *__restrict__ colof,
                          const int32_t *__restrict__ gene_dev,
                          const int32_t *__restrict__ cell_dev, const VT *__restrict__ val, int T,
                          bool cols_pass, int B, uint32_t *__restrict__ ent_out) {
    for (int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x; e < E;
         e += (int64_t)gridDim.x * kBlock) {
        const int64_t beg = ptr[e], end = ptr[e + 1];
        if (beg == end) continue;
        const int64_t len = end - beg, L4 = (len + 3) & ~(int64_t)3;
        uint32_t *dst = ent_out + (int64_t)ptr4[e] * 4;
        int cnt[8];
#pragma unroll
        for (int b = 0; b < 8; b++) cnt[b] = 0;
        for (int64_t t = beg; t < end; t++) {
            const uint32_t s = perm[t];
            const int32_t d = cols_pass ? gene_dev[rowidx[s]] : cell_dev[colof[s]];
            const int rr = (d % T) & 7;
            cnt[(rr >> 1) | ((rr & 1) << 2)]++;  // bucket order 0,2,4,6,1,3,5,7
        }
        int seen[8];
#pragma unroll
        for (int b = 0; b < 8; b++) seen[b] = 0;
        for (int64_t t = beg; t < end; t++) {
            const uint32_t s = perm[t];
            const int32_t d = cols_pass ? gene_dev[rowidx[s]] : cell_dev[colof[s]];
            const int local = d % T, rr = local & 7, b = (rr >> 1) | ((rr & 1) << 2);
            const int round = seen[b]++;
            int64_t pos = 0;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                pos += min(cnt[c], round);
                if (c < b && cnt[c] > round) pos++;
            }
            const uint32_t count = (uint32_t)val[s];
            dst[p16_position(pos, L4, B)] = (count << 16) | (uint32_t)local;
        }
        for (int64_t p = len; p < L4; p++) dst[p16_position(p, L4, B)] = 0u;
    }
}

// split[b] = first segment whose quad offset is >= b * total / nparts; split[nparts] = E
__global__ void split_p16_kernel(int nparts, int64_t E, const uint32_t *__restrict__ ptr4,
                                 int64_t *__restrict__ split) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nparts) return;
    if (b == nparts) { split[b] = E; return; }
    const int64_t target = (int64_t)((double)ptr4[E] * b / nparts);
    int64_t lo = 0, hi = E;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)ptr4[mid] < target) lo = mid + 1; else hi = mid;
    }
    split[b] = lo;
}

// split[b] = first segment whose start offset is >= b * nnz / nparts; split[nparts] = E
__global__ void split_kernel(int nparts, int64_t E, int64_t nnz, const int64_t *__restrict__ ptr,
                             int64_t *__restrict__ split) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nparts) return;
    if (b == nparts) { split[b] = E; return; }
    const int64_t target = (int64_t)((double)nnz * b / nparts);
    int64_t lo = 0, hi = E;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (ptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    split[b] = lo;
}

// One thread: the host half of an iteration moved onto the device.  Assembles the lower bound
// (src/vbnmf_update.cpp:67-90 in its nonzero-only form) from the reduced scalars, then applies the
// loop logic of vb_iterate (R/bayesian.R:342-348): hyper update, NaN stop, convergence test, lk0.
struct ControlArgs {
    double *ctl;
    const double *scal;   // [ewsum rs | wprior, sum log lw, sum ew]
    const double *tail;   // [ehsum rs | hprior, sum log lh, sum eh | enth, xlogp | entw]
    double *trace, *htrace;
    double n, m_global, lgx, tol;
    int r, rs, itmax, n0, dn;
    int flags[4];
};

__global__ void control_kernel(const ControlArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double *c = a.ctl;
    if (c[kCtlDone] != 0.0) return;
    const double aw = c[kCtlHyper + 0], bw = c[kCtlHyper + 1], ah = c[kCtlHyper + 2],
                 bh = c[kCtlHyper + 3];
    const double *ws = a.scal, *hs = a.tail;
    const int r = a.r, rs = a.rs;
    double U = 0.0;
    for (int k = 0; k < r; k++) {
        c[kCtlBew + k] = aw / bw + c[kCtlEhsum + k];   // what this iteration's W update used, :42-43
        c[kCtlBeh + k] = ah / bh + ws[k];              // and its H update, :52-53
        c[kCtlEhsum + k] = hs[k];                      // rowSums(eh_new) for the next iteration
        U -= ws[k] * hs[k];                            // -sum(ew.eh), :78
    }
    U -= hs[rs + 5] + hs[rs + 3] - hs[rs + 4];         // -x((A+B)/wth - log wth), :74-78
    U -= a.lgx;                                        // -lgamma(x+1), :81
    const double nr = a.n * r, mr = a.m_global * r;
    U += ws[rs + 0] + nr * (-lgamma(aw) + aw * log(aw / bw));   // :82-86
    U += hs[rs + 0] + mr * (-lgamma(ah) + ah * log(ah / bh));   // :87-89
    const double lkh = U / (a.n * a.m_global);                  // :90 (in double)
    for (int q = 0; q < 3; q++) { c[kCtlAcc + q] = ws[rs + q]; c[kCtlAcc + 3 + q] = hs[rs + q]; }
    const int it = (int)c[kCtlIt] + 1;
    c[kCtlIt] = it;
    c[kCtlLkh] = lkh;
    double hyper[4] = {aw, bw, ah, bh};
    if (it > a.n0 && it % a.dn == 0) {                 // R/bayesian.R:342-344
        const double mn[4] = {ws[rs + 1] / nr, hs[rs + 1] / mr, ws[rs + 2] / nr, hs[rs + 2] / mr};
        if (vb_hyper_update(a.flags, mn, hyper, 100, 1e-3)) {
            c[kCtlHyperErr] = 1.0;
            c[kCtlDone] = 1.0;
            return;
        }
        for (int q = 0; q < 4; q++) c[kCtlHyper + q] = hyper[q];
    }
    if (a.trace) a.trace[it - 1] = lkh;
    if (a.htrace)
        for (int q = 0; q < 4; q++) a.htrace[4 * (it - 1) + q] = hyper[q];
    const double lk0 = c[kCtlLk0];
    if (isnan(lkh)) {                                  // :345
        c[kCtlReason] = 2.0; c[kCtlDone] = 1.0;
    } else if (it > 1 && it > a.n0 && lkh >= lk0 && fabs(1 - lkh / lk0) < a.tol) {  // :346-347
        c[kCtlReason] = 1.0; c[kCtlDone] = 1.0;
    } else {
        c[kCtlLk0] = lkh;                              // :348
        if (it >= a.itmax) { c[kCtlReason] = 0.0; c[kCtlDone] = 1.0; }
    }
}

// cid[d] = 1 + index of the first maximum over k of alh[d][k] / beh[k]   (R/utils.R:906)
__global__ void __launch_bounds__(kBlock)
cluster_id_kernel(int64_t rows, int RS, int r, const double *__restrict__ alh,
                  const double *__restrict__ beh, int32_t *__restrict__ cid) {
    const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (j >= rows) return;
    int best = 0;
    double bv = alh[j * RS] / beh[0];
    for (int k = 1; k < r; k++) {
        const double v = alh[j * RS + k] / beh[k];
        if (v > bv) { bv = v; best = k; }
    }
    cid[j] = best + 1;
}

}  // namespace vb
