// Kernels that do not depend on the padded rank; included only by vbnmf.cu (single definition).
#pragma once
#include "kernels.cuh"

namespace vb {

// ---- one-time helpers ----------------------------------------------------------------------
// sum over nonzeros of lgamma(x+1) (src/vbnmf_update.cpp:80-81; zeros contribute 0) and of
// -x log x + x (R/factorize.R:45-46); out[0], out[1].  out[2] = number of values that are not
// integers in [0, 65535] (0 -> the packed 16-bit layout of the sweep applies); out[3] = number of
// negative or non-finite values (the handle is refused: the reference's bound would be NaN).
template <typename VT>
__global__ void __launch_bounds__(kBlock)
count_constants_kernel(int64_t nnz, const VT *__restrict__ val, double *__restrict__ part,
                       double *__restrict__ out, unsigned *counter) {
    __shared__ double sm[kWarpsPerBlock];
    double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < nnz;
         t += (int64_t)gridDim.x * kBlock) {
        const double x = (double)val[t];
        if (!(x >= 0.0 && x <= 1.7e308)) { d += 1.0; continue; }
        a += lgamma(x + 1.0);
        if (x > 0) b += -x * log(x) + x;
        if (!(x <= 65535.0 && x == floor(x))) c += 1.0;
    }
    a = block_sum(a, sm);
    b = block_sum(b, sm);
    c = block_sum(c, sm);
    d = block_sum(d, sm);
    if (threadIdx.x == 0) {
        part[blockIdx.x * 4 + 0] = a;
        part[blockIdx.x * 4 + 1] = b;
        part[blockIdx.x * 4 + 2] = c;
        part[blockIdx.x * 4 + 3] = d;
    }
    last_block_reduce(part, 4, out, counter, sm);
}

// expand CSC column pointers into a per-nonzero column index.  With counts != nullptr (first
// call for a matrix) also: count the entries with a non-zero value per row and column (an explicit
// zero does not make a row non-empty: the reference tests rowSums/colSums == 0,
// R/bayesian.R:244-247) and flag row indices outside [0, n) in bad[0] (those are not counted).
template <typename VT>
__global__ void __launch_bounds__(kBlock)
expand_cols_kernel(int64_t m, int64_t n, const int64_t *__restrict__ colptr,
                   const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                   int32_t *__restrict__ colof, unsigned long long *__restrict__ row_count,
                   unsigned long long *__restrict__ col_count, unsigned *__restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kBlock) >> 5;
    for (int64_t j = warp; j < m; j += nwarps) {
        const int64_t beg = colptr[j], end = colptr[j + 1];
        unsigned cc = 0;
        for (int64_t t = beg + lane; t < end; t += 32) {
            colof[t] = (int32_t)j;
            if (row_count) {
                const int32_t i = rowidx[t];
                if (i < 0 || i >= n) { *bad = 1u; continue; }
                if (val[t] != (VT)0) { atomicAdd(row_count + i, 1ull); cc++; }
            }
        }
        if (row_count) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cc += __shfl_xor_sync(kFull, cc, o);
            if (lane == 0) col_count[j] = cc;
        }
    }
}

// ---- renumbering on the device ----------------------------------------------------------------
// sort key for "descending count, stable": key = 2^31 - 1 - count (counts are < 2^31)
__global__ void __launch_bounds__(kBlock)
order_keys_kernel(int64_t count, const unsigned long long *__restrict__ cnt,
                  uint32_t *__restrict__ key, uint32_t *__restrict__ idx,
                  unsigned *__restrict__ nzero) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (t >= count) return;
    const unsigned long long c = cnt[t];
    key[t] = 0x7fffffffu - (uint32_t)(c > 0x7fffffffull ? 0x7fffffffull : c);
    idx[t] = (uint32_t)t;
    if (c == 0ull && nzero) atomicAdd(nzero, 1u);
}
// sorted position pos (descending count) -> device row: dealt round-robin over S slabs of T rows
__global__ void __launch_bounds__(kBlock)
deal_kernel(int64_t count, int S, int T, const uint32_t *__restrict__ order,
            int32_t *__restrict__ dev) {
    const int64_t pos = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (pos >= count) return;
    dev[order[pos]] = (int32_t)((pos % S) * T + pos / S);
}

// ---- panels between the caller's layout and the device layout -----------------------------------
// caller: W side src[k * cnt + i] (cnt x r, column-major), H side src[j * r + k] (r x cnt)
// device: panel[dev[i] * rs + k], rows/entries not addressed stay as they are (zeroed before)
// dst1 is an l panel (layout tsplit, panel_ofs), dst2 a row-major one
__global__ void __launch_bounds__(kBlock)
scatter_panel_kernel(int64_t cnt, int r, int rs, const int32_t *__restrict__ dev,
                     const double *__restrict__ src, bool wside, double *__restrict__ dst1,
                     double *__restrict__ dst2, int tsplit) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (t >= cnt * r) return;
    int64_t i; int k;
    if (wside) { k = (int)(t / cnt); i = t - (int64_t)k * cnt; }
    else { i = t / r; k = (int)(t - i * r); }
    const double v = src[t];
    const int64_t d = dev[i];
    if (dst1) dst1[panel_ofs(d, k, rs, tsplit)] = v;
    if (dst2) dst2[d * rs + k] = v;
}
// mode 0: out = panel; 1: out = panel / be_k; 2: out = panel / be_k^2 (0 when !has_post)
__global__ void __launch_bounds__(kBlock)
gather_panel_kernel(int64_t cnt, int r, int rs, const int32_t *__restrict__ dev,
                    const double *__restrict__ panel, bool wside, const double *__restrict__ be,
                    int mode, double *__restrict__ out, int tsplit) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (t >= cnt * r) return;
    int64_t i; int k;
    if (wside) { k = (int)(t / cnt); i = t - (int64_t)k * cnt; }
    else { i = t / r; k = (int)(t - i * r); }
    double v = panel[panel_ofs(dev[i], k, rs, tsplit)];
    if (mode >= 1) v = v / be[k];
    if (mode == 2) v = v / be[k];
    out[t] = v;
}
__global__ void __launch_bounds__(kBlock)
gather_i32_kernel(int64_t cnt, const int32_t *__restrict__ dev, const int32_t *__restrict__ src,
                  int32_t *__restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (t < cnt) out[t] = src[dev[t]];
}
// flags[k] = |max_i e_ik - min_i e_ik| < tol over the valid rows, e = al / be_k (R/bayesian.R:368-369)
// one CTA per k
__global__ void __launch_bounds__(kBlock)
uniform_columns_kernel(int64_t rows, int T, int S, int64_t nvalid, int rs,
                       const double *__restrict__ al, const double *__restrict__ be, double tol,
                       int32_t *__restrict__ flags) {
    __shared__ double smx[kWarpsPerBlock], smn[kWarpsPerBlock];
    const int k = blockIdx.x;
    double mx = -INFINITY, mn = INFINITY;
    for (int64_t row = threadIdx.x; row < rows; row += kBlock) {
        const int64_t slab = row / T, local = row - slab * T;
        if (local * S + slab >= nvalid) continue;
        const double v = al[row * rs + k] / be[k];
        mx = v > mx ? v : mx;
        mn = v < mn ? v : mn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(kFull, mx, o));
        mn = fmin(mn, __shfl_xor_sync(kFull, mn, o));
    }
    if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; smn[threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kWarpsPerBlock; w++) { mx = fmax(mx, smx[w]); mn = fmin(mn, smn[w]); }
        flags[k] = fabs(mx - mn) < tol ? 1 : 0;
    }
}

// sort key of every nonzero for one pass of the tiled layout:
//   key = slab(tile side) * NO + owner(device row of the owner side)
//   payload = nonzero index, or (packed-16 layout, val != nullptr) the packed word
//             {count << 16 | tile row} itself, so that the sorted payloads ARE the entries in
//             segment-major order and nothing has to be gathered through a permutation afterwards
template <typename VT>
__global__ void __launch_bounds__(kBlock)
make_keys_kernel(int64_t nnz, const int32_t *__restrict__ rowidx, const int32_t *__restrict__ colof,
                 const int32_t *__restrict__ gene_dev, const int32_t *__restrict__ cell_dev, int T,
                 int64_t NO, bool cols_pass, const VT *__restrict__ val, uint32_t *__restrict__ key,
                 uint32_t *__restrict__ payload) {
    for (int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x; t < nnz;
         t += (int64_t)gridDim.x * kBlock) {
        const int64_t gd = gene_dev[rowidx[t]], cd = cell_dev[colof[t]];
        const int64_t k = cols_pass ? (gd / T) * NO + cd : (cd / T) * NO + gd;
        key[t] = (uint32_t)k;
        payload[t] = val ? (((uint32_t)val[t] << 16) | (uint32_t)((cols_pass ? gd : cd) % T))
                         : (uint32_t)t;
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
// Conflict-aware schedule of one segment.  A group of the sweep processes NL nonzeros per step
// (NL = 8, or 4 when two lanes share a nonzero) and a step costs as many shared-memory wavefronts
// per gather as the largest number of its tile rows that fall into the same residue class
// (row mod 8; with NL = 4 the four rows of a step must also have equal parity, so the even and the
// odd residues are scheduled as two independent classes of NB = 4 buckets).  With bucket counts
// c_b the cheapest schedule in K steps has max(K, max_b c_b) wavefronts.  It is reached by
//   R "single" steps: step k holds the k-th nonzero of every bucket that still has one (lane =
//     bucket, empty lanes are holes), conflict free;
//   P "pair" steps of two half rows of <= NL/2 nonzeros with distinct buckets: the r_b = c_b - R
//     nonzeros left in each bucket are dealt cyclically over the 2P half rows (r_b <= 2P, so no
//     bucket repeats inside a half row) -> at most a 2-way conflict;
// with R + P = K and P the smallest value for which L <= R + 2P and sum_b r_b <= NL * P.  K is the
// number of steps the kernel executes anyway (whole chunks of 4 steps), so the schedule costs no
// extra instructions.  Holes are zero-count words that point at a row of an unused residue class.
struct SegPlan { int R, P; };

template <int NB>
__device__ __forceinline__ SegPlan plan_class(const int (&c)[NB], int NL, int K) {
    int L = 0;
#pragma unroll
    for (int b = 0; b < NB; b++) L = max(L, c[b]);
    SegPlan pl;
    for (int P = max(0, L - K);; P++) {
        const int R = K - P;
        int rest = 0;
#pragma unroll
        for (int b = 0; b < NB; b++) rest += max(0, c[b] - R);
        if (rest <= NL * P) { pl.R = R; pl.P = P; return pl; }
    }
}

// steps of a class with n nonzeros, the fullest bucket holding L of them: enough for all of them
// and for a 2-way conflict at worst (L <= 2K), rounded up to a multiple of `mult`
__device__ __forceinline__ int class_steps(int n, int L, int NL, int mult) {
    const int k0 = max((n + NL - 1) / NL, (L + 1) / 2);
    return ((k0 + mult - 1) / mult) * mult;
}

// residue (tile row mod 8) -> (class, bucket) and back.  NL = 8: one class, bucket = residue.
// NL = 4: class = parity, bucket = residue / 2.
__device__ __forceinline__ int res_class(int rr, int NL) { return NL == 8 ? 0 : (rr & 1); }
__device__ __forceinline__ int res_bucket(int rr, int NL) { return NL == 8 ? rr : (rr >> 1); }
__device__ __forceinline__ int bucket_res(int cls, int b, int NL) { return NL == 8 ? b : 2 * b + cls; }

struct SegSchedule {
    int K[2];        // steps per class
    SegPlan pl[2];
    int offr[2][8];  // pair part: first dealt index of each bucket
    int cnt[2][8];
};

// counts per residue -> schedule.  NL = 8: class 0 only, K a multiple of kmult (4: every stored
// step is scheduled; 1: the schedule uses the fewest steps and the rest of the last stored chunk
// of 4 steps is all holes, which kernels that skip the gathers of hole entries get for free).
// NL = 4: two classes of K multiples of 2 (an odd number of step pairs gets one more pair of hole
// steps in class 1).
// sbs ("side by side", the 4-unit split layout, NL = 8 lanes): the two parity classes are scheduled
// like the NL = 4 classes but run in the SAME steps, even rows in lanes 0..3 and odd rows in lanes
// 4..7, K = the larger of the two step counts (a multiple of kmult).
__device__ __forceinline__ void make_schedule(const int (&cnt8)[8], int NLp, int kmult, bool sbs,
                                              SegSchedule &sc, bool one4 = false) {
    if (one4) {  // one class of four buckets (cnt8[0..3]), four lanes, K a multiple of kmult
        for (int cl = 0; cl < 2; cl++) {
            sc.K[cl] = 0; sc.pl[cl].R = 0; sc.pl[cl].P = 0;
            for (int b = 0; b < 8; b++) { sc.cnt[cl][b] = 0; sc.offr[cl][b] = 0; }
        }
        int n = 0, L = 0, c[4];
        for (int b = 0; b < 4; b++) { c[b] = sc.cnt[0][b] = cnt8[b]; n += c[b]; L = max(L, c[b]); }
        sc.K[0] = class_steps(n, L, 4, kmult);
        if (sc.K[0] == 0) return;
        sc.pl[0] = plan_class<4>(c, 4, sc.K[0]);
        int acc = 0;
        for (int b = 0; b < 4; b++) { sc.offr[0][b] = acc; acc += max(0, c[b] - sc.pl[0].R); }
        return;
    }
    const int NL = sbs ? 4 : NLp;
    const int ncls = NL == 8 ? 1 : 2, NB = NL;
    for (int cl = 0; cl < 2; cl++) {
        sc.K[cl] = 0; sc.pl[cl].R = 0; sc.pl[cl].P = 0;
        for (int b = 0; b < 8; b++) { sc.cnt[cl][b] = 0; sc.offr[cl][b] = 0; }
    }
    for (int rr = 0; rr < 8; rr++) sc.cnt[res_class(rr, NL)][res_bucket(rr, NL)] = cnt8[rr];
    for (int cl = 0; cl < ncls; cl++) {
        int n = 0, L = 0;
        for (int b = 0; b < NB; b++) { n += sc.cnt[cl][b]; L = max(L, sc.cnt[cl][b]); }
        sc.K[cl] = class_steps(n, L, NL, (NL == 8 || sbs) ? kmult : 2);
    }
    if (sbs) sc.K[0] = sc.K[1] = max(sc.K[0], sc.K[1]);
    else if (ncls == 2 && ((sc.K[0] + sc.K[1]) & 3)) sc.K[1] += 2;
    for (int cl = 0; cl < ncls; cl++) {
        if (sc.K[cl] == 0) continue;
        if (NL == 8) {
            int c[8];
            for (int b = 0; b < 8; b++) c[b] = sc.cnt[cl][b];
            sc.pl[cl] = plan_class<8>(c, 8, sc.K[cl]);
        } else {
            int c[4];
            for (int b = 0; b < 4; b++) c[b] = sc.cnt[cl][b];
            sc.pl[cl] = plan_class<4>(c, 4, sc.K[cl]);
        }
        int acc = 0;
        for (int b = 0; b < NB; b++) {
            sc.offr[cl][b] = acc;
            acc += max(0, sc.cnt[cl][b] - sc.pl[cl].R);
        }
    }
}

// item index (step * NL + lane, steps of class 1 after those of class 0) of the k-th nonzero of
// residue rr
__device__ __forceinline__ int schedule_item(const SegSchedule &sc, int NLp, bool sbs, int rr,
                                             int k, bool one4 = false) {
    const int NL = sbs ? 4 : NLp;
    const int cl = one4 ? 0 : res_class(rr, NL), b = one4 ? rr : res_bucket(rr, NL);
    const int R = sc.pl[cl].R, P = sc.pl[cl].P;
    int step, lane;
    if (k < R) {
        step = k; lane = b;
    } else {
        const int t = sc.offr[cl][b] + (k - R);
        const int row = t % (2 * P), level = t / (2 * P);
        step = R + row % P;
        lane = (row >= P) ? NL - 1 - level : level;
    }
    if (sbs) return step * 8 + cl * 4 + lane;
    return ((cl ? sc.K[0] : 0) + step) * NL + lane;
}

// residue counts of a segment's words, one warp per segment (every lane gets all 8 counts)
__device__ __forceinline__ void warp_residue_counts(const uint32_t *__restrict__ words,
                                                    int64_t beg, int64_t end, int lane,
                                                    int (&cnt)[8]) {
#pragma unroll
    for (int b = 0; b < 8; b++) cnt[b] = 0;
    for (int64_t base = beg; base < end; base += 32) {
        const int64_t t = base + lane;
        const int rr = t < end ? (int)(words[t] & 7u) : 8;
#pragma unroll
        for (int b = 0; b < 8; b++) cnt[b] += __popc(__ballot_sync(kFull, rr == b));
    }
}

// quads (4 entries = 16 bytes) of every segment under the schedule above; len4[E] = 0.
// One warp per segment.
// Schedule modes of the 8-lane layouts.  kSchedPlain: eight residue classes (row mod 8), one lane
// each.  kSchedSbs: two parity classes side by side (4-unit split layout).  kSchedCls4: four
// classes (row mod 4) of TWO lanes each (split layout with two dense units in block B, ranks 19
// and 20): the nonzeros of class c are dealt alternately to the "virtual residues" c and c + 4,
// which are then scheduled like eight residue classes -- lane l of a single step holds a row of
// class l mod 4, and lanes l, l + 4 read the two units of their rows in opposite order.
// kSchedOne4 (4-lane groups, NL = 4): ONE set of four classes (row mod 4), one lane each.
enum { kSchedPlain = 0, kSchedSbs = 1, kSchedCls4 = 2, kSchedOne4 = 3 };

// counts per real residue -> counts per virtual residue (kSchedCls4)
__device__ __forceinline__ void virtual_counts(int (&cnt)[8]) {
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const int c4 = cnt[b] + cnt[b + 4];
        cnt[b] = (c4 + 1) >> 1;
        cnt[b + 4] = c4 >> 1;
    }
}

__global__ void __launch_bounds__(kBlock)
plan_p16_kernel(int64_t E, const int64_t *__restrict__ ptr, const uint32_t *__restrict__ words,
                int NL, int kmult, int mode, uint32_t *__restrict__ len4,
                uint8_t *__restrict__ dead) {
    const bool sbs = mode == kSchedSbs;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kBlock) >> 5;
    if (warp == 0 && lane == 0) len4[E] = 0u;
    for (int64_t e = warp; e < E; e += nwarps) {
        const int64_t beg = ptr[e], end = ptr[e + 1];
        if (beg == end) { if (lane == 0) { len4[e] = 0u; if (dead) dead[e] = 0; } continue; }
        int cnt[8];
        warp_residue_counts(words, beg, end, lane, cnt);
        if (mode == kSchedCls4) virtual_counts(cnt);
        if (mode == kSchedOne4) {
            int n = 0, L = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) { const int c4 = cnt[b] + cnt[b + 4]; n += c4; L = max(L, c4); }
            const int K = class_steps(n, L, 4, kmult), Kst = (K + 3) & ~3;
            if (lane == 0) {
                len4[e] = (uint32_t)Kst;  // Kst steps x 4 entries / 4
                if (dead) dead[e] = (uint8_t)(Kst - K);
            }
            continue;
        }
        int n0 = 0, n1 = 0, L0 = 0, L1 = 0;
        const int NLs = sbs ? 4 : NL;
#pragma unroll
        for (int rr = 0; rr < 8; rr++) {
            if (res_class(rr, NLs)) { n1 += cnt[rr]; L1 = max(L1, cnt[rr]); }
            else { n0 += cnt[rr]; L0 = max(L0, cnt[rr]); }
        }
        int K = class_steps(n0, L0, NLs, (NLs == 8 || sbs) ? kmult : 2);
        if (sbs) K = max(K, class_steps(n1, L1, 4, kmult));
        else if (NL != 8) K += class_steps(n1, L1, NL, 2);
        const int Kst = (K + 3) & ~3;  // stored steps: whole chunks of 4
        if (lane == 0) {
            len4[e] = (uint32_t)(Kst * NL / 4);
            if (dead) dead[e] = (uint8_t)(Kst - K);  // all-hole steps at the end (kmult = 1)
        }
    }
}

// ---- segments in the order of decreasing length inside windows of W owners ------------------------
// The four 8-lane groups of a warp run their chunk loops in lock step, so a warp takes as long as
// its longest segment: with ~115 nonzeros per segment (r = 20) the longest of four is ~9 % above
// the mean.  Inside every window of W consecutive owners of a slab the segments are therefore
// stored in the order of decreasing step count: neighbouring segments -- the ones a warp processes
// together -- have the same number of steps, while the owner rows and partial statistics a CTA
// touches stay within a window (sorting a whole slab scatters them over the panel: measured 25 %
// SLOWER in the cell-owner pass).  sort key of segment e: window, then 1023 - min(steps, 1023);
// value e (stable sort: equal lengths stay in owner order).
__global__ void __launch_bounds__(kBlock)
seg_order_keys_kernel(int64_t E, int64_t NO, int W, const uint32_t *__restrict__ len4,
                      const uint8_t *__restrict__ dead, int NL, uint32_t *__restrict__ key,
                      uint32_t *__restrict__ val) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e >= E) return;
    int steps = (int)(len4[e] * 4u / (uint32_t)NL) - (dead ? (int)dead[e] : 0);
    steps = steps > 1023 ? 1023 : steps;
    const int64_t slab = e / NO, o = e - slab * NO, nwin = (NO + W - 1) / W;
    key[e] = (uint32_t)(slab * nwin + o / W) * 1024u + (uint32_t)(1023 - steps);
    val[e] = (uint32_t)e;
}
// order[pos] = segment at position pos -> seg[pos] = its owner row, len4p / deadp = its quads and
// dead steps (len4p[E] = 0)
__global__ void __launch_bounds__(kBlock)
seg_permute_kernel(int64_t E, int64_t NO, const uint32_t *__restrict__ order,
                   const uint32_t *__restrict__ len4, const uint8_t *__restrict__ dead,
                   uint32_t *__restrict__ len4p, uint8_t *__restrict__ deadp,
                   uint32_t *__restrict__ seg) {
    const int64_t pos = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (pos > E) return;
    if (pos == E) { len4p[E] = 0u; return; }
    const uint32_t e = order ? order[pos] : (uint32_t)pos;
    len4p[pos] = len4[e];
    if (dead) deadp[pos] = dead[e];
    seg[pos] = (uint32_t)(e - (pos / NO) * NO);
}

// position of item p of a segment stored in blocks of B = 4*NL entries (4 steps): item p of a
// block -> quad p mod NL, word p div NL, so that word u of the quads the NL lanes load in one
// 128-bit access is step u of the block
__device__ __forceinline__ int64_t p16_position(int64_t p, int NL) {
    const int B = 4 * NL;
    const int64_t blk = p / B;
    const int pin = (int)(p - blk * B);
    return blk * B + 4 * (pin % NL) + pin / NL;
}

// One warp per segment: fill the segment with hole words, then scatter the nonzeros to their
// scheduled places as {count << 16 | tile row} (the k-th word of a residue in sorted order is the
// k-th nonzero of its bucket).  nvalid / S: rows of the tile side that exist (a hole must point
// at a real row: local * S + slab < nvalid; row 0 of a slab always is).
__global__ void __launch_bounds__(kBlock)
build_segments_p16_kernel(int64_t E, int64_t NO, const int64_t *__restrict__ ptr,
                          const uint32_t *__restrict__ ptr4, const uint32_t *__restrict__ seg,
                          const uint32_t *__restrict__ words,
                          int NL, int kmult, int mode, int64_t nvalid, int S,
                          uint32_t *__restrict__ ent_out) {
    const bool sbs = mode == kSchedSbs, cls4 = mode == kSchedCls4, one4 = mode == kSchedOne4;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t warp = ((int64_t)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * kBlock) >> 5;
    // pos: position in the stored order (ptr4, seg); e: segment in (slab, owner) order (ptr, words)
    for (int64_t pos = warp; pos < E; pos += nwarps) {
        const int64_t slab = pos / NO;
        const int64_t e = slab * NO + seg[pos];
        const int64_t beg = ptr[e], end = ptr[e + 1];
        if (beg == end) continue;
        uint32_t *dst = ent_out + (int64_t)ptr4[pos] * 4;
        const int nitems = (int)(ptr4[pos + 1] - ptr4[pos]) * 4;
        int cnt[8], cntv[8];
        warp_residue_counts(words, beg, end, lane, cnt);
#pragma unroll
        for (int b = 0; b < 8; b++) cntv[b] = cnt[b];
        if (cls4) virtual_counts(cntv);
        if (one4) {
#pragma unroll
            for (int b = 0; b < 4; b++) { cntv[b] = cnt[b] + cnt[b + 4]; cntv[b + 4] = 0; }
        }
        SegSchedule sc;
        make_schedule(cntv, NL, kmult, sbs, sc, one4);
        // holes: lane <-> bucket in the single steps, so the residue of the lane's bucket is free
        for (int p = lane; p < nitems; p += 32) {
            const int step = p / NL, ln = p - step * NL;
            const int cl = sbs ? (ln >> 2) : ((NL != 8 && !one4 && step >= sc.K[0]) ? 1 : 0);
            int rr = sbs ? 2 * (ln & 3) + cl : ((cls4 || one4) ? (ln & 3) : bucket_res(cl, ln, NL));
            if ((int64_t)rr * S + slab >= nvalid) rr = 0;
            dst[p16_position(p, NL)] = (uint32_t)rr;
        }
        __syncwarp();
        int seen[8];
#pragma unroll
        for (int b = 0; b < 8; b++) seen[b] = 0;
        for (int64_t base = beg; base < end; base += 32) {
            const int64_t t = base + lane;
            const bool valid = t < end;
            const uint32_t w = valid ? words[t] : 0u;
            const int rr = valid ? (int)(w & 7u) : 8;
            int k = 0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const unsigned mk = __ballot_sync(kFull, rr == b);
                if (rr == b) k = seen[b] + __popc(mk & lt);
                seen[b] += __popc(mk);
            }
            int vr = rr, vk = k;
            if (cls4 && valid) {
                // rank in the class (rows = c mod 4: residues c, then c + 4), dealt alternately
                const int kc = rr < 4 ? k : cnt[rr - 4] + k;
                vr = (rr & 3) + 4 * (kc & 1);
                vk = kc >> 1;
            }
            if (one4 && valid) {
                // rank in the class: residues c, then c + 4 -- the other way round in segments at
                // odd positions.  The two 4-lane groups of a bank phase sit at an even and an odd
                // position (CTA ranges start at even positions, split_p16_kernel): with a single
                // unit in block B (ranks 17, 18; bank group = row mod 8) lanes l and l + 4 collide
                // when their rows agree in bit 2, which this order avoids except around the
                // middle of the segments.
                vr = rr & 3;
                if ((pos & 1) == 0) vk = rr < 4 ? k : cnt[rr - 4] + k;
                else vk = rr >= 4 ? k : cnt[rr + 4] + k;
            }
            if (valid) dst[p16_position(schedule_item(sc, NL, sbs, vr, vk, one4), NL)] = w;
        }
        __syncwarp();
    }
}

// ptr4[e] |= number of all-hole steps at the end of segment e (0..3; quad pointers are multiples
// of 8, the sweep kernels of the split layout mask the tag off)
__global__ void __launch_bounds__(kBlock)
tag_dead_kernel(int64_t E, const uint8_t *__restrict__ dead, uint32_t *__restrict__ ptr4) {
    const int64_t e = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (e < E) ptr4[e] |= (uint32_t)dead[e];
}

// Work split of a packed-16 pass over the persistent CTAs: split[b] = first segment whose cost
// prefix is >= b * total / nparts, split[nparts] = E.  Cost of a segment = its quads + kappa: the
// per-segment part (pointers, owner row, cross-lane sum, store) is worth a few quads of gathers, so
// a CTA whose range holds the short segments of a slab (the owners are sorted by count) would be
// late under an equal-entries split (measured: sm__cycles_elapsed.max 5.7 % above the mean).
// The boundaries are even positions (see kSchedOne4 in build_segments_p16_kernel).
__global__ void split_p16_kernel(int nparts, int64_t E, const uint32_t *__restrict__ ptr4,
                                 double kappa, int64_t *__restrict__ split) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nparts) return;
    if (b == nparts) { split[b] = E; return; }
    const double target = ((double)ptr4[E] + kappa * (double)E) * b / nparts;
    int64_t lo = 0, hi = E;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((double)ptr4[mid] + kappa * (double)mid < target) lo = mid + 1; else hi = mid;
    }
    split[b] = lo & ~(int64_t)1;
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

// One thread: the host half of an ML iteration on the device (R/factorize.R:193,207-209).  Called
// once per loop turn t = 1 .. itmax + 1, after the cell-owner sweep at the current (w, h):
//   lik = (sum x log(wh) - sum_k colSums(w)_k rowSums(h)_k + sum(-x log x + x)) / (n m)   :40-49
// is the likelihood of the state left by iteration t - 1; |lkold - lk0| < tol |lkold| (:207) or
// t - 1 == itmax ends the run (the update kernels of this and later turns are then no-ops).
struct MlControlArgs {
    double *ctl;
    const double *scal;   // [colSums(w) rs | ...]
    const double *tail;   // [rowSums(h) rs | . . . | . , xlogp at rs + 4]
    double *trace;
    double n, m_global, mlconst, tol;
    int r, rs, itmax;
};
__global__ void ml_control_kernel(const MlControlArgs a) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double *c = a.ctl;
    if (c[kCtlDone] != 0.0) return;
    const int t = (int)c[kCtlIt] + 1;   // loop turn
    c[kCtlIt] = t;
    if (t == 1) return;                  // nothing to evaluate before the first update
    double swh = 0.0;
    for (int k = 0; k < a.r; k++) swh += a.scal[k] * a.tail[k];
    const double lk0 = (a.tail[a.rs + 4] - swh + a.mlconst) / a.n / a.m_global;
    const double lkold = c[kCtlLk0];
    a.trace[t - 2] = lk0;
    c[kCtlLkh] = lk0;
    if (fabs(lkold - lk0) < a.tol * fabs(lkold)) { c[kCtlReason] = 1.0; c[kCtlDone] = 1.0; return; }
    c[kCtlLk0] = lk0;
    if (t - 1 >= a.itmax) { c[kCtlReason] = 0.0; c[kCtlDone] = 1.0; }
}

// ---- on-device 'random' initialiser (vb_init(initializer = 'random'), R/bayesian.R:111-115) ----
// w_ik ~ Gamma(shape aw, scale bw/aw), h_kj ~ Gamma(shape ah, scale bh/ah).  R's RNG stream cannot be
// reproduced, so the draw is DEFINED here by a counter RNG keyed by (seed, side, global row, k):
// the same matrix entry gets the same value whatever the device layout or the sharding of the
// cells.  ccfindr_b200/synth.py:device_random_init_reference restates it for the tests.
//   stream:  state0 = mix(seed) ^ mix(id + 0x632BE59BD9B4E019), id = side << 62 | row << 6 | k;
//            next() = mix(state += 0x9E3779B97F4A7C15)            (splitmix64)
//   uniform: (next() >> 11 + 0.5) * 2^-53;  normal: sqrt(-2 ln u1) cos(2 pi u2) (Box-Muller)
//   gamma(a >= 1): Marsaglia-Tsang (d = a - 1/3, c = 1/sqrt(9d)); a < 1: gamma(a + 1) * u^(1/a)
VB_HD unsigned long long vb_mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct VbStream {
    unsigned long long s;
    VB_HD VbStream(unsigned long long seed, unsigned long long id)
        : s(vb_mix64(seed) ^ vb_mix64(id + 0x632BE59BD9B4E019ull)) {}
    VB_HD unsigned long long next() { s += 0x9E3779B97F4A7C15ull; return vb_mix64(s); }
    VB_HD double uniform() { return ((double)(next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
    VB_HD double normal() {
        const double u1 = uniform(), u2 = uniform();
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
    }
    VB_HD double gamma(double a) {
        const double a1 = a < 1.0 ? a + 1.0 : a;
        const double d = a1 - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        double g = 0.0;
        for (int it = 0; it < 1000; it++) {
            const double x = normal();
            const double v0 = 1.0 + c * x;
            const double u = uniform();
            if (v0 <= 0.0) continue;
            const double v = v0 * v0 * v0;
            if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) { g = d * v; break; }
        }
        if (a < 1.0) g *= pow(uniform(), 1.0 / a);
        return g;
    }
};

// one thread per (original row, k): panel[dev[row]][k] = scale * gamma(shape), also into `mirror`
// (alw/alh hold ew/eh = w/h before the first update, R/bayesian.R:170)
__global__ void __launch_bounds__(kBlock)
init_random_kernel(int64_t rows, int r, int rs, const int32_t *__restrict__ dev, int side,
                   int64_t row_offset, unsigned long long seed, double shape, double scale,
                   double *__restrict__ panel, double *__restrict__ mirror, int tsplit) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (t >= rows * r) return;
    const int64_t row = t / r;
    const int k = (int)(t - row * r);
    VbStream st(seed, ((unsigned long long)side << 62) |
                          ((unsigned long long)(row + row_offset) << 6) | (unsigned long long)k);
    const double v = scale * st.gamma(shape);
    const int64_t d = dev[row];
    panel[panel_ofs(d, k, rs, tsplit)] = v;
    mirror[d * rs + k] = v;
}

// ---- all-reduce of the W-side statistics over NVLink peer memory -----------------------------
// One process per GPU; every rank owns an exchange region (layout below) that its peers map
// through CUDA IPC.  An all-reduce with sequence number seq is two launches:
//   publish: copy the local vector into the own `in` buffer seq & 1, then (last CTA, after a
//            system-scope fence) store seq into arrive[my rank] of EVERY rank's region;
//   reduce:  reduce-scatter + all-gather (xchg_reduce_kernel): rank r sums slice r of all `in`
//            buffers in rank order with P2P loads, publishes the reduced slice, and collects the
//            other ranks' slices.
// A rank can only publish seq + 2 (the next use of the same buffers) after its reduce of seq + 1,
// which waits for every peer's publish of seq + 1, which follows that peer's reduce of seq in
// stream order: double buffering is enough, no extra barrier.  The payload is 1.6 MB at C2 (3.6 MB
// at C3) per iteration.  Round 1 had every rank read ALL peers' vectors (7 vectors over NVLink on 8
// GPUs) and measured 20 us per iteration slower than NCCL there (profiles/r01_peer_allreduce_ab.txt);
// the reduce-scatter form moves 1.75 vectors.  Opt-in (VBNMF_PEER_ALLREDUCE=1).
constexpr int kXchgFlags = 16;   // u64 flags per set (>= ranks of one NVLink domain we use)
// Region of a rank: [arrive flags (kXchgFlags) | slice flags (kXchgFlags) | in 0 | in 1 | res 0 | res 1]
struct XchgArgs {
    unsigned long long *peer[kXchgFlags];  // base of every rank's region (own region at [rank])
    int nranks, rank;
    unsigned long long seq;
    int64_t n2;                 // payload in double2 units
    int64_t buf_stride;         // bytes between consecutive buffers
    const double *ctl;
};
__device__ __forceinline__ double2 *xchg_buf(unsigned long long *base, const XchgArgs &a, int which) {
    char *p = reinterpret_cast<char *>(base) + 2 * kXchgFlags * 8 +
              ((int64_t)which * 2 + (int64_t)(a.seq & 1ull)) * a.buf_stride;
    return reinterpret_cast<double2 *>(p);
}

// publish: copy the local vector into the own `in` buffer of this sequence number, then (last CTA,
// after a system-scope fence) raise arrive[my rank] = seq in EVERY rank's region
__global__ void __launch_bounds__(kBlock)
xchg_publish_kernel(const XchgArgs a, const double2 *__restrict__ src, unsigned *counter) {
    if (a.ctl && a.ctl[kCtlDone] != 0.0) return;
    double2 *dst = xchg_buf(a.peer[a.rank], a, 0);
    for (int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x; i < a.n2;
         i += (int64_t)gridDim.x * kBlock)
        dst[i] = src[i];
    __shared__ bool is_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (!is_last) return;
    __threadfence_system();
    if (threadIdx.x < a.nranks) {
        volatile unsigned long long *f = a.peer[threadIdx.x] + a.rank;
        *f = a.seq;
    }
}

// bounded spin on a flag of the own region: a peer that never arrives traps this rank instead of
// hanging the GPU
__device__ __forceinline__ void xchg_wait(volatile unsigned long long *f, unsigned long long seq) {
    for (unsigned long long spin = 0; *f < seq; spin++)
        if (spin > (1ull << 31)) __trap();
}

// reduce-scatter + all-gather over peer memory in one launch.  Rank r owns slice r of the vector:
//   phase 1: wait for every rank's publish, sum slice r of all `in` buffers in rank order (P2P
//            loads) into the own `res` buffer and into `out`; last CTA raises slice[r] = seq in
//            every rank's region;
//   phase 2: for every other rank p, wait for slice[p] and copy ITS reduced slice into `out`.
// Every element is summed by exactly one rank in a fixed order and copied: bitwise identical on
// all ranks.  NVLink traffic per rank: 2 (nranks - 1)/nranks of the vector, against (nranks - 1)
// vectors for the all-to-all read this replaces.  All CTAs of the launch are resident together
// (grid <= SMs), so the CTAs spinning in phase 2 cannot keep phase 1 from finishing.
__global__ void __launch_bounds__(kBlock)
xchg_reduce_kernel(const XchgArgs a, double2 *__restrict__ out, unsigned *counter) {
    if (a.ctl && a.ctl[kCtlDone] != 0.0) return;
    const int64_t L = (a.n2 + a.nranks - 1) / a.nranks;       // slice length in double2
    unsigned long long *mine = a.peer[a.rank];
    if (threadIdx.x < a.nranks) xchg_wait(mine + threadIdx.x, a.seq);
    __threadfence_system();
    __syncthreads();
    {
        const int64_t lo = (int64_t)a.rank * L, hi = min(a.n2, lo + L);
        double2 *res = xchg_buf(mine, a, 1);
        for (int64_t i = lo + (int64_t)blockIdx.x * kBlock + threadIdx.x; i < hi;
             i += (int64_t)gridDim.x * kBlock) {
            double2 s = make_double2(0.0, 0.0);
            for (int p = 0; p < a.nranks; p++) {
                const double2 v = __ldcv(xchg_buf(a.peer[p], a, 0) + i);
                s.x += v.x;
                s.y += v.y;
            }
            res[i] = s;
            out[i] = s;
        }
    }
    __shared__ bool is_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
    __syncthreads();
    if (is_last) {
        __threadfence_system();
        if (threadIdx.x < a.nranks) {
            volatile unsigned long long *f = a.peer[threadIdx.x] + kXchgFlags + a.rank;
            *f = a.seq;
        }
    }
    for (int q = 1; q < a.nranks; q++) {
        const int p = (a.rank + q) % a.nranks;                 // staggered: not everyone on rank 0 first
        if (threadIdx.x == 0) xchg_wait(mine + kXchgFlags + p, a.seq);
        __syncthreads();
        __threadfence_system();
        const int64_t lo = (int64_t)p * L, hi = min(a.n2, lo + L);
        const double2 *src = xchg_buf(a.peer[p], a, 1);
        for (int64_t i = lo + (int64_t)blockIdx.x * kBlock + threadIdx.x; i < hi;
             i += (int64_t)gridDim.x * kBlock)
            out[i] = __ldcv(src + i);
    }
}

// ---- criterion = 'connectivity' of factorize() (R/factorize.R:194-206) ---------------------------
// lab[row] = index of the first maximum of the row of an l panel (h of the ML path), -1 for the
// padding rows: which.max(h[,j]) of connectivity() (R/factorize.R:51-60)
__global__ void __launch_bounds__(kBlock)
ml_labels_kernel(int64_t rows, int T, int S, int64_t nvalid, int rs, int r, int tsplit,
                 const double *__restrict__ v, int32_t *__restrict__ lab) {
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (row >= rows) return;
    const int64_t slab = row / T, local = row - slab * T;
    if (local * S + slab >= nvalid) { lab[row] = -1; return; }
    int best = 0;
    double bv = v[panel_ofs(row, 0, rs, tsplit)];
    for (int k = 1; k < r; k++) {
        const double x = v[panel_ofs(row, k, rs, tsplit)];
        if (x > bv) { bv = x; best = k; }
    }
    lab[row] = best;
}
// C[a * r + b] += number of cells with old label a and new label b (r x r doubles, zeroed before).
// The connectivity matrices of two labelings are equal iff every row and column of C has at most
// one non-zero; the number of changed pairs is sum_a C(n_a.,2) + sum_b C(n_.b,2) - 2 sum C(C_ab,2).
__global__ void __launch_bounds__(kBlock)
contingency_kernel(int64_t rows, int r, const int32_t *__restrict__ a, const int32_t *__restrict__ b,
                   double *__restrict__ Cm) {
    extern __shared__ unsigned int hist[];
    for (int i = threadIdx.x; i < r * r; i += kBlock) hist[i] = 0u;
    __syncthreads();
    for (int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x; row < rows;
         row += (int64_t)gridDim.x * kBlock) {
        const int la = a[row], lb = b[row];
        if (la >= 0 && lb >= 0) atomicAdd(&hist[la * r + lb], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < r * r; i += kBlock)
        if (hist[i]) atomicAdd(&Cm[i], (double)hist[i]);
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
