// MatrixMarket coordinate file -> CSC on the device (role of Matrix::readMM + as(., 'dgCMatrix') in
// read_10x, R/utils.R:28-54, for the 'matrix.mtx' a 10x pipeline writes).  The host only reads the
// file bytes and the header; the text is parsed on the GPU: newline positions are compacted into
// line starts, one thread parses one "row col value" line, the triples are sorted by (col, row)
// and the column pointers come from a binary search.  Nothing passes through a host-side sparse
// matrix.  Supported: `%%MatrixMarket matrix coordinate real|integer general` (what 10x writes);
// duplicates are rejected.  Included only by vbnmf.cu.
#pragma once
#include "kernels_common.cuh"

namespace vb {

// flag[b] = 1 where a data line starts: byte 0 of the region and every byte after a '\n' that is
// not itself a line break or the end of the data
__global__ void __launch_bounds__(kBlock)
mtx_line_flags_kernel(int64_t nbytes, const char *__restrict__ buf, uint8_t *__restrict__ flag) {
    for (int64_t b = (int64_t)blockIdx.x * kBlock + threadIdx.x; b < nbytes;
         b += (int64_t)gridDim.x * kBlock) {
        const char c = buf[b];
        const bool blank = c == '\n' || c == '\r';
        flag[b] = (!blank && (b == 0 || buf[b - 1] == '\n')) ? 1 : 0;
    }
}

__device__ __forceinline__ bool mtx_space(char c) { return c == ' ' || c == '\t' || c == '\r'; }

// one thread per line: "i j v" (1-based indices) -> key = (j-1) << 32 | (i-1), value; bad[0] is set
// for a malformed line or an index out of range
__global__ void __launch_bounds__(kBlock)
mtx_parse_kernel(int64_t nlines, int64_t nbytes, const char *__restrict__ buf,
                 const int64_t *__restrict__ start, int64_t n, int64_t m,
                 unsigned long long *__restrict__ key, double *__restrict__ val,
                 unsigned *__restrict__ bad) {
    const int64_t L = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (L >= nlines) return;
    int64_t p = start[L];
    auto peek = [&]() -> char { return p < nbytes ? buf[p] : '\n'; };
    auto skip = [&]() { while (mtx_space(peek())) p++; };
    auto uint_field = [&](long long &out) -> bool {
        skip();
        if (peek() < '0' || peek() > '9') return false;
        long long v = 0;
        while (peek() >= '0' && peek() <= '9') { v = v * 10 + (peek() - '0'); p++; if (v > (1ll << 40)) return false; }
        out = v;
        return true;
    };
    long long i = 0, j = 0;
    bool ok = uint_field(i) && uint_field(j);
    skip();
    // value: [+-] digits [. digits] [e|E [+-] digits]
    double sign = 1.0;
    if (peek() == '-') { sign = -1.0; p++; } else if (peek() == '+') p++;
    unsigned long long mant = 0;
    int nd = 0, dexp = 0;
    bool any = false;
    while (peek() >= '0' && peek() <= '9') {
        if (nd < 18) { mant = mant * 10 + (unsigned long long)(peek() - '0'); nd += (mant != 0); }
        else dexp++;
        any = true; p++;
    }
    if (peek() == '.') {
        p++;
        while (peek() >= '0' && peek() <= '9') {
            if (nd < 18) { mant = mant * 10 + (unsigned long long)(peek() - '0'); nd += (mant != 0); dexp--; }
            any = true; p++;
        }
    }
    if (peek() == 'e' || peek() == 'E') {
        p++;
        int es = 1, e = 0;
        if (peek() == '-') { es = -1; p++; } else if (peek() == '+') p++;
        if (peek() < '0' || peek() > '9') ok = false;
        while (peek() >= '0' && peek() <= '9') { e = e * 10 + (peek() - '0'); p++; if (e > 400) e = 400; }
        dexp += es * e;
    }
    skip();
    if (!any || peek() != '\n') ok = false;
    double v = (double)mant;
    if (dexp > 0) v *= pow(10.0, (double)dexp);
    else if (dexp < 0) v /= pow(10.0, (double)(-dexp));
    if (!ok || i < 1 || i > n || j < 1 || j > m) { *bad = 1u; key[L] = ~0ull; val[L] = 0.0; return; }
    key[L] = ((unsigned long long)(j - 1) << 32) | (unsigned long long)(i - 1);
    val[L] = sign * v;
}

// sorted keys -> row indices, values as fp32/fp64, duplicate detection
template <typename VT>
__global__ void __launch_bounds__(kBlock)
mtx_unpack_kernel(int64_t nnz, const unsigned long long *__restrict__ key,
                  const double *__restrict__ val, int32_t *__restrict__ rowidx,
                  VT *__restrict__ out, unsigned *__restrict__ bad) {
    const int64_t t = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (t >= nnz) return;
    const unsigned long long k = key[t];
    if (t > 0 && key[t - 1] == k) bad[1] = 1u;  // the same (row, col) twice
    rowidx[t] = (int32_t)(k & 0xffffffffull);
    out[t] = (VT)val[t];
    if (sizeof(VT) == 4 && (double)(float)val[t] != val[t]) bad[2] = 1u;  // not exact in fp32
}

// colptr[j] = first sorted position whose column is >= j, j in [0, m]
__global__ void __launch_bounds__(kBlock)
mtx_colptr_kernel(int64_t m, int64_t nnz, const unsigned long long *__restrict__ key,
                  int64_t *__restrict__ colptr) {
    const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (j > m) return;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)(key[mid] >> 32) < j) lo = mid + 1; else hi = mid;
    }
    colptr[j] = lo;
}

}  // namespace vb
