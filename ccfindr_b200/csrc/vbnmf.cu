// libvbnmf: host side of the B200-native VB Poisson-NMF engine and its C ABI (include/vbnmf.h).
//
// Replaces, for ccfindR's vb_factorize() path, the Rcpp step src/vbnmf_update.cpp:16-102 and the
// iteration loop R/bayesian.R:336-352 that calls it (plus hyper_update, R/bayesian.R:2-53), and
// for factorize() the loop R/factorize.R:189-212.  No CPU fallback: every entry point needs the
// CUDA device the handle was created on.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "../../include/vbnmf.h"
#include "kernels_common.cuh"
#include "rp_ranks.h"
#include "rp_table.h"

namespace vb {
#define F(RP) extern const RpTable rp_table_##RP;
VB_RP_LIST(F)
#undef F
const RpTable *rp_table(int rp) {
    switch (rp) {
#define F(RP) \
    case RP:  \
        return &rp_table_##RP;
        VB_RP_LIST(F)
#undef F
    }
    return nullptr;
}
}  // namespace vb

namespace {

constexpr int kMaxRank = 64;
constexpr int64_t kRowChunk = 4096;  // nonzeros per row-sweep work item

std::string g_create_error;

struct NcclApi {
    void *lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool load(std::string &err) {
        if (lib) return true;
        // libnccl.so.2 resolves to the copy already mapped by the host process (e.g. torch's) if any
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) {
            err = "libnccl lacks a required symbol";
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;

}  // namespace

struct vbnmf_comm {
    int nranks = 1, rank = 0, device = 0;
    ncclComm_t comm = nullptr;
};

struct vbnmf_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    int64_t n = 0, m = 0, nnz = 0, m_global = 0;
    int r = 0, rs = 0;
    int precision = VBNMF_FP64;
    bool val_float = true;
    bool borrowed = false;
    // CSC (cell columns) and CSR mirror (gene rows)
    int64_t *d_colptr = nullptr;
    int32_t *d_rowidx = nullptr;
    void *d_val = nullptr;
    int32_t *d_colidx = nullptr;
    void *d_valr = nullptr;
    // row-sweep work items
    int64_t n_items = 0;
    int32_t *d_item_row = nullptr, *d_item_len = nullptr;
    int64_t *d_item_beg = nullptr, *d_row_item_ptr = nullptr;
    // panels, n x rs and m x rs row-major (rank index fastest)
    double *d_lw = nullptr, *d_lh = nullptr, *d_alw = nullptr, *d_alh = nullptr;
    double *d_red = nullptr;  // [SwRaw n*rs | ehsum rs | hprior, sumloglh, sumeh | xlogp, enth | pad]
    double *d_ShRaw = nullptr, *d_SwPart = nullptr, *d_colx = nullptr, *d_cole = nullptr;
    double *d_scal = nullptr;  // [ewsum rs | wprior, sumloglw, sumew | entw | pad]
    double *h_scal = nullptr;  // pinned mirror: [d_scal (rs+8) | tail of d_red (rs+8)]
    double *d_partW = nullptr, *d_partH = nullptr, *d_partC = nullptr, *d_partE = nullptr;
    unsigned *d_counters = nullptr;
    unsigned long long *d_work = nullptr;
    int gridC = 0, gridE = 0;
    const vb::RpTable *tab = nullptr;
    int grid_cols = 0, grid_rows = 0;  // persistent sweep grids for the current rank
    double lgx = 0.0, mlconst = 0.0;  // global sums over nonzeros
    // host copies of the small vectors
    double ehsum[kMaxRank], ewsum[kMaxRank], bew[kMaxRank], beh[kMaxRank];
    double wacc[3], hacc[3], xlogp = 0, enth = 0, entw = 0;
    bool stats_valid = false, has_posterior = false;
    // NCCL
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    int64_t launches = 0;
    std::string err;
};

namespace {

using H = vbnmf_handle;

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                 \
            return VBNMF_ERR_CUDA;                                                       \
        }                                                                                \
    } while (0)

#define CKN(call)                                                                        \
    do {                                                                                 \
        ncclResult_t e_ = (call);                                                        \
        if (e_ != ncclSuccess) {                                                         \
            h->err = std::string(#call) + ": " +                                         \
                     (g_nccl.GetErrorString ? g_nccl.GetErrorString(e_) : "nccl error"); \
            return VBNMF_ERR_NCCL;                                                       \
        }                                                                                \
    } while (0)

int fail(H *h, int code, const std::string &msg) {
    h->err = msg;
    return code;
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline int pad_rank(int r) {
    int rs = (r + 1) & ~1;
    if (rs < 2) rs = 2;
    if (rs > 32) rs = (rs + 7) & ~7;
    return rs;
}
inline int64_t tail_off(const H *h) { return h->n * h->rs; }
inline int64_t red_len(const H *h) { return h->n * h->rs + h->rs + 8; }

int launch_sweep_cols(H *h) {
    CK(cudaMemsetAsync(h->d_work, 0, 2 * sizeof(unsigned long long), h->stream));
    double *tail = h->d_red + tail_off(h);
    vb::SweepColsArgs a{h->m, h->r, h->d_colptr, h->d_rowidx, h->d_val, h->d_lw, h->d_lh,
                        h->d_ShRaw, h->d_colx, h->d_cole, h->d_work};
    h->tab->sweep_cols(a, h->val_float, h->grid_cols, h->stream);
    vb::reduce_cols_kernel<<<h->gridC, vb::kBlock, 0, h->stream>>>(
        h->m, h->d_colx, h->d_cole, h->d_partC, tail + h->rs + 3, h->d_counters + 2);
    h->launches += 2;
    return 0;
}

int launch_sweep_rows(H *h) {
    vb::SweepRowsArgs a{h->n_items, h->d_item_row, h->d_item_beg, h->d_item_len, h->d_colidx,
                        h->d_valr, h->d_lw, h->d_lh, h->d_SwPart, h->d_work + 1};
    h->tab->sweep_rows(a, h->val_float, h->grid_rows, h->stream);
    vb::combine_rows_kernel<<<cdiv(h->n * h->rs, vb::kBlock), vb::kBlock, 0, h->stream>>>(
        h->n, h->rs, h->d_row_item_ptr, h->d_SwPart, h->d_red);
    h->launches += 2;
    return 0;
}

int launch_posterior(H *h, bool wside, double a, double b, double fud) {
    double *tail = h->d_red + tail_off(h);
    vb::PosteriorArgs p;
    p.rows = wside ? h->n : h->m;
    p.r = h->r;
    p.a = a; p.b = b; p.fud = fud;
    p.osum = wside ? tail : h->d_scal;
    p.SRaw = wside ? h->d_red : h->d_ShRaw;
    p.l = wside ? h->d_lw : h->d_lh;
    p.al_out = wside ? h->d_alw : h->d_alh;
    p.part = wside ? h->d_partW : h->d_partH;
    p.out = wside ? h->d_scal : tail;
    p.counter = h->d_counters + (wside ? 0 : 1);
    h->tab->posterior(p, h->stream);
    h->launches += 1;
    return 0;
}

int allreduce(H *h, double *buf, int64_t count) {
    if (h->nranks <= 1) return 0;
    CKN(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, h->comm, h->stream));
    h->launches += 1;
    return 0;
}

// statistics at the current lw, lh: ShRaw, SwRaw (global), xlogp, enth, and entw
int sweep(H *h) {
    int rc;
    if ((rc = launch_sweep_cols(h))) return rc;
    if ((rc = launch_sweep_rows(h))) return rc;
    if ((rc = allreduce(h, h->d_red, red_len(h)))) return rc;
    vb::entropy_w_kernel<<<h->gridE, vb::kBlock, 0, h->stream>>>(
        h->n, h->rs, h->r, h->d_lw, h->d_red, h->d_partE, h->d_scal + h->rs + 3,
        h->d_counters + 3);
    h->launches += 1;
    CK(cudaGetLastError());
    h->stats_valid = true;
    return 0;
}

int fetch_scalars(H *h) {
    const int w = h->rs + 8;
    CK(cudaMemcpyAsync(h->h_scal, h->d_scal, w * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(h->h_scal + w, h->d_red + tail_off(h), w * sizeof(double),
                       cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

// host half of an iteration: take the freshly fetched scalars, refresh the small host vectors and
// assemble the lower bound (src/vbnmf_update.cpp:67-90 in its nonzero-only form)
double absorb_scalars(H *h, const double *hyper) {
    const double aw = hyper[0], bw = hyper[1], ah = hyper[2], bh = hyper[3];
    const int rs = h->rs, r = h->r, w = rs + 8;
    const double *ws = h->h_scal, *hs = h->h_scal + w;
    for (int k = 0; k < r; k++) {
        h->bew[k] = aw / bw + h->ehsum[k];  // rowSums(eh_old), :42-43
        h->ewsum[k] = ws[k];
        h->beh[k] = ah / bh + h->ewsum[k];  // colSums(ew_new), :52-53
    }
    for (int k = 0; k < r; k++) h->ehsum[k] = hs[k];
    for (int c = 0; c < 3; c++) { h->wacc[c] = ws[rs + c]; h->hacc[c] = hs[rs + c]; }
    h->entw = ws[rs + 3];
    h->xlogp = hs[rs + 3];
    h->enth = hs[rs + 4];
    const double nr = (double)h->n * r, mr = (double)h->m_global * r;
    double U = 0.0;
    for (int k = 0; k < r; k++) U -= h->ewsum[k] * h->ehsum[k];       // -sum(ew.eh), :78
    U -= h->entw + h->enth - h->xlogp;                                // -x((A+B)/wth - log wth), :74-78
    U -= h->lgx;                                                      // -lgamma(x+1), :81
    U += h->wacc[0] + nr * (-lgamma(aw) + aw * log(aw / bw));         // :82-86
    U += h->hacc[0] + mr * (-lgamma(ah) + ah * log(ah / bh));         // :87-89
    h->has_posterior = true;
    return U / ((double)h->n * (double)h->m_global);                  // :90 (in double, not int)
}

// One reference iteration given valid statistics: posterior update (W then H), sweep at the new
// factors, lower bound.  src/vbnmf_update.cpp:38-90.
int iterate(H *h, const double *hyper, double fud, double *lkh) {
    int rc;
    if ((rc = launch_posterior(h, true, hyper[0], hyper[1], fud))) return rc;
    if ((rc = launch_posterior(h, false, hyper[2], hyper[3], fud))) return rc;
    if ((rc = sweep(h))) return rc;
    if ((rc = fetch_scalars(h))) return rc;
    *lkh = absorb_scalars(h, hyper);
    return 0;
}

void means_of(const H *h, double *means) {
    const double nr = (double)h->n * h->r, mr = (double)h->m_global * h->r;
    means[0] = h->wacc[1] / nr;
    means[1] = h->hacc[1] / mr;
    means[2] = h->wacc[2] / nr;
    means[3] = h->hacc[2] / mr;
}

// R/bayesian.R:2-53
int hyper_update(const int *flags, const double *mn, double *hyper, int niter, double tol) {
    if (flags[0] + flags[1] + flags[2] + flags[3] == 0) return 0;
    const double lwm = mn[0], lhm = mn[1], ewm = mn[2], ehm = mn[3];
    double aw0 = hyper[0], ah0 = hyper[2];
    const double bw0 = hyper[1], bh0 = hyper[3];
    double aw1 = aw0, ah1 = ah0;
    if (flags[0] + flags[2] > 0) {
        int i = 1;
        while (i < niter) {
            double dw = 0.0, dh = 0.0;
            if (flags[0])
                dw = (log(aw0) - vb_digamma(aw0) - ewm / bw0 + 1.0 + lwm - log(bw0)) /
                     (1.0 / aw0 - vb_trigamma(aw0));
            if (flags[2])
                dh = (log(ah0) - vb_digamma(ah0) - ehm / bh0 + 1.0 + lhm - log(bh0)) /
                     (1.0 / ah0 - vb_trigamma(ah0));
            aw1 = aw0 - dw;
            ah1 = ah0 - dh;
            while (aw1 <= 0) { dw = dw / 2; aw1 = aw0 - dw; }
            while (ah1 <= 0) { dh = dh / 2; ah1 = ah0 - dh; }
            const double df =
                (1 - aw1 / aw0) * (1 - aw1 / aw0) + (1 - ah1 / ah0) * (1 - ah1 / ah0);
            if (df < tol) break;
            aw0 = aw1;
            ah0 = ah1;
            i++;
        }
        if (i == niter) return VBNMF_ERR_HYPER;
    }
    hyper[0] = aw1;
    hyper[1] = flags[1] ? ewm : bw0;
    hyper[2] = ah1;
    hyper[3] = ehm;  // R/bayesian.R:50-51 assigns ehm in both branches
    return 0;
}

void free_panels(H *h) {
    double **ps[] = {&h->d_lw, &h->d_lh, &h->d_alw, &h->d_alh, &h->d_red, &h->d_ShRaw,
                     &h->d_SwPart, &h->d_scal, &h->d_partW, &h->d_partH};
    for (auto p : ps) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    if (h->h_scal) cudaFreeHost(h->h_scal);
    h->h_scal = nullptr;
}

int alloc_panels(H *h, int r) {
    const int rs = pad_rank(r);
    if (rs == h->rs && h->d_lw) { h->r = r; return 0; }
    const vb::RpTable *tab = vb::rp_table(rs);
    if (!tab) return fail(h, VBNMF_ERR_ARG, "no kernels compiled for this rank");
    free_panels(h);
    h->tab = tab;
    h->grid_cols = tab->sweep_cols_ctas_per_sm(h->val_float) * h->num_sms;
    h->grid_rows = tab->sweep_rows_ctas_per_sm(h->val_float) * h->num_sms;
    h->r = r;
    h->rs = rs;
    const size_t nr = (size_t)h->n * rs * 8, mr = (size_t)h->m * rs * 8;
    CK(cudaMalloc(&h->d_lw, nr));
    CK(cudaMalloc(&h->d_alw, nr));
    CK(cudaMalloc(&h->d_lh, mr));
    CK(cudaMalloc(&h->d_alh, mr));
    CK(cudaMalloc(&h->d_ShRaw, mr));
    CK(cudaMalloc(&h->d_red, (size_t)red_len(h) * 8));
    CK(cudaMalloc(&h->d_SwPart, (size_t)h->n_items * rs * 8));
    CK(cudaMalloc(&h->d_scal, (size_t)(rs + 8) * 8));
    CK(cudaMalloc(&h->d_partW, (size_t)cdiv(h->n, vb::kBlock) * (rs + 3) * 8));
    CK(cudaMalloc(&h->d_partH, (size_t)cdiv(h->m, vb::kBlock) * (rs + 3) * 8));
    CK(cudaMallocHost(&h->h_scal, (size_t)2 * (rs + 8) * 8));
    CK(cudaMemsetAsync(h->d_red, 0, (size_t)red_len(h) * 8, h->stream));
    CK(cudaMemsetAsync(h->d_scal, 0, (size_t)(rs + 8) * 8, h->stream));
    return 0;
}

// host n x r column-major -> device n x rs row-major (zero padded)
int upload_cm(H *h, double *dst, const double *src, int64_t rows, int r) {
    const int rs = h->rs;
    std::vector<double> tmp((size_t)rows * rs, 0.0);
    for (int k = 0; k < r; k++)
        for (int64_t i = 0; i < rows; i++) tmp[(size_t)i * rs + k] = src[(size_t)k * rows + i];
    CK(cudaMemcpyAsync(dst, tmp.data(), tmp.size() * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
// host r x m column-major (= m rows of r) -> device m x rs
int upload_rm(H *h, double *dst, const double *src, int64_t rows, int r) {
    const int rs = h->rs;
    if (rs != r) CK(cudaMemsetAsync(dst, 0, (size_t)rows * rs * 8, h->stream));
    CK(cudaMemcpy2DAsync(dst, (size_t)rs * 8, src, (size_t)r * 8, (size_t)r * 8, (size_t)rows,
                         cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

template <typename VT>
int build_layouts_t(H *h) {
    const int64_t n = h->n, m = h->m, nnz = h->nnz;
    if (nnz >= (int64_t)UINT32_MAX) return fail(h, VBNMF_ERR_ARG, "nnz per GPU must be < 2^32-1");
    const VT *val = (const VT *)h->d_val;
    // constants over the nonzeros
    const int gridK = h->num_sms * 8;
    double *d_part = nullptr, *d_out = nullptr;
    CK(cudaMalloc(&d_part, (size_t)gridK * 2 * 8));
    CK(cudaMalloc(&d_out, 2 * 8));
    vb::count_constants_kernel<VT><<<gridK, vb::kBlock, 0, h->stream>>>(nnz, val, d_part, d_out,
                                                                       h->d_counters + 4);
    double consts[2];
    CK(cudaMemcpyAsync(consts, d_out, 16, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->lgx = consts[0];
    h->mlconst = consts[1];
    cudaFree(d_part);
    cudaFree(d_out);
    // CSR mirror: stable sort of the nonzeros by gene row
    int32_t *d_colof = nullptr, *d_keys_out = nullptr;
    uint32_t *d_perm_in = nullptr, *d_perm_out = nullptr;
    unsigned long long *d_rowcount = nullptr;
    CK(cudaMalloc(&d_colof, (size_t)nnz * 4));
    CK(cudaMalloc(&d_rowcount, (size_t)(n + 1) * 8));
    CK(cudaMemsetAsync(d_rowcount, 0, (size_t)(n + 1) * 8, h->stream));
    vb::expand_cols_kernel<<<h->num_sms * 8, vb::kBlock, 0, h->stream>>>(m, h->d_colptr, h->d_rowidx,
                                                                        d_colof, d_rowcount);
    CK(cudaMalloc(&d_keys_out, (size_t)nnz * 4));
    CK(cudaMalloc(&d_perm_in, (size_t)nnz * 4));
    CK(cudaMalloc(&d_perm_out, (size_t)nnz * 4));
    vb::iota_kernel<<<h->num_sms * 8, vb::kBlock, 0, h->stream>>>(nnz, d_perm_in);
    int bits = 1;
    while (((int64_t)1 << bits) < n) bits++;
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->d_rowidx, d_keys_out, d_perm_in,
                                       d_perm_out, nnz, 0, bits, h->stream));
    void *d_tmp = nullptr;
    CK(cudaMalloc(&d_tmp, tmp_bytes));
    CK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, h->d_rowidx, d_keys_out, d_perm_in,
                                       d_perm_out, nnz, 0, bits, h->stream));
    CK(cudaMalloc(&h->d_colidx, (size_t)nnz * 4));
    CK(cudaMalloc(&h->d_valr, (size_t)nnz * sizeof(VT)));
    vb::gather_csr_kernel<VT><<<h->num_sms * 8, vb::kBlock, 0, h->stream>>>(
        nnz, d_perm_out, d_colof, val, h->d_colidx, (VT *)h->d_valr);
    // row pointers on the host -> work items
    std::vector<unsigned long long> rc((size_t)n + 1);
    CK(cudaMemcpyAsync(rc.data(), d_rowcount, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost,
                       h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    cudaFree(d_tmp); cudaFree(d_colof); cudaFree(d_keys_out); cudaFree(d_perm_in);
    cudaFree(d_perm_out); cudaFree(d_rowcount);
    std::vector<int32_t> item_row, item_len;
    std::vector<int64_t> item_beg, row_item_ptr((size_t)n + 1);
    int64_t off = 0;
    for (int64_t i = 0; i < n; i++) {
        row_item_ptr[i] = (int64_t)item_row.size();
        const int64_t cnt = (int64_t)rc[i];
        for (int64_t b = 0; b < cnt; b += kRowChunk) {
            item_row.push_back((int32_t)i);
            item_beg.push_back(off + b);
            item_len.push_back((int32_t)std::min(kRowChunk, cnt - b));
        }
        off += cnt;
    }
    row_item_ptr[n] = (int64_t)item_row.size();
    h->n_items = (int64_t)item_row.size();
    const size_t ni = std::max<size_t>(1, item_row.size());
    CK(cudaMalloc(&h->d_item_row, ni * 4));
    CK(cudaMalloc(&h->d_item_len, ni * 4));
    CK(cudaMalloc(&h->d_item_beg, ni * 8));
    CK(cudaMalloc(&h->d_row_item_ptr, (size_t)(n + 1) * 8));
    CK(cudaMemcpy(h->d_item_row, item_row.data(), item_row.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_item_len, item_len.data(), item_len.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_item_beg, item_beg.data(), item_beg.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_row_item_ptr, row_item_ptr.data(), (size_t)(n + 1) * 8,
                  cudaMemcpyHostToDevice));
    return 0;
}

int init_common(H *h, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(h, VBNMF_ERR_CUDA, "no CUDA device available (libvbnmf has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(h, VBNMF_ERR_ARG, "bad device ordinal");
    h->device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    CK(cudaMalloc(&h->d_counters, 16 * sizeof(unsigned)));
    CK(cudaMemset(h->d_counters, 0, 16 * sizeof(unsigned)));
    CK(cudaMalloc(&h->d_work, 2 * sizeof(unsigned long long)));
    h->gridC = std::min<int64_t>(h->num_sms * 4, std::max<int64_t>(1, cdiv(h->m, vb::kBlock)));
    h->gridE = std::min<int64_t>(h->num_sms * 4, std::max<int64_t>(1, cdiv(h->n * 2, vb::kBlock)));
    CK(cudaMalloc(&h->d_colx, (size_t)std::max<int64_t>(1, h->m) * 8));
    CK(cudaMalloc(&h->d_cole, (size_t)std::max<int64_t>(1, h->m) * 8));
    CK(cudaMalloc(&h->d_partC, (size_t)h->gridC * 2 * 8));
    CK(cudaMalloc(&h->d_partE, (size_t)h->gridE * 8));
    h->m_global = h->m;
    return 0;
}

}  // namespace

extern "C" {

const char *vbnmf_last_error(const vbnmf_handle *h) {
    return h ? h->err.c_str() : g_create_error.c_str();
}

void vbnmf_destroy(vbnmf_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_panels(h);
    if (!h->borrowed) {
        cudaFree(h->d_colptr); cudaFree(h->d_rowidx); cudaFree(h->d_val);
    }
    cudaFree(h->d_colidx); cudaFree(h->d_valr);
    cudaFree(h->d_item_row); cudaFree(h->d_item_len); cudaFree(h->d_item_beg);
    cudaFree(h->d_row_item_ptr);
    cudaFree(h->d_colx); cudaFree(h->d_cole); cudaFree(h->d_partC); cudaFree(h->d_partE);
    cudaFree(h->d_counters); cudaFree(h->d_work);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int vbnmf_create(vbnmf_handle **out, int64_t n, int64_t m, int64_t nnz, const int32_t *colptr32,
                 const int64_t *colptr64, const int32_t *rowidx, const double *values, int device) {
    if (!out) return VBNMF_ERR_ARG;
    *out = nullptr;
    H *h = new H();
    auto bail = [&](int rc) {
        g_create_error = h->err;
        vbnmf_destroy(h);
        return rc;
    };
    if (n <= 0 || m <= 0 || nnz <= 0 || (!colptr32 == !colptr64) || !rowidx || !values)
        return bail(fail(h, VBNMF_ERR_ARG, "vbnmf_create: bad arguments"));
    h->n = n; h->m = m; h->nnz = nnz;
    int rc = init_common(h, device);
    if (rc) return bail(rc);
    std::vector<int64_t> cp((size_t)m + 1);
    for (int64_t j = 0; j <= m; j++) cp[j] = colptr64 ? colptr64[j] : (int64_t)colptr32[j];
    if (cp[0] != 0 || cp[m] != nnz) return bail(fail(h, VBNMF_ERR_ARG, "colptr does not span nnz"));
    bool as_float = true;
    for (int64_t t = 0; t < nnz; t++) {
        if (rowidx[t] < 0 || rowidx[t] >= n)
            return bail(fail(h, VBNMF_ERR_ARG, "row index out of range"));
        if ((double)(float)values[t] != values[t]) as_float = false;
    }
    h->val_float = as_float;
    auto up = [&]() -> int {
        CK(cudaMalloc(&h->d_colptr, (size_t)(m + 1) * 8));
        CK(cudaMalloc(&h->d_rowidx, (size_t)nnz * 4));
        CK(cudaMemcpy(h->d_colptr, cp.data(), (size_t)(m + 1) * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_rowidx, rowidx, (size_t)nnz * 4, cudaMemcpyHostToDevice));
        if (as_float) {
            std::vector<float> vf((size_t)nnz);
            for (int64_t t = 0; t < nnz; t++) vf[t] = (float)values[t];
            CK(cudaMalloc(&h->d_val, (size_t)nnz * 4));
            CK(cudaMemcpy(h->d_val, vf.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice));
        } else {
            CK(cudaMalloc(&h->d_val, (size_t)nnz * 8));
            CK(cudaMemcpy(h->d_val, values, (size_t)nnz * 8, cudaMemcpyHostToDevice));
        }
        return 0;
    };
    if ((rc = up())) return bail(rc);
    rc = as_float ? build_layouts_t<float>(h) : build_layouts_t<double>(h);
    if (rc) return bail(rc);
    *out = h;
    return 0;
}

int vbnmf_create_from_device(vbnmf_handle **out, int64_t n, int64_t m, int64_t nnz,
                             const int64_t *d_colptr, const int32_t *d_rowidx, const float *d_values,
                             int device) {
    if (!out) return VBNMF_ERR_ARG;
    *out = nullptr;
    H *h = new H();
    auto bail = [&](int rc) {
        g_create_error = h->err;
        vbnmf_destroy(h);
        return rc;
    };
    if (n <= 0 || m <= 0 || nnz <= 0 || !d_colptr || !d_rowidx || !d_values)
        return bail(fail(h, VBNMF_ERR_ARG, "vbnmf_create_from_device: bad arguments"));
    h->n = n; h->m = m; h->nnz = nnz;
    int rc = init_common(h, device);
    if (rc) return bail(rc);
    h->borrowed = true;
    h->val_float = true;
    h->d_colptr = const_cast<int64_t *>(d_colptr);
    h->d_rowidx = const_cast<int32_t *>(d_rowidx);
    h->d_val = const_cast<float *>(d_values);
    if ((rc = build_layouts_t<float>(h))) return bail(rc);
    *out = h;
    return 0;
}

int vbnmf_set_precision(vbnmf_handle *h, int precision) {
    if (!h) return VBNMF_ERR_ARG;
    if (precision != VBNMF_FP64)
        return fail(h, VBNMF_ERR_ARG, "only VBNMF_FP64 is implemented in this build");
    h->precision = precision;
    return 0;
}

int vbnmf_set_stream(vbnmf_handle *h, void *cuda_stream) {
    if (!h) return VBNMF_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->own_stream = false;
    return 0;
}

int vbnmf_nccl_unique_id(void *uid128) {
    std::string err;
    if (!uid128 || !g_nccl.load(err)) { g_create_error = err; return VBNMF_ERR_NCCL; }
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return VBNMF_ERR_NCCL;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(uid128, &id, 128);
    return 0;
}

int vbnmf_comm_create(vbnmf_comm **out, int nranks, int rank, const void *uid128, int device) {
    if (!out || nranks < 1 || rank < 0 || rank >= nranks || !uid128) return VBNMF_ERR_ARG;
    *out = nullptr;
    if (!g_nccl.load(g_create_error)) return VBNMF_ERR_NCCL;
    if (cudaSetDevice(device) != cudaSuccess) {
        g_create_error = "vbnmf_comm_create: bad device";
        return VBNMF_ERR_CUDA;
    }
    vbnmf_comm *c = new vbnmf_comm();
    c->nranks = nranks; c->rank = rank; c->device = device;
    ncclUniqueId id;
    memcpy(&id, uid128, 128);
    ncclResult_t e = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
    if (e != ncclSuccess) {
        g_create_error = std::string("ncclCommInitRank: ") +
                         (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error");
        delete c;
        return VBNMF_ERR_NCCL;
    }
    *out = c;
    return 0;
}

void vbnmf_comm_destroy(vbnmf_comm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
}

int vbnmf_attach_comm(vbnmf_handle *h, vbnmf_comm *c) {
    if (!h || !c) return VBNMF_ERR_ARG;
    if (c->device != h->device) return fail(h, VBNMF_ERR_ARG, "communicator is on another device");
    if (h->nranks > 1) return fail(h, VBNMF_ERR_STATE, "a communicator is already attached");
    if (c->nranks == 1) return 0;
    CK(cudaSetDevice(h->device));
    h->comm = c->comm;
    h->nranks = c->nranks;
    h->rank = c->rank;
    // global constants: total cells and the sums over nonzeros
    double *d3 = nullptr, h3[3] = {(double)h->m, h->lgx, h->mlconst};
    CK(cudaMalloc(&d3, 24));
    CK(cudaMemcpyAsync(d3, h3, 24, cudaMemcpyHostToDevice, h->stream));
    int rc = allreduce(h, d3, 3);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h3, d3, 24, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d3);
    h->m_global = (int64_t)llround(h3[0]);
    h->lgx = h3[1];
    h->mlconst = h3[2];
    h->stats_valid = false;
    return 0;
}

int vbnmf_set_state(vbnmf_handle *h, int r, const double *lw, const double *lh, const double *ew,
                    const double *eh) {
    if (!h || !lw || !lh) return VBNMF_ERR_ARG;
    if (r < 1 || r > kMaxRank) return fail(h, VBNMF_ERR_ARG, "rank must be in 1..64");
    CK(cudaSetDevice(h->device));
    int rc;
    if ((rc = alloc_panels(h, r))) return rc;
    if ((rc = upload_cm(h, h->d_lw, lw, h->n, r))) return rc;
    if ((rc = upload_rm(h, h->d_lh, lh, h->m, r))) return rc;
    // before the first update ew/eh are whatever the caller holds (vb_init: ew = w, eh = h)
    if ((rc = upload_cm(h, h->d_alw, ew ? ew : lw, h->n, r))) return rc;
    if ((rc = upload_rm(h, h->d_alh, eh ? eh : lh, h->m, r))) return rc;
    const double *e = eh ? eh : lh;
    double *tail = h->d_red + tail_off(h);
    std::vector<double> s((size_t)h->rs + 8, 0.0);
    for (int64_t j = 0; j < h->m; j++)
        for (int k = 0; k < r; k++) s[k] += e[(size_t)j * r + k];
    CK(cudaMemcpyAsync(tail, s.data(), s.size() * 8, cudaMemcpyHostToDevice, h->stream));
    if ((rc = allreduce(h, tail, h->rs + 8))) return rc;
    CK(cudaMemcpyAsync(s.data(), tail, s.size() * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < kMaxRank; k++) {
        h->ehsum[k] = k < r ? s[k] : 0.0;
        h->ewsum[k] = 0.0;
        h->bew[k] = h->beh[k] = 1.0;
    }
    h->stats_valid = false;
    h->has_posterior = false;
    return 0;
}

int vbnmf_step(vbnmf_handle *h, const double hyper[4], double fudge, double *lkh) {
    if (!h || !hyper || !lkh) return VBNMF_ERR_ARG;
    if (!h->d_lw) return fail(h, VBNMF_ERR_STATE, "vbnmf_set_state has not been called");
    CK(cudaSetDevice(h->device));
    int rc;
    if (!h->stats_valid) {
        // the statistics pass must not disturb the H-side sums already in the reduce buffer
        std::vector<double> keep((size_t)h->rs + 8, 0.0);
        for (int k = 0; k < h->r; k++) keep[k] = h->ehsum[k];
        if ((rc = sweep(h))) return rc;
        CK(cudaMemcpyAsync(h->d_red + tail_off(h), keep.data(), (size_t)h->rs * 8,
                           cudaMemcpyHostToDevice, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return iterate(h, hyper, fudge, lkh);
}

int vbnmf_run(vbnmf_handle *h, const vbnmf_cfg *cfg, double hyper[4], double *lkh_trace,
              double *hyper_trace, int *niter, double *lml, int *stop_reason) {
    if (!h || !cfg || !hyper || !niter || !lml || !stop_reason) return VBNMF_ERR_ARG;
    if (!h->d_lw) return fail(h, VBNMF_ERR_STATE, "vbnmf_set_state has not been called");
    if (cfg->itmax < 1 || cfg->dn < 1) return fail(h, VBNMF_ERR_ARG, "itmax and dn must be >= 1");
    double lk0 = 0.0, lkh = 0.0;  // R/bayesian.R:336
    int it, reason = VBNMF_STOP_ITMAX, rc = 0;
    for (it = 1; it <= cfg->itmax; it++) {                                  // :337
        if ((rc = vbnmf_step(h, hyper, cfg->fudge, &lkh))) return rc;       // :339
        if (it > cfg->n0 && it % cfg->dn == 0) {                            // :342-344
            double mn[4];
            means_of(h, mn);
            if ((rc = hyper_update(cfg->hyper_update, mn, hyper, 100, 1e-3)))
                return fail(h, rc, "Hyper-parameter update failed to converge");
        }
        if (lkh_trace) lkh_trace[it - 1] = lkh;
        if (hyper_trace) memcpy(hyper_trace + 4 * (it - 1), hyper, 4 * sizeof(double));
        if (std::isnan(lkh)) { reason = VBNMF_STOP_NAN; break; }            // :345
        if (it > 1 && it > cfg->n0 && lkh >= lk0 && fabs(1 - lkh / lk0) < cfg->tol) {  // :346-347
            reason = VBNMF_STOP_CONVERGED;
            break;
        }
        lk0 = lkh;                                                          // :348
    }
    if (it > cfg->itmax) it = cfg->itmax;
    *niter = it;
    *lml = lk0;                                                             // :379
    *stop_reason = reason;
    return 0;
}

int vbnmf_get_means(vbnmf_handle *h, double means[4]) {
    if (!h || !means) return VBNMF_ERR_ARG;
    if (!h->has_posterior) return fail(h, VBNMF_ERR_STATE, "no update has been run yet");
    means_of(h, means);
    return 0;
}

int vbnmf_get_state(vbnmf_handle *h, double *lw, double *lh, double *ew, double *eh, double *dw,
                    double *dh) {
    if (!h) return VBNMF_ERR_ARG;
    if (!h->d_lw) return fail(h, VBNMF_ERR_STATE, "vbnmf_set_state has not been called");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    const int r = h->r, rs = h->rs;
    const int64_t n = h->n, m = h->m;
    if (lw || ew || dw) {
        std::vector<double> tmp((size_t)n * rs);
        if (lw) {
            CK(cudaMemcpy(tmp.data(), h->d_lw, tmp.size() * 8, cudaMemcpyDeviceToHost));
            for (int k = 0; k < r; k++)
                for (int64_t i = 0; i < n; i++) lw[(size_t)k * n + i] = tmp[(size_t)i * rs + k];
        }
        if (ew || dw) {
            CK(cudaMemcpy(tmp.data(), h->d_alw, tmp.size() * 8, cudaMemcpyDeviceToHost));
            for (int k = 0; k < r; k++)
                for (int64_t i = 0; i < n; i++) {
                    const double a = tmp[(size_t)i * rs + k], b = h->bew[k];
                    if (ew) ew[(size_t)k * n + i] = a / b;                      // :44
                    if (dw) dw[(size_t)k * n + i] = h->has_posterior ? a / b / b : 0.0;  // :46
                }
        }
    }
    if (lh)
        CK(cudaMemcpy2D(lh, (size_t)r * 8, h->d_lh, (size_t)rs * 8, (size_t)r * 8, (size_t)m,
                        cudaMemcpyDeviceToHost));
    if (eh || dh) {
        std::vector<double> tmp((size_t)m * r);
        CK(cudaMemcpy2D(tmp.data(), (size_t)r * 8, h->d_alh, (size_t)rs * 8, (size_t)r * 8,
                        (size_t)m, cudaMemcpyDeviceToHost));
        for (int64_t j = 0; j < m; j++)
            for (int k = 0; k < r; k++) {
                const double a = tmp[(size_t)j * r + k], b = h->beh[k];
                if (eh) eh[(size_t)j * r + k] = a / b;                          // :54
                if (dh) dh[(size_t)j * r + k] = h->has_posterior ? a / b / b : 0.0;  // :56
            }
    }
    return 0;
}

int vbnmf_cluster_id(vbnmf_handle *h, int32_t *cid) {
    if (!h || !cid) return VBNMF_ERR_ARG;
    if (!h->d_lw) return fail(h, VBNMF_ERR_STATE, "vbnmf_set_state has not been called");
    CK(cudaSetDevice(h->device));
    double *d_beh = nullptr;
    int32_t *d_cid = nullptr;
    CK(cudaMalloc(&d_beh, kMaxRank * 8));
    CK(cudaMalloc(&d_cid, (size_t)h->m * 4));
    CK(cudaMemcpyAsync(d_beh, h->beh, kMaxRank * 8, cudaMemcpyHostToDevice, h->stream));
    vb::cluster_id_kernel<<<cdiv(h->m, vb::kBlock), vb::kBlock, 0, h->stream>>>(
        h->m, h->rs, h->r, h->d_alh, d_beh, d_cid);
    CK(cudaMemcpyAsync(cid, d_cid, (size_t)h->m * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_beh);
    cudaFree(d_cid);
    return 0;
}

int vbnmf_uniform_columns(vbnmf_handle *h, double tol, int32_t *flags) {
    if (!h || !flags) return VBNMF_ERR_ARG;
    if (!h->d_lw) return fail(h, VBNMF_ERR_STATE, "vbnmf_set_state has not been called");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    const int r = h->r, rs = h->rs;
    std::vector<double> tmp((size_t)h->n * rs);
    CK(cudaMemcpy(tmp.data(), h->d_alw, tmp.size() * 8, cudaMemcpyDeviceToHost));
    for (int k = 0; k < r; k++) {
        double mx = -INFINITY, mn = INFINITY;
        for (int64_t i = 0; i < h->n; i++) {
            const double v = tmp[(size_t)i * rs + k] / h->bew[k];
            mx = v > mx ? v : mx;
            mn = v < mn ? v : mn;
        }
        flags[k] = fabs(mx - mn) < tol ? 1 : 0;  // R/bayesian.R:368-369
    }
    return 0;
}

int vbnmf_bench_iterations(vbnmf_handle *h, const double hyper[4], double fudge, int iters,
                           double ms[4], int64_t *launches, double *lkh_last) {
    if (!h || !hyper || !ms || iters < 1) return VBNMF_ERR_ARG;
    if (!h->d_lw) return fail(h, VBNMF_ERR_STATE, "vbnmf_set_state has not been called");
    CK(cudaSetDevice(h->device));
    int rc;
    double lkh = 0.0;
    if (!h->stats_valid && (rc = vbnmf_step(h, hyper, fudge, &lkh))) return rc;
    std::vector<cudaEvent_t> ev((size_t)iters * 4 + 2);
    for (auto &e : ev) CK(cudaEventCreate(&e));
    const int64_t l0 = h->launches;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventRecord(ev[0], h->stream));
    const double aw = hyper[0], bw = hyper[1], ah = hyper[2], bh = hyper[3];
    for (int it = 0; it < iters; it++) {
        // same sequence as iterate(), with events around the two sweep kernels
        if ((rc = launch_posterior(h, true, aw, bw, fudge))) return rc;
        if ((rc = launch_posterior(h, false, ah, bh, fudge))) return rc;
        CK(cudaEventRecord(ev[2 + it * 4 + 0], h->stream));
        if ((rc = launch_sweep_cols(h))) return rc;
        CK(cudaEventRecord(ev[2 + it * 4 + 1], h->stream));
        CK(cudaEventRecord(ev[2 + it * 4 + 2], h->stream));
        if ((rc = launch_sweep_rows(h))) return rc;
        CK(cudaEventRecord(ev[2 + it * 4 + 3], h->stream));
        if ((rc = allreduce(h, h->d_red, red_len(h)))) return rc;
        vb::entropy_w_kernel<<<h->gridE, vb::kBlock, 0, h->stream>>>(
            h->n, h->rs, h->r, h->d_lw, h->d_red, h->d_partE, h->d_scal + h->rs + 3,
            h->d_counters + 3);
        h->launches += 1;
        if ((rc = fetch_scalars(h))) return rc;  // the per-iteration host readback of the loop
        lkh = absorb_scalars(h, hyper);
    }
    CK(cudaEventRecord(ev[1], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, ev[0], ev[1]));
    ms[0] = t;
    ms[1] = ms[2] = 0.0;
    for (int it = 0; it < iters; it++) {
        CK(cudaEventElapsedTime(&t, ev[2 + it * 4 + 0], ev[2 + it * 4 + 1]));
        ms[1] += t;
        CK(cudaEventElapsedTime(&t, ev[2 + it * 4 + 2], ev[2 + it * 4 + 3]));
        ms[2] += t;
    }
    ms[3] = ms[0] - ms[1] - ms[2];
    for (auto &e : ev) cudaEventDestroy(e);
    if (launches) *launches = h->launches - l0;
    if (lkh_last) *lkh_last = lkh;
    return 0;
}

int vbnmf_info(const vbnmf_handle *h, int64_t info[8]) {
    if (!h || !info) return VBNMF_ERR_ARG;
    info[0] = h->n; info[1] = h->m; info[2] = h->nnz; info[3] = h->r; info[4] = h->rs;
    info[5] = h->precision; info[6] = h->nranks; info[7] = h->m_global;
    return 0;
}

int mlnmf_run(vbnmf_handle *h, int r, const double *w0, const double *h0, int itmax, double tol,
              double *w, double *h_out, double *lik_trace, int *niter) {
    if (!h || !w0 || !h0 || !niter || itmax < 1) return VBNMF_ERR_ARG;
    if (r < 1 || r > kMaxRank) return fail(h, VBNMF_ERR_ARG, "rank must be in 1..64");
    CK(cudaSetDevice(h->device));
    int rc;
    if ((rc = alloc_panels(h, r))) return rc;
    if ((rc = upload_cm(h, h->d_lw, w0, h->n, r))) return rc;
    if ((rc = upload_rm(h, h->d_lh, h0, h->m, r))) return rc;
    h->stats_valid = false;
    h->has_posterior = false;
    const double eps = 2.220446049250313e-16;  // .Machine$double.eps, R/factorize.R:15,24
    const int rs = h->rs, wd = rs + 8;
    double *tail = h->d_red + tail_off(h);
    auto colsum = [&](bool wside) -> int {
        vb::ColsumArgs a{wside ? h->n : h->m, wside ? h->d_lw : h->d_lh,
                         wside ? h->d_partW : h->d_partH, wside ? h->d_scal : tail,
                         h->d_counters + (wside ? 0 : 1)};
        h->tab->colsum(a, h->stream);
        h->launches += 1;
        return 0;
    };
    auto mlupd = [&](bool wside) -> int {
        vb::MlUpdateArgs a{wside ? h->n : h->m, r, eps, wside ? tail : h->d_scal,
                           wside ? h->d_red : h->d_ShRaw, wside ? h->d_lw : h->d_lh,
                           wside ? h->d_partW : h->d_partH, wside ? h->d_scal : tail,
                           h->d_counters + (wside ? 0 : 1)};
        h->tab->ml_update(a, h->stream);
        h->launches += 1;
        return 0;
    };
    auto lik_now = [&](double *lik) -> int {
        // likelihood (R/factorize.R:40-49) of the current w, h; needs xlogp of a column sweep at
        // these factors, colSums(w) in d_scal and rowSums(h) in the tail
        if ((rc = fetch_scalars(h))) return rc;
        double swh = 0.0;
        for (int k = 0; k < r; k++) swh += h->h_scal[k] * h->h_scal[wd + k];
        *lik = (h->h_scal[wd + rs + 3] - swh + h->mlconst) / (double)h->n / (double)h->m_global;
        return 0;
    };
    if ((rc = colsum(true))) return rc;   // colSums(w0)
    if ((rc = colsum(false))) return rc;  // rowSums(h0) (local)
    if ((rc = allreduce(h, tail, rs))) return rc;
    double lkold = -INFINITY, lk0 = NAN;
    int it, done = 0;
    for (it = 1; it <= itmax; it++) {                                      // R/factorize.R:191
        if ((rc = launch_sweep_cols(h))) return rc;  // ShRaw and xlogp at (w, h) of iteration it-1
        if ((rc = allreduce(h, tail + rs + 3, 2))) return rc;
        if (it > 1) {
            if ((rc = lik_now(&lk0))) return rc;                           // :193
            if (lik_trace) lik_trace[it - 2] = lk0;
            if (fabs(lkold - lk0) < tol * fabs(lkold)) { done = it - 1; break; }  // :207
            lkold = lk0;
        }
        if ((rc = mlupd(false))) return rc;          // h update, :8-15 -> rowSums(h_new) in tail
        if ((rc = launch_sweep_rows(h))) return rc;  // SwRaw at (w, h_new), :17
        if ((rc = allreduce(h, h->d_red, tail_off(h) + rs))) return rc;
        if ((rc = mlupd(true))) return rc;           // w update, :17-24 -> colSums(w_new) in d_scal
    }
    if (!done) {
        if ((rc = launch_sweep_cols(h))) return rc;
        if ((rc = allreduce(h, tail + rs + 3, 2))) return rc;
        if ((rc = lik_now(&lk0))) return rc;
        if (lik_trace) lik_trace[itmax - 1] = lk0;
        done = itmax;
    }
    CK(cudaGetLastError());
    *niter = done;
    // export w, h
    if (w || h_out) {
        if ((rc = vbnmf_get_state(h, w, h_out, nullptr, nullptr, nullptr, nullptr))) return rc;
    }
    return 0;
}

}  // extern "C"
