// One translation unit per padded rank: nvcc -DVB_RP=<RP> rp_inst.cu
#include "kernels.cuh"
#include "rp_table.h"

#ifndef VB_RP
#error "compile with -DVB_RP=<padded rank>"
#endif
#define VB_CAT2(a, b) a##b
#define VB_CAT(a, b) VB_CAT2(a, b)

namespace vb {
namespace {

constexpr int RP = VB_RP;

void sweep_cols(const SweepColsArgs &a, bool vf, int grid, cudaStream_t s) {
    if (vf)
        sweep_cols_kernel<RP, float><<<grid, kBlock, 0, s>>>(
            a.m, a.r, a.colptr, a.rowidx, (const float *)a.val, a.lw, a.lh, a.ShRaw, a.col_xlogp,
            a.col_enth, a.work_counter);
    else
        sweep_cols_kernel<RP, double><<<grid, kBlock, 0, s>>>(
            a.m, a.r, a.colptr, a.rowidx, (const double *)a.val, a.lw, a.lh, a.ShRaw, a.col_xlogp,
            a.col_enth, a.work_counter);
}

void sweep_rows(const SweepRowsArgs &a, bool vf, int grid, cudaStream_t s) {
    if (vf)
        sweep_rows_kernel<RP, float><<<grid, kBlock, 0, s>>>(
            a.n_items, a.item_row, a.item_beg, a.item_len, a.colidx, (const float *)a.val, a.lw,
            a.lh, a.SwPart, a.work_counter);
    else
        sweep_rows_kernel<RP, double><<<grid, kBlock, 0, s>>>(
            a.n_items, a.item_row, a.item_beg, a.item_len, a.colidx, (const double *)a.val, a.lw,
            a.lh, a.SwPart, a.work_counter);
}

template <typename K>
int ctas_per_sm(K k) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, kBlock, 0) != cudaSuccess || nb < 1)
        nb = 1;
    return nb;
}
int occ_cols(bool vf) {
    return vf ? ctas_per_sm(sweep_cols_kernel<RP, float>) : ctas_per_sm(sweep_cols_kernel<RP, double>);
}
int occ_rows(bool vf) {
    return vf ? ctas_per_sm(sweep_rows_kernel<RP, float>) : ctas_per_sm(sweep_rows_kernel<RP, double>);
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

void posterior(const PosteriorArgs &a, cudaStream_t s) {
    posterior_kernel<RP><<<cdiv(a.rows, kBlock), kBlock, 0, s>>>(
        a.rows, a.r, a.a, a.b, a.fud, a.osum, a.SRaw, a.l, a.al_out, a.part, a.out, a.counter);
}
void ml_update(const MlUpdateArgs &a, cudaStream_t s) {
    ml_update_kernel<RP><<<cdiv(a.rows, kBlock), kBlock, 0, s>>>(a.rows, a.r, a.eps, a.osum, a.SRaw,
                                                                 a.v, a.part, a.out, a.counter);
}
void colsum(const ColsumArgs &a, cudaStream_t s) {
    panel_colsum_kernel<RP><<<cdiv(a.rows, kBlock), kBlock, 0, s>>>(a.rows, a.v, a.part, a.out,
                                                                   a.counter);
}

}  // namespace

extern const RpTable VB_CAT(rp_table_, VB_RP);
const RpTable VB_CAT(rp_table_, VB_RP) = {RP,       sweep_cols, sweep_rows, occ_cols,
                                          occ_rows, posterior,  ml_update,  colsum};

}  // namespace vb
