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
constexpr int RS = row_stride(RP);

int sweep_prepare(int smem_bytes) {
    cudaError_t e = cudaSuccess;
#define VB_OPT(K)                                                                               \
    if (e == cudaSuccess)                                                                       \
        e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    VB_OPT((sweep_tiled_kernel<RP, float, true, double>))
    VB_OPT((sweep_tiled_kernel<RP, float, false, double>))
    VB_OPT((sweep_tiled_kernel<RP, double, true, double>))
    VB_OPT((sweep_tiled_kernel<RP, double, false, double>))
    VB_OPT((sweep_tiled_kernel<RP, float, true, float>))
    VB_OPT((sweep_tiled_kernel<RP, float, false, float>))
    VB_OPT((sweep_tiled_kernel<RP, double, true, float>))
    VB_OPT((sweep_tiled_kernel<RP, double, false, float>))
    VB_OPT((sweep_p16_kernel<RP, true, double>))
    VB_OPT((sweep_p16_kernel<RP, false, double>))
    VB_OPT((sweep_p16_kernel<RP, true, float>))
    VB_OPT((sweep_p16_kernel<RP, false, float>))
    if constexpr (split_rank(RP)) {
        VB_OPT((sweep_p16_kernel<RP, true, double, true>))
        VB_OPT((sweep_p16_kernel<RP, false, double, true>))
    }
    if constexpr (SweepCfg<RP, double>::g4()) {
        VB_OPT((sweep_p16_kernel<RP, true, double, true, 4>))
        VB_OPT((sweep_p16_kernel<RP, false, double, true, 4>))
    }
#undef VB_OPT
    return e == cudaSuccess ? 0 : 1;
}

// split: 0 plain rows, 1 split layout (8-lane groups), 2 split layout with 4-lane groups
template <typename PT>
void sweep_pt(const SweepTiledArgs &a, bool cols, int fmt, int split, int grid, int smem,
              cudaStream_t s) {
    constexpr int NT = SweepCfg<RP, PT>::kThreads;
    const bool vf = fmt == kEntF32;
    constexpr int NTC = SweepCfg<RP, PT>::p16_threads(true), NTR = SweepCfg<RP, PT>::p16_threads(false);
    if constexpr (SweepCfg<RP, PT>::g4()) {
        if (fmt == kEntP16 && split == 2) {
            if (cols) sweep_p16_kernel<RP, true, PT, true, 4><<<grid, NTC, smem, s>>>(a);
            else sweep_p16_kernel<RP, false, PT, true, 4><<<grid, NTR, smem, s>>>(a);
            return;
        }
    }
    if constexpr (split_rank(RP) && sizeof(PT) == 8) {
        if (fmt == kEntP16 && split) {
            if (cols) sweep_p16_kernel<RP, true, PT, true><<<grid, NTC, smem, s>>>(a);
            else sweep_p16_kernel<RP, false, PT, true><<<grid, NTR, smem, s>>>(a);
            return;
        }
    }
    if (fmt == kEntP16) {
        if (cols) sweep_p16_kernel<RP, true, PT><<<grid, NTC, smem, s>>>(a);
        else sweep_p16_kernel<RP, false, PT><<<grid, NTR, smem, s>>>(a);
    } else if (cols) {
        if (vf) sweep_tiled_kernel<RP, float, true, PT><<<grid, NT, smem, s>>>(a);
        else sweep_tiled_kernel<RP, double, true, PT><<<grid, NT, smem, s>>>(a);
    } else {
        if (vf) sweep_tiled_kernel<RP, float, false, PT><<<grid, NT, smem, s>>>(a);
        else sweep_tiled_kernel<RP, double, false, PT><<<grid, NT, smem, s>>>(a);
    }
}

void sweep(const SweepTiledArgs &a, bool cols, int fmt, bool pf32, int split, int grid, int smem,
           cudaStream_t s) {
    if (pf32) sweep_pt<float>(a, cols, fmt, 0, grid, smem, s);
    else sweep_pt<double>(a, cols, fmt, split, grid, smem, s);
}

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

void combine(const CombineArgs &a, cudaStream_t s) {
    combine_kernel<RP><<<a.grid, kBlock, 0, s>>>(a.NO, a.nslabs, a.r, a.Part, a.l, a.SRaw, a.part,
                                                 a.out, a.counter, a.xl_part, a.nxl, a.ctl,
                                                 a.tsplit);
}
void posterior(const PosteriorArgs &a, cudaStream_t s) {
    posterior_kernel<RP><<<cdiv(a.rows, a.rows_per_cta), post_threads(RS), 0, s>>>(
        a.rows, a.T, a.S, a.nvalid, a.r, a.a, a.b, a.fud, a.osum, a.SRaw, a.l, a.al_out, a.part,
        a.out, a.counter, a.l32, a.ctl, a.hoff, a.rows_per_cta, a.tsplit);
}
void ml_update(const MlUpdateArgs &a, cudaStream_t s) {
    ml_update_kernel<RP><<<cdiv(a.rows, kPostRows), post_threads(RS), 0, s>>>(
        a.rows, a.T, a.S, a.nvalid, a.r, a.eps, a.osum, a.SRaw, a.v, a.part, a.out, a.counter,
        a.l32, a.tsplit, a.ctl);
}
void mirror(int64_t rows, const double *v, float *v32, cudaStream_t s) {
    mirror_kernel<RP><<<cdiv(rows, kBlock), kBlock, 0, s>>>(rows, v, v32);
}
void colsum(const ColsumArgs &a, cudaStream_t s) {
    panel_colsum_kernel<RP><<<cdiv(a.rows, kBlock), kBlock, 0, s>>>(a.rows, a.v, a.part, a.out,
                                                                   a.counter, a.tsplit);
}

}  // namespace

extern const RpTable VB_CAT(rp_table_, VB_RP);
const RpTable VB_CAT(rp_table_, VB_RP) = {
    RP,      RS,        row_stride_f32(RP),
    SweepCfg<RP, double>::kNPG, SweepCfg<RP, float>::kNPG, split_units(RP),
    SweepCfg<RP, double>::stage_total(), SweepCfg<RP, float>::stage_total(),
    SweepCfg<RP, double>::g4() ? SweepCfg<RP, double>::stage_total(4) : -1, sweep_prepare,
    sweep, mirror,
    combine, posterior, ml_update,          colsum};

}  // namespace vb
