// Internal (not exported) launch table: one entry per padded rank RP, each compiled in its own
// translation unit (rp_inst.cu with -DVB_RP=<RP>) so that the build parallelises.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace vb {

struct CombineArgs {
    int64_t NO;
    int nslabs, r;
    const double *Part, *l;
    double *SRaw, *part, *out;
    unsigned *counter;
    const double *xl_part;  // optional per-CTA partial sums folded into out[1]
    int nxl;
    int grid;
    const double *ctl;      // device loop control block or nullptr
    int tsplit;             // layout of l (panel_ofs)
};

struct PosteriorArgs {
    int64_t rows;
    int T, S;
    int64_t nvalid;
    int r;
    double a, b, fud;
    const double *osum, *SRaw;
    double *l, *al_out, *part, *out;
    unsigned *counter;
    float *l32;  // fp32 mirror of l (fp32-storage mode) or nullptr
    const double *ctl;  // device loop control block (hypers read from it) or nullptr
    int hoff;           // 0: (aw, bw), 2: (ah, bh)
    int rows_per_cta;   // post_rows_per_cta(rows, rs, SMs)
    int tsplit;         // layout of l (panel_ofs)
};

struct MlUpdateArgs {
    int64_t rows;
    int T, S;
    int64_t nvalid;
    int r;
    double eps;
    const double *osum, *SRaw;
    double *v, *part, *out;
    unsigned *counter;
    float *l32;
    int tsplit;  // layout of v (panel_ofs)
    const double *ctl;  // device loop control block or nullptr
};

struct ColsumArgs {
    int64_t rows;
    const double *v;
    double *part, *out;
    unsigned *counter;
    int tsplit;  // layout of v (panel_ofs)
};

struct RpTable {
    int rp, rs, rsf;  // padded rank, fp64 panel stride (doubles), fp32 mirror stride (floats)
    int npg64, npg32;  // nonzeros per 8-lane group step of the sweep (fp64 / fp32 panels)
    int split64;       // units of block A of the split layout the fp64 packed-16 sweep reads lw/lh in
                       // (panel_ofs): 8 (ranks 16..20), 4 (ranks 8..14) or 0 (plain rows)
    int stage64, stage32;  // bytes behind the tile the packed-16 sweeps use as owner staging slots
    int stage64_g4;        // ... of the fp64 split kernels with 4-lane groups; -1: no such kernels
    // cols: cell-owner pass; fmt = storage format of the nonzeros (kEnt*); grid = CTAs (one per
    // SM, persistent)
    int (*sweep_prepare)(int smem_bytes);  // opt in to the dynamic shared memory size; 0 = ok
    // split: 0 plain rows, 1 split layout, 2 split layout with 4-lane groups
    void (*sweep)(const SweepTiledArgs &, bool cols, int fmt, bool panels_f32, int split, int grid,
                  int smem_bytes, cudaStream_t);
    void (*mirror)(int64_t rows, const double *v, float *v32, cudaStream_t);
    void (*combine)(const CombineArgs &, cudaStream_t);
    void (*posterior)(const PosteriorArgs &, cudaStream_t);
    void (*ml_update)(const MlUpdateArgs &, cudaStream_t);
    void (*colsum)(const ColsumArgs &, cudaStream_t);
};

const RpTable *rp_table(int rp);  // nullptr when rp is not instantiated

}  // namespace vb
