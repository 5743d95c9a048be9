// Internal (not exported) launch table: one entry per padded rank RP, each compiled in its own
// translation unit (rp_inst.cu with -DVB_RP=<RP>) so that the build parallelises.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vb {

struct SweepColsArgs {
    int64_t m;
    int r;
    const int64_t *colptr;
    const int32_t *rowidx;
    const void *val;
    const double *lw, *lh;
    double *ShRaw, *col_xlogp, *col_enth;
    unsigned long long *work_counter;
};

struct SweepRowsArgs {
    int64_t n_items;
    const int32_t *item_row;
    const int64_t *item_beg;
    const int32_t *item_len;
    const int32_t *colidx;
    const void *val;
    const double *lw, *lh;
    double *SwPart;
    unsigned long long *work_counter;
};

struct PosteriorArgs {
    int64_t rows;
    int r;
    double a, b, fud;
    const double *osum, *SRaw;
    double *l, *al_out, *part, *out;
    unsigned *counter;
};

struct MlUpdateArgs {
    int64_t rows;
    int r;
    double eps;
    const double *osum, *SRaw;
    double *v, *part, *out;
    unsigned *counter;
};

struct ColsumArgs {
    int64_t rows;
    const double *v;
    double *part, *out;
    unsigned *counter;
};

struct RpTable {
    int rp;
    // val_is_float selects the count storage type; grid = CTAs to launch (persistent kernels)
    void (*sweep_cols)(const SweepColsArgs &, bool val_is_float, int grid, cudaStream_t);
    void (*sweep_rows)(const SweepRowsArgs &, bool val_is_float, int grid, cudaStream_t);
    int (*sweep_cols_ctas_per_sm)(bool val_is_float);
    int (*sweep_rows_ctas_per_sm)(bool val_is_float);
    void (*posterior)(const PosteriorArgs &, cudaStream_t);
    void (*ml_update)(const MlUpdateArgs &, cudaStream_t);
    void (*colsum)(const ColsumArgs &, cudaStream_t);
};

const RpTable *rp_table(int rp);  // nullptr when rp is not instantiated

}  // namespace vb
