/* TEST INFRASTRUCTURE ONLY (oracle/).  CPU restatement, in plain C + OpenMP on a CSC matrix, of
 * ccfindR's variational-Bayes NMF hot path.  It is the checker for the CUDA engine and the timed
 * CPU baseline of bench.py; the product (ccfindr_b200/) never links, imports or calls it.
 *
 * Reference lines restated (paths relative to /root/reference):
 *   src/vbnmf_update.cpp:33-36   statistics  p = lw.lh, q = x/p, sw, sh        -> sweep()
 *   src/vbnmf_update.cpp:38-46   W posterior alw, bew, ew, dw                   -> w_update()
 *   src/vbnmf_update.cpp:48-56   H posterior alh, beh, eh, dh                   -> h_update()
 *   src/vbnmf_update.cpp:58-65   lw, lh = max(exp(psi(alpha))/beta, fudge)      -> w_update()/h_update()
 *   src/vbnmf_update.cpp:67-90   variational lower bound lkh                    -> bound()
 *   R/bayesian.R:2-53            hyper_update (Newton on aw, ah; bw, bh)        -> osp_hyper_update()
 *   R/bayesian.R:336-352         iteration loop of vb_iterate                   -> osp_vb_run()
 *   R/factorize.R:2-27,40-49     nmf_updateR + likelihood (ML path)             -> osp_ml_run()
 *
 * Parity pin: osp_vb_step() is checked in tests/test_oracle.py against (1) the reference's own
 * src/vbnmf_update.cpp compiled in place into oracle/_ref (with stand-in Eigen/Rcpp/GSL headers,
 * see oracle/shim/) and (2) the golden vectors in tests/golden/ produced by that binary.  The
 * R-level driver pieces (hyper_update, the loop, the ML path) exist only as R source and R is not
 * installed here: for those this file is a restatement with PARITY UNPINNED.
 *
 * The dense reference touches all n*m entries; this file touches only the nonzeros.  The two
 * agree because q = x/p is 0 wherever x = 0 (p > 0 is guaranteed by fudge > 0), lgamma(0+1) = 0,
 * sum_ij (ew.eh)_ij = sum_k colsum(ew)_k rowsum(eh)_k, and
 *   sum_ij x_ij (A+B)_ij / p_ij = sum_ik log(lw_ik) lw_ik Sw_ik + sum_kj log(lh_kj) lh_kj Sh_kj
 * with A = (lw o log lw).lh, B = lw.(lh o log lh) (vbnmf_update.cpp:69-75) and Sw, Sh the raw
 * sums of :35-36 taken at the same lw, lh.  One deliberate difference: the reference divides by
 * the int product n*m (:90), which overflows past 2^31-1; here the divisor is (double)n*(double)m.
 *
 * Interface layout = R's: column-major doubles, lw/ew/dw n x r, lh/eh/dh r x m.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t n, m;
    int r;
    const int64_t *colptr;
    const int32_t *rowidx;
    const double *val;
} csc_t;

/* ---- special functions (double) ------------------------------------------------------- */
double osp_digamma(double x) {
    /* upward recurrence to xs >= 10, Stirling series there, then the 1/(x+j) terms are added
     * smallest first so that the dominant -1/x enters last (keeps small-x results to ~1 ulp) */
    const int nstep = x < 10.0 ? (int)ceil(10.0 - x) : 0;
    const double xs = x + (double)nstep;
    const double xi = 1.0 / xs, x2 = xi * xi;
    const double s = x2 * (1.0 / 12.0 - x2 * (1.0 / 120.0 - x2 * (1.0 / 252.0 - x2 * (1.0 / 240.0 -
                     x2 * (1.0 / 132.0 - x2 * (691.0 / 32760.0 - x2 * (1.0 / 12.0 -
                     x2 * (3617.0 / 8160.0))))))));
    double acc = log(xs) - 0.5 * xi - s;
    for (int j = nstep - 1; j >= 0; j--) acc -= 1.0 / (x + (double)j);
    return acc;
}

double osp_trigamma(double x) {
    const int nstep = x < 10.0 ? (int)ceil(10.0 - x) : 0;
    const double xs = x + (double)nstep;
    const double xi = 1.0 / xs, x2 = xi * xi;
    /* 1/x + 1/(2x^2) + sum B_2k / x^(2k+1) */
    const double s = xi * x2 * (1.0 / 6.0 - x2 * (1.0 / 30.0 - x2 * (1.0 / 42.0 - x2 * (1.0 / 30.0 -
                     x2 * (5.0 / 66.0 - x2 * (691.0 / 2730.0 - x2 * (7.0 / 6.0)))))));
    double acc = xi + 0.5 * x2 + s;
    for (int j = nstep - 1; j >= 0; j--) acc += 1.0 / ((x + (double)j) * (x + (double)j));
    return acc;
}

int osp_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- one pass over the nonzeros (vbnmf_update.cpp:33-36 and the data term of :67-81) ---- */
/* lwT: n x r row-major (gene row contiguous); lh: r x m column-major (cell column contiguous).
 * Outputs: SwRaw[i*r+k] = sum_j q_ij lh_kj, ShRaw[j*r+k] = sum_i lw_ik q_ij,
 *          *xlogp = sum x_ij log p_ij.  SwRaw/ShRaw may be NULL (bound-only pass). */
static void sweep(const csc_t *X, const double *lwT, const double *lh, double *SwRaw, double *ShRaw,
                  double *xlogp) {
    const int64_t n = X->n, m = X->m;
    const int r = X->r;
    const int nt = osp_num_threads();
    double *swp = NULL;
    if (SwRaw) swp = (double *)calloc((size_t)nt * (size_t)n * (size_t)r, sizeof(double));
    double xl_tot = 0.0;
#pragma omp parallel reduction(+ : xl_tot)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        double *sw = swp ? swp + (size_t)tid * (size_t)n * (size_t)r : NULL;
        double shacc[256];
#pragma omp for schedule(dynamic, 64)
        for (int64_t j = 0; j < m; j++) {
            const double *lhj = lh + j * r;
            for (int k = 0; k < r; k++) shacc[k] = 0.0;
            for (int64_t t = X->colptr[j]; t < X->colptr[j + 1]; t++) {
                const int64_t i = X->rowidx[t];
                const double x = X->val[t];
                const double *lwi = lwT + i * r;
                double p = 0.0;
                for (int k = 0; k < r; k++) p += lwi[k] * lhj[k];
                const double q = x / p;
                xl_tot += x * log(p);
                if (sw) {
                    double *swi = sw + i * r;
                    for (int k = 0; k < r; k++) {
                        shacc[k] += lwi[k] * q;
                        swi[k] += q * lhj[k];
                    }
                }
            }
            if (ShRaw)
                for (int k = 0; k < r; k++) ShRaw[j * r + k] = shacc[k];
        }
    }
    if (SwRaw) {
        const int64_t nr = n * (int64_t)r;
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < nr; e++) {
            double a = 0.0;
            for (int t = 0; t < nt; t++) a += swp[(size_t)t * (size_t)nr + e];
            SwRaw[e] = a;
        }
        free(swp);
    }
    *xlogp = xl_tot;
}

/* vbnmf_update.cpp:38-46, 58-61, 82-86: W posterior from the raw statistics.
 * ehsum[k] = rowSums(eh_old) (:42-43).  Writes alw, new lwT, ewsum[k] = colSums(ew_new);
 * returns in acc[0..3]: W prior/entropy part of the bound, sum log lw_new, sum ew_new, unused. */
static void w_update(int64_t n, int r, const double *SwRaw, double *lwT, double *alwT,
                     const double *ehsum, double aw, double bw, double fud, double *ewsum,
                     double *acc) {
    double bew[256], lbew[256];
    for (int k = 0; k < r; k++) {
        bew[k] = aw / bw + ehsum[k];
        lbew[k] = log(bew[k]);
        ewsum[k] = 0.0;
    }
    const double lga = -lgamma(aw) + aw * log(aw / bw);
    double prior = 0.0, sll = 0.0, sew = 0.0;
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < r; k++) {
            const int64_t e = i * r + k;
            const double alw = aw + lwT[e] * SwRaw[e];
            const double ew = alw / bew[k];
            const double tmp = exp(osp_digamma(alw)) / bew[k];
            const double lwn = tmp > fud ? tmp : fud;
            alwT[e] = alw;
            lwT[e] = lwn;
            ewsum[k] += ew;
            sew += ew;
            sll += log(lwn);
            prior += -(aw / bw) * ew + lga + alw * (1.0 - lbew[k]) + lgamma(alw);
        }
    acc[0] = prior;
    acc[1] = sll;
    acc[2] = sew;
    acc[3] = 0.0;
}

/* vbnmf_update.cpp:48-56, 62-65, 87-89: H posterior. ewsum[k] = colSums(ew_new) (:52-53). */
static void h_update(int64_t m, int r, const double *ShRaw, double *lh, double *alh,
                     const double *ewsum, double ah, double bh, double fud, double *ehsum,
                     double *acc) {
    double beh[256], lbeh[256];
    for (int k = 0; k < r; k++) {
        beh[k] = ah / bh + ewsum[k];
        lbeh[k] = log(beh[k]);
    }
    const double lga = -lgamma(ah) + ah * log(ah / bh);
    const int nt = osp_num_threads();
    double *part = (double *)calloc((size_t)nt * (size_t)(r + 4), sizeof(double));
#pragma omp parallel
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        double *pp = part + (size_t)tid * (size_t)(r + 4);
#pragma omp for schedule(static)
        for (int64_t j = 0; j < m; j++)
            for (int k = 0; k < r; k++) {
                const int64_t e = j * r + k;
                const double a = ah + lh[e] * ShRaw[e];
                const double eh = a / beh[k];
                const double tmp = exp(osp_digamma(a)) / beh[k];
                const double lhn = tmp > fud ? tmp : fud;
                alh[e] = a;
                lh[e] = lhn;
                pp[k] += eh;
                pp[r + 0] += -(ah / bh) * eh + lga + a * (1.0 - lbeh[k]) + lgamma(a);
                pp[r + 1] += log(lhn);
                pp[r + 2] += eh;
            }
    }
    for (int k = 0; k < r; k++) ehsum[k] = 0.0;
    acc[0] = acc[1] = acc[2] = acc[3] = 0.0;
    for (int t = 0; t < nt; t++) {
        const double *pp = part + (size_t)t * (size_t)(r + 4);
        for (int k = 0; k < r; k++) ehsum[k] += pp[k];
        acc[0] += pp[r + 0];
        acc[1] += pp[r + 1];
        acc[2] += pp[r + 2];
    }
    free(part);
}

/* entropy-collapse terms of the bound at the (new) lw, lh with their own raw statistics */
static double ent_w(int64_t n, int r, const double *lwT, const double *SwRaw) {
    double s = 0.0;
    for (int64_t e = 0; e < n * (int64_t)r; e++) s += log(lwT[e]) * lwT[e] * SwRaw[e];
    return s;
}
static double ent_h(int64_t m, int r, const double *lh, const double *ShRaw) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t e = 0; e < m * (int64_t)r; e++) s += log(lh[e]) * lh[e] * ShRaw[e];
    return s;
}

/* sum over nonzeros of lgamma(x+1) (vbnmf_update.cpp:80-81; zero entries contribute 0) */
double osp_lgx_sum(int64_t nnz, const double *val) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t t = 0; t < nnz; t++) s += lgamma(val[t] + 1.0);
    return s;
}

static void transpose_in(int64_t n, int r, const double *cm, double *rm) { /* n x r col-major -> row-major */
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < r; k++) rm[i * r + k] = cm[(int64_t)k * n + i];
}
static void transpose_out(int64_t n, int r, const double *rm, double *cm) {
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < r; k++) cm[(int64_t)k * n + i] = rm[i * r + k];
}

/* workspace of one factorization */
typedef struct {
    csc_t X;
    double *lwT, *alwT, *SwRaw, *lh, *alh, *ShRaw;
    double ewsum[256], ehsum[256], bew[256], beh[256];
    double wacc[4], hacc[4];
    double lgx;
} work_t;

static int work_alloc(work_t *w, int64_t n, int64_t m, int r, const int64_t *colptr,
                      const int32_t *rowidx, const double *val) {
    if (r > 256 || r < 1) return 1;
    memset(w, 0, sizeof(*w));
    w->X.n = n; w->X.m = m; w->X.r = r;
    w->X.colptr = colptr; w->X.rowidx = rowidx; w->X.val = val;
    const size_t nr = (size_t)n * r, rm = (size_t)m * r;
    w->lwT = (double *)malloc(nr * 8); w->alwT = (double *)malloc(nr * 8);
    w->SwRaw = (double *)malloc(nr * 8);
    w->lh = (double *)malloc(rm * 8); w->alh = (double *)malloc(rm * 8);
    w->ShRaw = (double *)malloc(rm * 8);
    w->lgx = osp_lgx_sum(colptr[m], val);
    return 0;
}
static void work_free(work_t *w) {
    free(w->lwT); free(w->alwT); free(w->SwRaw); free(w->lh); free(w->alh); free(w->ShRaw);
}

/* posterior update from the statistics already in SwRaw/ShRaw, then a sweep at the new lw, lh that
 * yields both the bound of THIS iteration and the statistics of the NEXT one. */
static double iterate(work_t *w, const double *hyper, double fud) {
    const int64_t n = w->X.n, m = w->X.m;
    const int r = w->X.r;
    const double aw = hyper[0], bw = hyper[1], ah = hyper[2], bh = hyper[3];
    for (int k = 0; k < r; k++) w->bew[k] = aw / bw + w->ehsum[k];
    w_update(n, r, w->SwRaw, w->lwT, w->alwT, w->ehsum, aw, bw, fud, w->ewsum, w->wacc);
    for (int k = 0; k < r; k++) w->beh[k] = ah / bh + w->ewsum[k];
    h_update(m, r, w->ShRaw, w->lh, w->alh, w->ewsum, ah, bh, fud, w->ehsum, w->hacc);
    double xlogp;
    sweep(&w->X, w->lwT, w->lh, w->SwRaw, w->ShRaw, &xlogp);
    double U = 0.0;
    for (int k = 0; k < r; k++) U -= w->ewsum[k] * w->ehsum[k];
    U -= ent_w(n, r, w->lwT, w->SwRaw) + ent_h(m, r, w->lh, w->ShRaw) - xlogp;
    U -= w->lgx;
    U += w->wacc[0] + w->hacc[0];
    return U / ((double)n * (double)m);
}

static void export_state(const work_t *w, double *lw, double *lh, double *ew, double *eh,
                         double *dw, double *dh) {
    const int64_t n = w->X.n, m = w->X.m;
    const int r = w->X.r;
    if (lw) transpose_out(n, r, w->lwT, lw);
    if (lh) memcpy(lh, w->lh, sizeof(double) * (size_t)m * r);
    for (int64_t i = 0; i < n; i++)
        for (int k = 0; k < r; k++) {
            const double a = w->alwT[i * r + k], b = w->bew[k];
            if (ew) ew[(int64_t)k * n + i] = a / b;
            if (dw) dw[(int64_t)k * n + i] = a / b / b;
        }
    for (int64_t j = 0; j < m; j++)
        for (int k = 0; k < r; k++) {
            const double a = w->alh[j * r + k], b = w->beh[k];
            if (eh) eh[j * r + k] = a / b;
            if (dh) dh[j * r + k] = a / b / b;
        }
}

/* One call of the reference's vbnmf_update() (src/vbnmf_update.cpp:16-102) on CSC input.
 * hyper = {aw, bw, ah, bh}.  lw, lh: in/out.  eh_in: the eh of the incoming wh (only its row sums
 * are read, :42-43).  means4 (optional) = mean(log lw), mean(log lh), mean(ew), mean(eh) of the
 * result, the inputs of hyper_update (R/bayesian.R:8-11). */
int osp_vb_step(int64_t n, int64_t m, int r, const int64_t *colptr, const int32_t *rowidx,
                const double *val, double *lw, double *lh, const double *eh_in, double *ew,
                double *eh, double *dw, double *dh, const double *hyper, double fudge,
                double *lkh, double *means4) {
    work_t w;
    if (work_alloc(&w, n, m, r, colptr, rowidx, val)) return 1;
    transpose_in(n, r, lw, w.lwT);
    memcpy(w.lh, lh, sizeof(double) * (size_t)m * r);
    for (int k = 0; k < r; k++) {
        double s = 0.0;
        for (int64_t j = 0; j < m; j++) s += eh_in[j * r + k];
        w.ehsum[k] = s;
    }
    double xl;
    sweep(&w.X, w.lwT, w.lh, w.SwRaw, w.ShRaw, &xl);
    *lkh = iterate(&w, hyper, fudge);
    export_state(&w, lw, lh, ew, eh, dw, dh);
    if (means4) {
        means4[0] = w.wacc[1] / ((double)n * r);
        means4[1] = w.hacc[1] / ((double)m * r);
        means4[2] = w.wacc[2] / ((double)n * r);
        means4[3] = w.hacc[2] / ((double)m * r);
    }
    work_free(&w);
    return 0;
}

/* R/bayesian.R:2-53.  flags = hyper.update (aw, bw, ah, bh); hyper = {aw, bw, ah, bh} in/out.
 * Returns 0, or 2 when the Newton loop exhausts niter ('Hyper-parameter update failed to
 * converge', :43). */
int osp_hyper_update(const int *flags, double lwm, double lhm, double ewm, double ehm,
                     double *hyper, int niter, double tol) {
    if (flags[0] + flags[1] + flags[2] + flags[3] == 0) return 0;
    double aw0 = hyper[0], ah0 = hyper[2];
    const double bw0 = hyper[1], bh0 = hyper[3];
    double aw1, ah1;
    if (flags[0] + flags[2] > 0) {
        int i = 1;
        aw1 = aw0; ah1 = ah0;
        while (i < niter) {
            double dw = 0.0, dh = 0.0;
            if (flags[0])
                dw = (log(aw0) - osp_digamma(aw0) - ewm / bw0 + 1.0 + lwm - log(bw0)) /
                     (1.0 / aw0 - osp_trigamma(aw0));
            if (flags[2])
                dh = (log(ah0) - osp_digamma(ah0) - ehm / bh0 + 1.0 + lhm - log(bh0)) /
                     (1.0 / ah0 - osp_trigamma(ah0));
            aw1 = aw0 - dw;
            ah1 = ah0 - dh;
            while (aw1 <= 0) { dw = dw / 2; aw1 = aw0 - dw; }
            while (ah1 <= 0) { dh = dh / 2; ah1 = ah0 - dh; }
            const double df = (1 - aw1 / aw0) * (1 - aw1 / aw0) + (1 - ah1 / ah0) * (1 - ah1 / ah0);
            if (df < tol) break;
            aw0 = aw1;
            ah0 = ah1;
            i++;
        }
        if (i == niter) return 2;
    } else {
        aw1 = aw0;
        ah1 = ah0;
    }
    hyper[0] = aw1;
    hyper[1] = flags[1] ? ewm : bw0;
    hyper[2] = ah1;
    hyper[3] = ehm; /* both branches of R/bayesian.R:50-51 assign ehm */
    return 0;
}

/* The iteration loop of vb_iterate for one rank (R/bayesian.R:336-352).
 * cfg_i = {itmax, n0, dn, hu_aw, hu_bw, hu_ah, hu_bh};  cfg_d = {tol, fudge}.
 * lw/lh in: initial factors (vb_init sets lw=ew=w, lh=eh=h, :170); out: final state.
 * lkh_trace[it-1] = lkh of iteration it; hyper_trace[4*(it-1)..] = hyper AFTER iteration it.
 * lml = lk0 as stored at R/bayesian.R:379.  stop_reason: 0 Itmax, 1 converged, 2 NaN.
 * Returns 0, or 2 on hyper-update failure. */
int osp_vb_run(int64_t n, int64_t m, int r, const int64_t *colptr, const int32_t *rowidx,
               const double *val, double *lw, double *lh, double *ew, double *eh, double *dw,
               double *dh, const int *cfg_i, const double *cfg_d, double *hyper,
               double *lkh_trace, double *hyper_trace, int *niter, double *lml, int *stop_reason) {
    const int itmax = cfg_i[0], n0 = cfg_i[1], dn = cfg_i[2];
    const int *flags = cfg_i + 3;
    const double tol = cfg_d[0], fud = cfg_d[1];
    work_t w;
    if (work_alloc(&w, n, m, r, colptr, rowidx, val)) return 1;
    transpose_in(n, r, lw, w.lwT);
    memcpy(w.lh, lh, sizeof(double) * (size_t)m * r);
    for (int k = 0; k < r; k++) {
        double s = 0.0;
        for (int64_t j = 0; j < m; j++) s += lh[j * r + k]; /* eh = h at init */
        w.ehsum[k] = s;
    }
    double xl, lk0 = 0.0;
    int it, rc = 0, reason = 0;
    sweep(&w.X, w.lwT, w.lh, w.SwRaw, w.ShRaw, &xl);
    for (it = 1; it <= itmax; it++) {
        const double lkh = iterate(&w, hyper, fud);
        /* bew/beh of THIS iteration are what ew/dw/eh/dh derive from; iterate() stored them
         * from the hypers in force before the hyper update below */
        if (it > n0 && it % dn == 0) {
            const double means[4] = {w.wacc[1] / ((double)n * r), w.hacc[1] / ((double)m * r),
                                     w.wacc[2] / ((double)n * r), w.hacc[2] / ((double)m * r)};
            rc = osp_hyper_update(flags, means[0], means[1], means[2], means[3], hyper, 100, 1e-3);
            if (rc) break;
        }
        if (lkh_trace) lkh_trace[it - 1] = lkh;
        if (hyper_trace) memcpy(hyper_trace + 4 * (it - 1), hyper, 4 * sizeof(double));
        if (isnan(lkh)) { reason = 2; break; }
        if (it > 1 && it > n0 && lkh >= lk0 && fabs(1 - lkh / lk0) < tol) { reason = 1; break; }
        lk0 = lkh;
    }
    if (it > itmax) it = itmax; /* R leaves `it` at Itmax when the for loop runs out */
    export_state(&w, lw, lh, ew, eh, dw, dh);
    *niter = it;
    *lml = lk0;
    *stop_reason = reason;
    work_free(&w);
    return rc;
}

/* ---- maximum-likelihood path (R/factorize.R:2-27 nmf_updateR, :40-49 likelihood, loop :189-212)
 * w n x r col-major, h r x m col-major, in/out.  lik_trace[it-1] = likelihood after iteration it.
 * Stops when |lkold - lk0| < tol*|lkold| (:207).  eps = .Machine$double.eps clamp (:15,:24). */
static double ml_lik_const(int64_t nnz, const double *val) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (int64_t t = 0; t < nnz; t++)
        if (val[t] > 0) s += -val[t] * log(val[t]) + val[t];
    return s;
}

int osp_ml_run(int64_t n, int64_t m, int r, const int64_t *colptr, const int32_t *rowidx,
               const double *val, double *w_io, double *h_io, int itmax, double tol,
               double *lik_trace, int *niter) {
    if (r > 256) return 1;
    const double eps = 2.220446049250313e-16;
    csc_t X = {n, m, r, colptr, rowidx, val};
    const size_t nr = (size_t)n * r, rm = (size_t)m * r;
    double *wT = (double *)malloc(nr * 8), *Sw = (double *)malloc(nr * 8);
    double *h = (double *)malloc(rm * 8), *Sh = (double *)malloc(rm * 8);
    transpose_in(n, r, w_io, wT);
    memcpy(h, h_io, rm * 8);
    const double lconst = ml_lik_const(colptr[m], val);
    double lkold = -INFINITY, xl;
    int it;
    for (it = 1; it <= itmax; it++) {
        double csum[256], rsum[256];
        /* H update (:8-15): h <- h o (w^T (x/(w h))) / colSums(w), clamp */
        sweep(&X, wT, h, Sw, Sh, &xl);
        for (int k = 0; k < r; k++) { csum[k] = 0.0; rsum[k] = 0.0; }
        for (int64_t i = 0; i < n; i++)
            for (int k = 0; k < r; k++) csum[k] += wT[i * r + k];
        for (int64_t j = 0; j < m; j++)
            for (int k = 0; k < r; k++) {
                double v = h[j * r + k] * Sh[j * r + k] / csum[k];
                if (v < eps) v = eps;
                h[j * r + k] = v;
            }
        /* W update with the new h (:17-24) */
        sweep(&X, wT, h, Sw, Sh, &xl);
        for (int64_t j = 0; j < m; j++)
            for (int k = 0; k < r; k++) rsum[k] += h[j * r + k];
        for (int64_t i = 0; i < n; i++)
            for (int k = 0; k < r; k++) {
                double v = wT[i * r + k] * Sw[i * r + k] / rsum[k];
                if (v < eps) v = eps;
                wT[i * r + k] = v;
            }
        /* likelihood (:40-49) at the new w, h */
        sweep(&X, wT, h, NULL, NULL, &xl);
        for (int k = 0; k < r; k++) csum[k] = 0.0;
        for (int64_t i = 0; i < n; i++)
            for (int k = 0; k < r; k++) csum[k] += wT[i * r + k];
        double swh = 0.0;
        for (int k = 0; k < r; k++) swh += csum[k] * rsum[k];
        const double lk0 = (xl - swh + lconst) / (double)n / (double)m;
        if (lik_trace) lik_trace[it - 1] = lk0;
        if (fabs(lkold - lk0) < tol * fabs(lkold)) break;
        lkold = lk0;
    }
    if (it > itmax) it = itmax;
    transpose_out(n, r, wT, w_io);
    memcpy(h_io, h, rm * 8);
    *niter = it;
    free(wT); free(Sw); free(h); free(Sh);
    return 0;
}

/* timing helper for bench.py's cpu_baseline: `iters` steady-state iterations (posterior update +
 * one sweep each) on the given matrix, hypers fixed.  Returns seconds per iteration. */
double osp_time_iterations(int64_t n, int64_t m, int r, const int64_t *colptr,
                           const int32_t *rowidx, const double *val, const double *lw,
                           const double *lh, const double *hyper, double fudge, int iters,
                           double *lkh_last) {
    work_t w;
    if (work_alloc(&w, n, m, r, colptr, rowidx, val)) return -1.0;
    transpose_in(n, r, lw, w.lwT);
    memcpy(w.lh, lh, sizeof(double) * (size_t)m * r);
    for (int k = 0; k < r; k++) {
        double s = 0.0;
        for (int64_t j = 0; j < m; j++) s += lh[j * r + k];
        w.ehsum[k] = s;
    }
    double xl, lkh = 0.0;
    sweep(&w.X, w.lwT, w.lh, w.SwRaw, w.ShRaw, &xl);
#ifdef _OPENMP
    const double t0 = omp_get_wtime();
#else
    const double t0 = 0.0;
#endif
    for (int it = 0; it < iters; it++) lkh = iterate(&w, hyper, fudge);
#ifdef _OPENMP
    const double t1 = omp_get_wtime();
#else
    const double t1 = 0.0;
#endif
    if (lkh_last) *lkh_last = lkh;
    work_free(&w);
    return (t1 - t0) / (double)iters;
}
