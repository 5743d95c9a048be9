// TEST INFRASTRUCTURE ONLY (oracle/): C entry point around the reference's own
// vbnmf_update() (/root/reference/src/vbnmf_update.cpp:16-102), which oracle/Makefile compiles
// in place against the stand-in headers in oracle/shim/.  Nothing here restates the
// algorithm; this file only marshals plain arrays into the Rcpp::List arguments the
// reference function takes (the job src/RcppExports.cpp:11-22 does inside R) and back.
#include <RcppEigen.h>
#include <cstring>

Rcpp::List vbnmf_update(const Eigen::MatrixXd &X, const Rcpp::List &wh, const Rcpp::List &hyper,
                        const Rcpp::NumericVector &fudge);

static Eigen::MatrixXd wrap(const double *p, int r, int c) {
    Eigen::MatrixXd m(r, c);
    std::memcpy(m.data(), p, sizeof(double) * (size_t)r * (size_t)c);
    return m;
}
static void unwrap(const Rcpp::List &z, const char *key, double *out) {
    if (!out) return;
    Eigen::MatrixXd m = z[key];
    std::memcpy(out, m.data(), sizeof(double) * (size_t)m.rows() * (size_t)m.cols());
}

extern "C" {

// All matrices column-major doubles, exactly as R hands them to .Call:
// X n x m, lw/ew n x r, lh/eh r x m.  hyper = {aw, bw, ah, bh}.  Outputs may be NULL.
int ref_vbnmf_update(int n, int m, int r, const double *X, const double *lw, const double *lh,
                     const double *ew, const double *eh, const double *hyper, double fudge,
                     double *o_lw, double *o_lh, double *o_ew, double *o_eh, double *o_dw,
                     double *o_dh, double *o_lkh) {
    Rcpp::List wh, hy;
    wh.set("lw", wrap(lw, n, r));
    wh.set("lh", wrap(lh, r, m));
    wh.set("ew", wrap(ew, n, r));
    wh.set("eh", wrap(eh, r, m));
    hy.set("aw", hyper[0]);
    hy.set("bw", hyper[1]);
    hy.set("ah", hyper[2]);
    hy.set("bh", hyper[3]);
    Rcpp::NumericVector fv{fudge};
    Rcpp::List z = vbnmf_update(wrap(X, n, m), wh, hy, fv);
    unwrap(z, "lw", o_lw);
    unwrap(z, "lh", o_lh);
    unwrap(z, "ew", o_ew);
    unwrap(z, "eh", o_eh);
    unwrap(z, "dw", o_dw);
    unwrap(z, "dh", o_dh);
    if (o_lkh) *o_lkh = (double)z["lkh"];
    return 0;
}

const char *ref_source_path(void) { return REF_SOURCE_PATH; }
}
