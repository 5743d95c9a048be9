"""TEST INFRASTRUCTURE ONLY (oracle/).  Literal NumPy/SciPy restatement, on DENSE matrices, of
ccfindR's variational-Bayes NMF path and its R driver.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this; the product (ccfindr_b200/) never does.

Reference lines restated (paths relative to /root/reference):
  vbnmf_update()     src/vbnmf_update.cpp:16-102  (the native step, `useC=TRUE`)
  vbnmf_updateR()    R/bayesian.R:56-106          (the pure-R twin, reciprocal bew/beh)
  hyper_update()     R/bayesian.R:2-53
  vb_iterate()       R/bayesian.R:303-390         (one run over all ranks)
  vb_factorize()     R/bayesian.R:265-299         (best run per rank, measure table)
  nmf_updateR()      R/factorize.R:2-27
  likelihood()       R/factorize.R:40-49
  ml_iterate()       R/factorize.R:187-212        (inner loop of factorize(), criterion='likelihood')
  cluster_id()       R/utils.R:903-909

Parity pin: vbnmf_update() here is checked against the reference's own src/vbnmf_update.cpp,
compiled in place into oracle/_ref (stand-in Eigen/Rcpp/GSL headers), and against the golden
vectors that binary produced (tests/golden/).  Everything that exists only as R source
(hyper_update, vb_iterate, the ML path) has no executable reference here (no R interpreter):
PARITY UNPINNED for those, restated line by line.

Third-party arithmetic: GSL gsl_sf_psi / gsl_sf_lngamma and R's digamma / psigamma / lgamma are
restated with scipy.special.digamma / polygamma / gammaln (double precision contracts).
Arrays are float64; R's column-major layout is irrelevant to NumPy semantics here.
"""
import numpy as np
from scipy.special import digamma, gammaln, polygamma

EPS = float(np.finfo(np.float64).eps)  # .Machine$double.eps


def vbnmf_update(X, wh, hyper, fudge):
    """src/vbnmf_update.cpp:16-102, statement by statement."""
    X = np.asarray(X, dtype=np.float64)
    fud = float(fudge)
    n, m = X.shape                                              # :20-21
    lw = np.array(wh["lw"], dtype=np.float64)                   # :22-25
    lh = np.array(wh["lh"], dtype=np.float64)
    eh = np.array(wh["eh"], dtype=np.float64)
    r = lw.shape[1]                                             # :27
    aw, ah, bw, bh = (float(hyper[k]) for k in ("aw", "ah", "bw", "bh"))  # :28-31

    wth = lw @ lh                                               # :33
    xwh = X / wth                                               # :34
    sw = lw * (xwh @ lh.T)                                      # :35
    sh = lh * (lw.T @ xwh)                                      # :36

    alw = aw + sw                                               # :38-39
    bew = np.full((n, r), aw / bw) + eh.sum(axis=1)[None, :]    # :40-43 (old eh)
    ew = alw / bew                                              # :44
    dw = alw / bew / bew                                        # :46

    alh = ah + sh                                               # :48-49
    beh = np.full((r, m), ah / bh) + ew.sum(axis=0)[:, None]    # :50-53 (new ew)
    eh = alh / beh                                              # :54
    dh = alh / beh / beh                                        # :56

    lw = np.maximum(np.exp(digamma(alw)) / bew, fud)            # :58-61
    lh = np.maximum(np.exp(digamma(alh)) / beh, fud)            # :62-65

    wth = lw @ lh                                               # :67
    A = (lw * np.log(lw)) @ lh                                  # :69-70
    B = lw @ (lh * np.log(lh))                                  # :71-72
    lwth = np.log(wth)                                          # :73
    U1 = (A + B) / wth - lwth                                   # :74-76
    U1 = X * U1                                                 # :77
    U1 = -ew @ eh - U1                                          # :78
    U = float(np.sum(U1 - gammaln(X + 1.0)))                    # :79-81
    lga = -gammaln(aw) + aw * np.log(aw / bw)                   # :82
    U += float(np.sum(-(aw / bw) * ew + lga + alw * (1.0 - np.log(bew)) + gammaln(alw)))  # :84-86
    lga = -gammaln(ah) + ah * np.log(ah / bh)                   # :87
    U += float(np.sum(-(ah / bh) * eh + lga + alh * (1.0 - np.log(beh)) + gammaln(alh)))  # :88-89
    # :90 divides by the int product n*m (overflows past 2^31-1); identical below that.
    U /= float(n) * float(m)
    return dict(w=ew, h=eh, lw=lw, lh=lh, ew=ew, eh=eh, lkh=U, dw=dw, dh=dh)  # :92-100


def vbnmf_updateR(x, wh, r, hyper, fudge=None):
    """R/bayesian.R:56-106 (pure-R twin; bew/beh held as reciprocals)."""
    x = np.asarray(x, dtype=np.float64)
    n, m = x.shape
    lw, lh = np.array(wh["lw"], float), np.array(wh["lh"], float)
    eh = np.array(wh["eh"], float)
    aw, bw, ah, bh = (float(hyper[k]) for k in ("aw", "bw", "ah", "bh"))
    wth = lw @ lh                                               # :71
    sw = lw * ((x / wth) @ lh.T)                                # :72
    sh = lh * (lw.T @ (x / wth))                                # :73
    alw = aw + sw                                               # :75
    bew = 1.0 / (aw / bw + eh.sum(axis=1))[None, :].repeat(n, 0)  # :76
    ew = alw * bew                                              # :77
    alh = ah + sh                                               # :79
    beh = 1.0 / (ah / bh + ew.sum(axis=0))[:, None].repeat(m, 1)  # :80
    eh = alh * beh                                              # :81
    lw = np.exp(digamma(alw)) * bew                             # :83
    lh = np.exp(digamma(alh)) * beh                             # :84
    if fudge is None:
        fudge = EPS                                             # :85
    lw[lw < fudge] = fudge                                      # :86-87
    lh[lh < fudge] = fudge
    wth = lw @ lh                                               # :89
    U1 = -ew @ eh - gammaln(x + 1) - x * ((((lw * np.log(lw)) @ lh) + lw @ (lh * np.log(lh))) / wth
                                          - np.log(wth))        # :90-91
    U2 = -(aw / bw) * ew - gammaln(aw) + aw * np.log(aw / bw) + alw * (1 + np.log(bew)) + gammaln(alw)
    U3 = -(ah / bh) * eh - gammaln(ah) + ah * np.log(ah / bh) + alh * (1 + np.log(beh)) + gammaln(alh)
    U = (U1.sum() + U2.sum() + U3.sum()) / (float(n) * float(m))  # :96-97
    dw = alw * bew ** 2                                         # :102
    dh = alh * beh ** 2                                         # :103
    return dict(w=ew, h=eh, lw=lw, lh=lh, ew=ew, eh=eh, lkh=float(U), dw=dw, dh=dh)


class HyperUpdateError(RuntimeError):
    """'Hyper-parameter update failed to converge' (R/bayesian.R:43)."""


def hyper_update(hyper_update_flags, wh, hyper, Niter=100, Tol=1e-4):
    """R/bayesian.R:2-53."""
    hu = [bool(v) for v in hyper_update_flags]
    if sum(hu) == 0:                                            # :4
        return dict(hyper)
    aw0, ah0 = float(hyper["aw"]), float(hyper["ah"])           # :6-7
    lwm = float(np.mean(np.log(wh["lw"])))                      # :8-11
    lhm = float(np.mean(np.log(wh["lh"])))
    ewm = float(np.mean(wh["ew"]))
    ehm = float(np.mean(wh["eh"]))
    bw0, bh0 = float(hyper["bw"]), float(hyper["bh"])           # :12-13
    if hu[0] + hu[2] > 0:                                       # :15
        i = 1
        while i < Niter:                                        # :17
            if hu[0]:
                dw = (np.log(aw0) - digamma(aw0) - ewm / bw0 + 1 + lwm - np.log(bw0)) / \
                     (1 / aw0 - polygamma(1, aw0))              # :19-20
            else:
                dw = 0.0
            if hu[2]:
                dh = (np.log(ah0) - digamma(ah0) - ehm / bh0 + 1 + lhm - np.log(bh0)) / \
                     (1 / ah0 - polygamma(1, ah0))              # :23-24
            else:
                dh = 0.0
            aw1 = aw0 - dw                                      # :26-27
            ah1 = ah0 - dh
            while aw1 <= 0:                                     # :28-31
                dw = dw / 2
                aw1 = aw0 - dw
            while ah1 <= 0:                                     # :32-35
                dh = dh / 2
                ah1 = ah0 - dh
            df = (1 - aw1 / aw0) ** 2 + (1 - ah1 / ah0) ** 2    # :37
            if df < Tol:                                        # :38
                break
            aw0, ah0 = aw1, ah1                                 # :39-40
            i += 1
        if i == Niter:                                          # :43
            raise HyperUpdateError("Hyper-parameter update failed to converge")
    else:
        aw1, ah1 = aw0, ah0                                     # :44-47
    bw1 = ewm if hu[1] else bw0                                 # :48-49
    bh1 = ehm                                                   # :50-51 (both branches)
    return dict(aw=float(aw1), bw=float(bw1), ah=float(ah1), bh=float(bh1))


def vb_init_from(w, h):
    """The list vb_init returns (R/bayesian.R:161-170) for given w, h."""
    w = np.array(w, float)
    h = np.array(h, float)
    return dict(w=w, h=h, lw=w.copy(), lh=h.copy(), ew=w.copy(), eh=h.copy(),
                dw=np.zeros_like(w), dh=np.zeros_like(h))


def vb_run_one_rank(mat, w0, h0, hyper0, *, Itmax=10000, hyper_update_flags=(True,) * 4,
                    Tol=1e-5, n0=10, dn=1, fudge=EPS, update=vbnmf_update):
    """Loop body of vb_iterate for one rank: R/bayesian.R:333-352.
    Returns (wh, hyper, lk0, it, lkh_trace, hyper_trace, stop_reason)."""
    hyper = dict(hyper0)
    wh = vb_init_from(w0, h0)                                   # :334
    lk0 = 0.0                                                   # :336
    trace, htrace = [], []
    reason = 0
    it = 0
    for it in range(1, int(Itmax) + 1):                         # :337
        if update is vbnmf_updateR:
            wh = vbnmf_updateR(mat, wh, w0.shape[1], hyper, fudge=fudge)
        else:
            wh = update(mat, wh, hyper, fudge)                  # :339
        if it > n0 and it % dn == 0:                            # :342-344
            hyper = hyper_update(hyper_update_flags, wh, hyper, Niter=100, Tol=1e-3)
        trace.append(wh["lkh"])
        htrace.append([hyper["aw"], hyper["bw"], hyper["ah"], hyper["bh"]])
        if np.isnan(wh["lkh"]):                                 # :345
            reason = 2
            break
        if it > 1 and it > n0 and wh["lkh"] >= lk0 and abs(1 - wh["lkh"] / lk0) < Tol:  # :346-347
            reason = 1
            break
        lk0 = wh["lkh"]                                         # :348
    return wh, hyper, lk0, it, np.array(trace), np.array(htrace), reason


def uniform_columns(ew, Tol):
    """R/bayesian.R:368-369: columns of ew with |max - min| < Tol."""
    ew = np.asarray(ew)
    return np.abs(ew.max(axis=0) - ew.min(axis=0)) < Tol


def vb_iterate(mat, ranks, inits, *, gamma_a=1.0, gamma_b=1.0, unif_stop=True, **kw):
    """R/bayesian.R:303-390 for one run.  inits[irank] = (w0, h0) replaces vb_init's RNG draw.
    Returns dict(rdat, wdat, hdat, dwdat, dhdat, hyperp, nunif, niter)."""
    ga = np.atleast_1d(np.asarray(gamma_a, float))
    gb = np.atleast_1d(np.asarray(gamma_b, float))
    nrank = len(ranks)
    rdat = [-np.inf] * nrank                                    # :309
    out = dict(rdat=rdat, wdat=[None] * nrank, hdat=[None] * nrank, dwdat=[None] * nrank,
               dhdat=[None] * nrank, hyperp=[None] * nrank, nunif=[0] * nrank,
               niter=[0] * nrank)
    Tol = kw.get("Tol", 1e-5)
    nrow, ncol = mat.shape
    for irank, rank in enumerate(ranks):                        # :316
        if rank > min(nrow, ncol):                              # :319-320
            raise ValueError("Rank exceeded min(nrow,ncol)")
        hyper0 = dict(aw=ga[0], ah=ga[-1], bw=gb[0], bh=gb[-1])  # :321-326
        w0, h0 = inits[irank]
        wh, hyper, lk0, it, _, _, _ = vb_run_one_rank(mat, w0, h0, hyper0, **kw)
        cu = uniform_columns(wh["ew"], Tol)                     # :368-369
        if cu.sum() > 0 and unif_stop:                          # :370-378
            if irank == 0:
                raise RuntimeError("Rerun with lower ranks")
            break
        rdat[irank] = lk0                                       # :379-384
        out["wdat"][irank] = wh["ew"]
        out["hdat"][irank] = wh["eh"]
        out["dwdat"][irank] = np.sqrt(wh["dw"])
        out["dhdat"][irank] = np.sqrt(wh["dh"])
        out["hyperp"][irank] = hyper
        out["niter"][irank] = it
    return out


def vb_select(vb_runs, ranks):
    """R/bayesian.R:265-299: per rank, the run with the largest rdat (strict >, first wins)."""
    res = dict(ranks=[], lml=[], basis=[], coeff=[], dbasis=[], dcoeff=[], aw=[], bw=[], ah=[],
               bh=[], nunif=[], run=[])
    for k in range(len(ranks)):
        rmax, imax = -np.inf, None
        for i, vb in enumerate(vb_runs):
            if vb["rdat"][k] > rmax:                            # :271
                imax, rmax = i, vb["rdat"][k]
        if rmax == -np.inf:                                     # :276
            continue
        vb = vb_runs[imax]
        res["ranks"].append(ranks[k]); res["lml"].append(rmax); res["run"].append(imax)
        res["basis"].append(vb["wdat"][k]); res["coeff"].append(vb["hdat"][k])
        res["dbasis"].append(vb["dwdat"][k]); res["dcoeff"].append(vb["dhdat"][k])
        for key in ("aw", "bw", "ah", "bh"):
            res[key].append(vb["hyperp"][k][key])
        res["nunif"].append(vb["nunif"][k])
    return res


def nmf_updateR(x, w, h):
    """R/factorize.R:2-27 with prior=FALSE (the only call site, :192)."""
    x = np.asarray(x, float)
    w = np.array(w, float)
    h = np.array(h, float)
    up = h * (w.T @ (x / (w @ h)))                              # :8
    down = w.sum(axis=0)[:, None]                               # :9
    h = up / down                                               # :14
    h[h < EPS] = EPS                                            # :15
    up = w * ((x / (w @ h)) @ h.T)                              # :17
    down = h.sum(axis=1)[None, :]                               # :18
    w = up / down                                               # :23
    w[w < EPS] = EPS                                            # :24
    return dict(ew=w, eh=h)


def likelihood(mat, w, h):
    """R/factorize.R:40-49."""
    wh = (w @ h).ravel()
    amat = np.asarray(mat, float).ravel()
    x = np.sum(amat * np.log(wh) - wh)                          # :44
    z = amat[amat > 0]
    x = x + np.sum(-z * np.log(z) + z)                          # :45-46
    return float(x / mat.shape[0] / mat.shape[1])               # :47


def connectivity(h):
    """R/factorize.R:51-60, literally: the m(m-1)/2 vector of outer(cid, cid, '==') below the
    diagonal (small m only)."""
    cid = np.argmax(np.asarray(h), axis=0)
    cnn = cid[:, None] == cid[None, :]
    return cnn.T[np.tril_indices(len(cid), -1)[::-1]].astype(float)


def dispersion(cnn, nc):
    """R/factorize.R:62-67."""
    return 1.0 / nc + 8.0 * float(np.sum((cnn - 0.5) ** 2)) / nc ** 2


def cophenet(conav, nc, method="average"):
    """R/factorize.R:69-78 with scipy standing in for stats::hclust / cophenetic / cor."""
    from scipy.cluster.hierarchy import cophenet as sc_cophenet, linkage
    tmp = np.zeros((nc, nc))
    tmp[np.tril_indices(nc, -1)[::-1]] = 1.0 - conav
    d = (tmp + tmp.T)[np.triu_indices(nc, 1)]
    z = linkage(d, method=method)
    return float(np.corrcoef(d, sc_cophenet(z))[0, 1])


def ml_iterate(mat, w0, h0, *, Itmax=10000, Tol=1e-5, criterion="likelihood", ncnn_step=40):
    """R/factorize.R:187-212, both stopping criteria.  Returns (w, h, lk0, it, trace) and, for
    criterion='connectivity', the per-iteration nchange as a sixth element."""
    wh = dict(ew=np.array(w0, float), eh=np.array(h0, float))
    lkold = -np.inf                                             # :190
    zstep = 0                                                   # :189
    trace, nch = [], []
    it = 0
    lk0 = np.nan
    ncol = np.asarray(mat).shape[1]
    npair = ncol * (ncol - 1) / 2
    cnn0 = None
    for it in range(1, int(Itmax) + 1):                         # :191
        wh = nmf_updateR(mat, wh["ew"], wh["eh"])               # :192
        lk0 = likelihood(mat, wh["ew"], wh["eh"])               # :193
        trace.append(lk0)
        if criterion == "connectivity":                         # :194-204
            cnn = connectivity(wh["eh"])
            nchange = npair if it == 1 else float(np.sum(cnn != cnn0))
            nch.append(nchange)
            zstep = zstep + 1 if nchange == 0 else 0
            if zstep == ncnn_step:
                break
            cnn0 = cnn
        elif criterion == "likelihood":
            if abs(lkold - lk0) < Tol * abs(lkold):             # :207
                break
            lkold = lk0                                         # :209
        else:
            raise ValueError("Unknown stopping criterion.")     # :212
    if criterion == "connectivity":
        return wh["ew"], wh["eh"], lk0, it, np.array(trace), np.array(nch)
    return wh["ew"], wh["eh"], lk0, it, np.array(trace)


def cluster_id(h):
    """R/utils.R:903-909: apply(h, 2, which.max) -> 1-based index of the FIRST maximum."""
    return np.argmax(np.asarray(h), axis=0).astype(np.int32) + 1
