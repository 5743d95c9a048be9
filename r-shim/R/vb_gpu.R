# Drop-in for the hot loop of ccfindR::vb_factorize.  SOURCE ONLY (R is not installed where this
# repository is built and tested); the same C ABI is exercised from Python in tests/.
#
# vb_iterate_gpu() is R/bayesian.R:303-390 with the `for(it in seq_len(bundle$Itmax))` loop
# (:337-352: vbnmf_update + hyper_update + convergence test) replaced by ONE .Call on a
# device-resident handle.  Everything outside that loop is the reference's code path: vb_init,
# the uniform-column rule, the per-run bookkeeping and vb_factorize's best-run selection.

vbnmf_handle <- function(mat, device = 0L) {
  mat <- methods::as(mat, "CsparseMatrix")          # dgCMatrix: @p, @i, @x, @Dim (no as.matrix())
  .Call(C_vbnmf_create, mat@p, mat@i, as.numeric(mat@x), mat@Dim, as.integer(device))
}

vb_iterate_gpu <- function(irun, bundle) {
  nrow <- dim(bundle$mat)[1]; ncol <- dim(bundle$mat)[2]
  nrank <- length(bundle$ranks)
  rdat <- rep(-Inf, nrank)
  wdat <- hdat <- dwdat <- dhdat <- hyperp <- list()
  nunif <- rep(0, nrank)
  h <- if (is.null(bundle$handle)) vbnmf_handle(bundle$mat, bundle$device) else bundle$handle
  if (bundle$verbose >= 2) if (bundle$nrun > 1) cat('Run ', irun, '\n', sep = '')
  for (irank in seq_len(nrank)) {
    rank <- bundle$ranks[[irank]]
    if (rank > min(nrow, ncol)) stop('Rank exceeded min(nrow,ncol)')
    hyper <- list(aw = bundle$gamma.a[1], ah = bundle$gamma.a[length(bundle$gamma.a)],
                  bw = bundle$gamma.b[1], bh = bundle$gamma.b[length(bundle$gamma.b)])
    wh <- vb_init(nrow, ncol, bundle$mat, rank, hyper = hyper, initializer = bundle$initializer)
    .Call(C_vbnmf_set_state, h, wh$lw, wh$lh, wh$ew, wh$eh)
    res <- .Call(C_vbnmf_run, h, c(hyper$aw, hyper$bw, hyper$ah, hyper$bh),
                 as.integer(bundle$Itmax), bundle$Tol, as.logical(bundle$hyper.update),
                 as.integer(bundle$hyper.update.n0), as.integer(bundle$hyper.update.dn),
                 bundle$fudge)
    hyper <- list(aw = res$hyper[1], bw = res$hyper[2], ah = res$hyper[3], bh = res$hyper[4])
    lk0 <- res$lml; it <- res$niter
    if (bundle$verbose >= 3)
      for (i in seq_len(it)) cat(i, ', log(evidence) = ', res$lkh_trace[i], '\n', sep = '')
    if (bundle$verbose >= 2)
      cat('Rank = ', rank, ': Nsteps =', it, ', log(evidence) =', lk0, ', hyper = (', hyper$aw, ',',
          hyper$bw, ',', hyper$ah, ',', hyper$bh, ')\n', sep = '')
    contains.unif <- .Call(C_vbnmf_uniform_columns, h, bundle$Tol)
    if (sum(contains.unif) > 0) {
      warning('Rank ', rank, ' row/column ', paste(which(contains.unif), collapse = ','), ' constant.')
      if (bundle$unif.stop) {
        warning('Rank scan stopped for rank >= ', rank)
        if (irank == 1) stop('Rerun with lower ranks')
        break
      }
    }
    st <- .Call(C_vbnmf_get_state, h, NULL)
    rdat[irank] <- lk0
    wdat[[irank]] <- st$ew; hdat[[irank]] <- st$eh
    dwdat[[irank]] <- sqrt(st$dw); dhdat[[irank]] <- sqrt(st$dh)
    hyperp[[irank]] <- hyper
  }
  list(rdat = rdat, wdat = wdat, hdat = hdat, hyperp = hyperp, nunif = nunif,
       dwdat = dwdat, dhdat = dhdat)
}
