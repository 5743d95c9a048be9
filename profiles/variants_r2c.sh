#!/bin/bash
# owner rows staged through shared memory (default, 8-lane kernels) against the request before the
# previous segment's cross-lane sum (libvbnmf_S0.so, VB_OWN_STAGE=0), over ranks (C2 matrix)
run() { echo "== $*"; env "${@:2}" python profiles/prof_run.py --workload $1 2>&1 | grep -v "^\[vbnmf" | cut -c1-420; }
for r in 8 10 12 14 24 30; do
  for lib in libvbnmf.so libvbnmf_S0.so; do
    run "c2 --rank $r --iters 10" VBNMF_LIB_NAME=$lib
  done
done
