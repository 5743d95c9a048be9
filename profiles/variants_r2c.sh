#!/bin/bash
# current build (+ variants) against the round-2 baseline (libvbnmf_t.so = commit 58af0c5)
run() { echo "== $*"; env "${@:2}" python profiles/prof_run.py --workload $1 2>&1 | grep -v "^\[vbnmf" | cut -c1-420; }
for lib in libvbnmf_t.so libvbnmf.so libvbnmf_S0.so; do
  [ -f ccfindr_b200/$lib ] || continue
  run "c3 --cells 200000 --iters 10" VBNMF_LIB_NAME=$lib
  run "c2 --iters 20" VBNMF_LIB_NAME=$lib
  run "c2 --iters 20 --precision 1" VBNMF_LIB_NAME=$lib
done
