#!/bin/bash
# threads per CTA of the 4-lane split kernels (rows of 97..160 bytes): 384 (default) against 352 / 416
run() { echo "== $*"; env "${@:2}" python profiles/prof_run.py --workload $1 2>&1 | grep -v "^\[vbnmf" | cut -c1-420; }
# (build the variants first: VBNMF_MID_THREADS=352 VBNMF_LIB_NAME=libvbnmf_T352.so VBNMF_OBJ_SUFFIX=_T352 python -m ccfindr_b200.build ...)
for lib in libvbnmf.so libvbnmf_T352.so libvbnmf_T416.so libvbnmf_C352.so; do
  [ -f ccfindr_b200/$lib ] || continue
  run "c3 --cells 200000 --iters 10" VBNMF_LIB_NAME=$lib
done
