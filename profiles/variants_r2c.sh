#!/bin/bash
run() { echo "== $*"; env "${@:2}" python profiles/prof_run.py --workload $1 2>&1 | grep -v "^\[vbnmf" | cut -c1-420; }
run "c3 --cells 200000 --iters 10" VBNMF_SEG_WINDOW=4096
run "c3 --cells 200000 --iters 10" VBNMF_SEG_WINDOW=8192
run "c3 --cells 200000 --iters 10" VBNMF_SEG_WINDOW=4096 VBNMF_LIB_NAME=libvbnmf_S0.so
run "c3 --cells 200000 --iters 10" VBNMF_SEG_WINDOW=8192 VBNMF_LIB_NAME=libvbnmf_S0.so
run "c3 --cells 200000 --iters 10 --precision 1" VBNMF_SEG_WINDOW=4096
run "c3 --cells 200000 --iters 10 --precision 1" VBNMF_SEG_WINDOW=0
