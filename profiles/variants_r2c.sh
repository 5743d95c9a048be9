#!/bin/bash
# dense block B (rows of exactly RP doubles) + four-class schedule for ranks 19/20 against the round-2 baseline
run() { echo "== $*"; env "$@" python profiles/prof_run.py --workload c3 --cells 200000 --iters 10 2>&1 | grep -v "^\[vbnmf" | cut -c1-420; }
run VBNMF_LIB_NAME=libvbnmf_t.so
run VBNMF_LIB_NAME=libvbnmf.so
run VBNMF_LIB_NAME=libvbnmf.so VBNMF_NO_CLS4=1
