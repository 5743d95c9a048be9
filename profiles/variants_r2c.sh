#!/bin/bash
# 4-lane groups (default for ranks 15/16/19/20) against 8-lane groups (VBNMF_NO_G4=1)
run() { echo "== $*"; env "${@:2}" python profiles/prof_run.py --workload $1 2>&1 | grep -v "^\[vbnmf" | cut -c1-420; }
run "c3 --cells 200000 --iters 10" VBNMF_NO_G4=1
run "c3 --cells 200000 --iters 10" VBNMF_X=1
run "c3 --cells 200000 --iters 10" VBNMF_SEG_WINDOW=4096
run "c3 --cells 200000 --iters 10" VBNMF_SEG_WINDOW=1024
