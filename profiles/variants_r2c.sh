#!/bin/bash
run() { echo "== $*"; env "${@:2}" python profiles/prof_run.py --workload $1 2>&1 | grep -v "^\[vbnmf" | cut -c1-420; }
run "c2 --rank 18 --iters 10" X=1
run "c3 --cells 200000 --rank 18 --iters 10" X=1
run "c3 --cells 200000 --iters 10" X=1
run "c2 --rank 16 --iters 10" X=1
