#!/bin/bash
# length-sorted windows against owner order over the ranks of the C4 sweep (C2 matrix)
for r in 2 4 6 8 14 18 24 30; do
  for W in 0 4096; do
    echo "== r=$r W=$W"; VBNMF_SEG_WINDOW=$W python profiles/prof_run.py --workload c2 --rank $r --iters 10 2>&1 | grep -v "^\[vbnmf" | cut -c1-420
  done
done
