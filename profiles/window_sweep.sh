#!/bin/bash
# window of the length-sorted segment order (0 = owner order)
for W in 2048 4096 6144 12288 24576 49152; do
  for w in "c3 --cells 200000 --iters 10" "c2 --iters 20"; do
    echo "== W=$W $w"; VBNMF_SEG_WINDOW=$W python profiles/prof_run.py --workload $w 2>&1 | grep -v "^\[vbnmf" | cut -c1-420
  done
done
