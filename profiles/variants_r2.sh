#!/bin/bash
# A/B of the 4-unit split layout (ranks 8..14, fp64) against the lock-step layout
for w in "c2 --iters 20"; do
  echo "== $w: split (default)"; python profiles/prof_run.py --workload $w
  echo "== $w: lock-step (VBNMF_NO_SPLIT4=1)"; VBNMF_NO_SPLIT4=1 python profiles/prof_run.py --workload $w
done
