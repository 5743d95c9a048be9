#!/bin/bash
# time per iteration against the tile height (segments per pass ~ 1/T): per-segment overhead
for T in 1440 1280 1152 960 800 640 480; do
  echo "== T=$T"; VBNMF_TILE_ROWS=$T python profiles/prof_run.py --workload c3 --cells 200000 --iters 10 2>&1 | grep -v "^\[vbnmf" | cut -c1-420
done
