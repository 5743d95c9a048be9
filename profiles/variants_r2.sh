#!/bin/bash
# A/B of build knobs on a C3-shaped matrix (20k genes x 200k cells, r = 20)
P="python profiles/prof_run.py --workload c3 --cells 200000 --iters 10"
echo "== default (immediate slow path of the log-product)"; $P
echo "== b: cell-owner pass with 352 threads"; VBNMF_LIB_NAME=libvbnmf_b.so $P
echo "== c: per-chunk slow path (round-2 baseline)"; VBNMF_LIB_NAME=libvbnmf_c.so $P
echo "== d: 352 threads, per-chunk slow path"; VBNMF_LIB_NAME=libvbnmf_d.so $P
