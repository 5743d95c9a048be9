#!/bin/bash
# A/B of the split-layout sweep kernels on a C3-shaped matrix (20k genes x 200k cells, r = 20)
P="python profiles/prof_run.py --workload c3 --cells 200000 --iters 10"
echo "== default (split, dead steps skipped, 2 chains, 384 thr)"; $P
echo "== b: 4 dot chains"; VBNMF_LIB_NAME=libvbnmf_b.so $P
echo "== c: dead steps executed"; VBNMF_LIB_NAME=libvbnmf_c.so $P
echo "== d: 256 threads"; VBNMF_LIB_NAME=libvbnmf_d.so $P
echo "== e: 320 threads"; VBNMF_LIB_NAME=libvbnmf_e.so $P
echo "== f: 256 threads + 4 chains"; VBNMF_LIB_NAME=libvbnmf_f.so $P
