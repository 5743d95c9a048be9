#!/bin/bash
# A/B of build knobs on a C3-shaped matrix (20k genes x 200k cells, r = 20)
P="python profiles/prof_run.py --workload c3 --cells 200000 --iters 10"
echo "== default"; $P
echo "== b: LP_BITS=2"; VBNMF_LIB_NAME=libvbnmf_b.so $P
echo "== c: 448 threads"; VBNMF_LIB_NAME=libvbnmf_c.so $P
echo "== d: LP_BITS=4"; VBNMF_LIB_NAME=libvbnmf_d.so $P
