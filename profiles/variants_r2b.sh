#!/bin/bash
# A/B of owner-row request timing / L2 prefetch / raw pointers (libvbnmf_<X>.so built with VBNMF_DEFS)
O=gpurun_out
for lib in libvbnmf_t.so libvbnmf.so libvbnmf_F.so libvbnmf_G.so libvbnmf_H.so libvbnmf_t.so; do
  [ -f ccfindr_b200/$lib ] || continue
  for w in "c3 --cells 200000 --iters 10" "c2 --iters 20"; do
    echo "== $lib $w"
    VBNMF_LIB_NAME=$lib python profiles/prof_run.py --workload $w 2>&1 | grep -v "^\[vbnmf" | cut -c1-330
  done
done
