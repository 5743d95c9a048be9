#!/bin/bash
# Round-2 final profiling pass (one GPU).  Every ncu capture follows a plain run of the same command.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/r02_gputests_1gpu.txt
python bench.py > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err
# 1. launch list of one bench run (gpu__time_duration per launch, cold caches, serialised)
python bench.py --steps 2 --warmup 3 --no-e2e --parity-iters 0 --no-cpu --secondary none --workload c2 > $O/r02_ll_plain.json 2> $O/r02_ll_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launch_list_c2.csv \
    -k regex:'sweep_|combine_kernel|posterior_kernel|control_kernel|make_keys|segment_ptr|plan_p16|build_segments|seg_order|seg_permute|split_|tag_dead|expand_cols|count_constants|order_keys|deal_kernel|scatter_panel|gather_panel|panel_colsum|mirror_kernel|cluster_id|DeviceRadixSort|DeviceScan' \
    python bench.py --steps 2 --warmup 3 --no-e2e --parity-iters 0 --no-cpu --secondary none --workload c2 > $O/r02_ll_ncu.log 2>&1
# 2. DRAM traffic of the two sweep launches at the bench sizes (one pass, two metrics)
for w in c2 c3; do
  python profiles/prof_run.py --workload $w --iters 2 > $O/r02_prof_$w.log 2>&1 &&
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:sweep_ -s 8 -c 2 --csv --log-file $O/r02_traffic_$w.csv \
      python profiles/prof_run.py --workload $w --iters 2 > $O/r02_traffic_$w.log 2>&1
done
# 3. full captures of the sweep kernels: C3-shaped (r = 20, 200k cells) and C2 (r = 10), fp64; C2 fp32-storage
ncu --set full --clock-control none --import-source on -k regex:sweep_p16 -s 8 -c 2 -o $O/r02_sweep_r20_fp64 -f \
    python profiles/prof_run.py --workload c3 --cells 200000 --iters 2 > $O/r02_ncu_r20.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_p16 -s 8 -c 2 -o $O/r02_sweep_r10_fp64 -f \
    python profiles/prof_run.py --workload c2 --iters 2 > $O/r02_ncu_r10.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sweep_p16 -s 8 -c 2 -o $O/r02_sweep_r10_fp32 -f \
    python profiles/prof_run.py --workload c2 --precision 1 --iters 2 > $O/r02_ncu_r10_32.log 2>&1
# 4. BASELINE config 4 to convergence on one GPU
python profiles/c4_converged.py > $O/r02_c4_converged_1gpu.json 2> $O/r02_c4_converged_1gpu.err
ls -la $O/*.ncu-rep $O/r02_*.csv
