set -x
O=gpurun_out
for w in c2 c3; do
  python profiles/prof_run.py --workload $w --iters 2 > $O/r02_prof_$w.log 2>&1 &&
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k regex:sweep_ -s 8 -c 2 --csv --log-file $O/r02_traffic_$w.csv \
      python profiles/prof_run.py --workload $w --iters 2 > $O/r02_traffic_$w.log 2>&1
done
