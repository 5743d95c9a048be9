python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload c2 --secondary none > gpurun_out/r2_bench_c2only.json 2> gpurun_out/r2_bench_c2only.err; tail -c 300 gpurun_out/r2_bench_c2only.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_c2only.json").read().strip().splitlines()[-1])
print(json.dumps(d["ml_path"])[:1200])
print(d["ms_per_step"], d["roofline"]["ms_per_launch"], d["e2e"]["seconds"], d["parity"]["lkh_rel_err"])
PY
python profiles/prof_run.py --workload c3 --cells 200000 --iters 10 2>&1 | grep -v "^\[vbnmf" | cut -c1-200
