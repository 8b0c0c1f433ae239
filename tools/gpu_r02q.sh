#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02q_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r02q_pytest.log | tail -10
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r02q_bench.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["ms"])
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02q_launches.csv \
  python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02q_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02q_launches.csv 4 40 > gpurun_out/r02q_launches_summary.txt 2>&1
head -24 gpurun_out/r02q_launches_summary.txt
