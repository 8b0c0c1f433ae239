#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -k "trunk_conv" > gpurun_out/r02s_trunk.log 2>&1
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/r02s_trunk.log | tail -12
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -m gpu -q > gpurun_out/r02s_model.log 2>&1
tail -3 gpurun_out/r02s_model.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r02s_bench.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["ms"])
for r in d["roofline_other"]:
    print("  ", r["kernel"][:60], r.get("ms"), r.get("frac"))
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02s_launches.csv \
  python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02s_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02s_launches.csv 4 40 > gpurun_out/r02s_launches_summary.txt 2>&1
grep -n "trunk\|launches in" gpurun_out/r02s_launches_summary.txt
