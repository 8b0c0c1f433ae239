#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s -x -k "gwnet" > gpurun_out/r02x_gw.log 2>&1
grep -E "passed|failed|FAILED|Error|^E  " gpurun_out/r02x_gw.log | tail -12 | cut -c1-250
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r02x_bench.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for r in d["roofline_other"]:
    if "G2" in r["kernel"]: print("  ", r["kernel"][:40], r["ms"], r["frac"])
P
tail -3 gpurun_out/r02x_bench.err | cut -c1-300
