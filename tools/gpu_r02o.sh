#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_dropout.py -m gpu -q -s > gpurun_out/r02o_tc.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r02o_tc.log | tail -8
timeout 300 python tools/attn_trace.py 6624 48 > gpurun_out/r02o_attn_trace.txt 2>&1
tail -4 gpurun_out/r02o_attn_trace.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
tail -c 200 gpurun_out/r02o_bench.json; tail -3 gpurun_out/r02o_bench.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r02o_bench.json")); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["ms"])
P
