#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -x -k "fused_ffn" > gpurun_out/r02n_ffn.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r02n_ffn.log | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
tail -c 200 gpurun_out/r02n_bench.json; tail -3 gpurun_out/r02n_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02n_launches.csv \
  python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02n_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02n_launches.csv 4 40 > gpurun_out/r02n_launches_summary.txt 2>&1
head -12 gpurun_out/r02n_launches_summary.txt
