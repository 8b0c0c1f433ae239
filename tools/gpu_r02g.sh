#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "fused_layer_kernel" > gpurun_out/r02g_fused.log 2>&1
tail -15 gpurun_out/r02g_fused.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02g_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r02g_pytest.log | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
tail -c 400 gpurun_out/r02g_bench.json; tail -3 gpurun_out/r02g_bench.err
STEP_B200_GW_FUSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02g_bench_gwsplit.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02g_launches.csv \
  python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02g_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02g_launches.csv 4 40 > gpurun_out/r02g_launches_summary.txt 2>&1
head -30 gpurun_out/r02g_launches_summary.txt
