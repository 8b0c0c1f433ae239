#!/bin/bash
# round-2 call d: GEMM chain debug, fused token-block kernel, dropout / training tests, whole suite, bench, launch list
set -x
mkdir -p gpurun_out
timeout 300 python tools/debug_gemm.py > gpurun_out/r02d_gemm_debug.log 2>&1
tail -70 gpurun_out/r02d_gemm_debug.log
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s -k "fused_token or bf16_close" > gpurun_out/r02d_fused.log 2>&1
tail -15 gpurun_out/r02d_fused.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02d_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r02d_pytest.log | tail -25
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
tail -c 600 gpurun_out/r02d_bench.json; tail -5 gpurun_out/r02d_bench.err
STEP_B200_TS_FUSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02d_bench_unfused.json 2> gpurun_out/r02d_bench_unfused.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02d_launches.csv \
  python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02d_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02d_launches.csv 4 30 > gpurun_out/r02d_launches_summary.txt 2>&1
head -34 gpurun_out/r02d_launches_summary.txt
