#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -x -k "fused_ffn" > gpurun_out/r02m_ffn.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r02m_ffn.log | tail -8
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02m_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r02m_pytest.log | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
tail -c 300 gpurun_out/r02m_bench.json; tail -3 gpurun_out/r02m_bench.err
STEP_B200_FFN_FUSED=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02m_bench_ffnsplit.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02m_launches.csv \
  python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02m_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02m_launches.csv 4 40 > gpurun_out/r02m_launches_summary.txt 2>&1
head -16 gpurun_out/r02m_launches_summary.txt
