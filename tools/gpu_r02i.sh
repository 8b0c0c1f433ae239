#!/bin/bash
# round-2 call i: attention with bound prefetch; tc tests; bench; attention ncu detail + source-level stall sampling
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_dropout.py -m gpu -q -s > gpurun_out/r02i_tc.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r02i_tc.log | tail -15
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02i_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r02i_pytest.log | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
tail -c 300 gpurun_out/r02i_bench.json; tail -3 gpurun_out/r02i_bench.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'tc_attn_kernel' -s 4 -c 1 -o /tmp/r02i_attn -f \
  python bench.py --steps 1 --warmup 2 --only-resident > gpurun_out/r02i_ncu_attn.log 2>&1
python tools/ncu_summary.py /tmp/r02i_attn.ncu-rep > gpurun_out/r02i_ncu_attn.txt 2>&1
ncu -i /tmp/r02i_attn.ncu-rep --page details > gpurun_out/r02i_ncu_attn_details.txt 2>&1
ncu -i /tmp/r02i_attn.ncu-rep --page source --csv > gpurun_out/r02i_ncu_attn_source.csv 2>&1
ls -la gpurun_out | tail -8
