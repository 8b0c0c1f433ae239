#!/bin/bash
# round-2 call h: one-pass bounded softmax + bf16x2 dropout masks in tc_attn_kernel; ncu detail of attention and the fused GWNet layer
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_dropout.py -m gpu -q -s > gpurun_out/r02h_tc.log 2>&1
grep -E "passed|failed|FAILED|Error" gpurun_out/r02h_tc.log | tail -15
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02h_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r02h_pytest.log | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
tail -c 300 gpurun_out/r02h_bench.json; tail -3 gpurun_out/r02h_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02h_launches.csv \
  python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02h_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02h_launches.csv 4 40 > gpurun_out/r02h_launches_summary.txt 2>&1
head -16 gpurun_out/r02h_launches_summary.txt
timeout 600 ncu --set full --clock-control none -k regex:'tc_attn_kernel' -s 4 -c 1 -o /tmp/r02h_attn -f \
  python bench.py --steps 1 --warmup 2 --only-resident > gpurun_out/r02h_ncu_attn.log 2>&1
python tools/ncu_summary.py /tmp/r02h_attn.ncu-rep > gpurun_out/r02h_ncu_attn.txt 2>&1
ncu -i /tmp/r02h_attn.ncu-rep --page details > gpurun_out/r02h_ncu_attn_details.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'gw_fused_fwd_kernel' -s 7 -c 2 -o /tmp/r02h_gwf -f \
  python bench.py --steps 1 --warmup 2 --only-resident > gpurun_out/r02h_ncu_gwf.log 2>&1
python tools/ncu_summary.py /tmp/r02h_gwf.ncu-rep > gpurun_out/r02h_ncu_gwf.txt 2>&1
ncu -i /tmp/r02h_gwf.ncu-rep --page details > gpurun_out/r02h_ncu_gwf_details.txt 2>&1
ls -la gpurun_out | tail -12
