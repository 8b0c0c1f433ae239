#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -x -k "fused_ffn and 2-9-168-0.0" > gpurun_out/r02k_ffn_memcheck.log 2>&1
grep -n "Invalid\|Error\|error\|at 0x\|by thread\|Address" gpurun_out/r02k_ffn_memcheck.log | head -30
STEP_B200_FFN_FUSED=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:'tc_attn_kernel' -s 4 -c 1 -o /tmp/r02k_attn -f \
  python bench.py --steps 1 --warmup 2 --only-resident > gpurun_out/r02k_ncu_attn.log 2>&1
python tools/ncu_summary.py /tmp/r02k_attn.ncu-rep > gpurun_out/r02k_ncu_attn.txt 2>&1
ncu -i /tmp/r02k_attn.ncu-rep --page source --csv > gpurun_out/r02k_ncu_attn_source.csv 2>&1
ls -la gpurun_out | tail -5
