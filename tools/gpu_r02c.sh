#!/bin/bash
# round-2 call c: new kernels (gemm, G1/G3, pre-training backward, fc v2), whole suite, new bench, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -s -k "gemm or prologue or trunk_fc or attention_backward or add_layernorm" > gpurun_out/r02c_new.log 2>&1
tail -25 gpurun_out/r02c_new.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02c_pytest.log 2>&1
tail -25 gpurun_out/r02c_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -c 1500 gpurun_out/r02c_bench.json; tail -5 gpurun_out/r02c_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches.csv \
  python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02c_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r02c_launches.csv 4 70 > gpurun_out/r02c_launches_summary.txt 2>&1
head -50 gpurun_out/r02c_launches_summary.txt
