#!/bin/bash
# round-2 baseline capture at HEAD: bench line, then one --set full pass over every GWNet / trunk kernel of a step
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'gw_layer|tc_mix_kernel|tc_dP_kernel|trunk_conv2_bwd|trunk_conv1_bwd|gw_skip' -c 96 \
  -o gpurun_out/r02a_gw -f python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02a_ncu.log 2>&1
ls -la gpurun_out
