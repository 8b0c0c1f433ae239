#!/bin/bash
# round-2 call b: new fc kernels + whole-model parity on every config, bench, compact ncu summaries of the GWNet / trunk kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "trunk_fc" -s > gpurun_out/r02b_fc.log 2>&1
tail -15 gpurun_out/r02b_fc.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02b_pytest.log 2>&1
tail -40 gpurun_out/r02b_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -c 600 gpurun_out/r02b_bench.json
timeout 600 ncu --set full --clock-control none \
  -k regex:'gw_layer|tc_mix_kernel|tc_dP_kernel|trunk_conv2_bwd|trunk_conv1_bwd|gw_skip|fc_fwd|fc_dx|fc_dw' -c 100 \
  -o /tmp/r02b_gw -f python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02b_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r02b_gw.ncu-rep > gpurun_out/r02b_ncu_gw_summary.txt 2>&1
ls -la /tmp/r02b_gw.ncu-rep gpurun_out
