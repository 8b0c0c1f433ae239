#!/bin/bash
# round-2 evidence run at HEAD (1 GPU): tests, full bench line, reference arm, launch list, per-step DRAM totals, ncu captures
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02u_gputest.log 2>&1
tail -3 gpurun_out/r02u_gputest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02u_bench_1gpu.json 2> gpurun_out/r02u_bench_1gpu.err
tail -c 300 gpurun_out/r02u_bench_1gpu.json; tail -2 gpurun_out/r02u_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02u_bench_reference.json 2> gpurun_out/r02u_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02u_launches.csv \
  python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02u_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r02u_launches.csv 4 70 > gpurun_out/r02u_launches_summary.txt 2>&1
head -8 gpurun_out/r02u_launches_summary.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
  --log-file /tmp/r02u_dram.csv python bench.py --steps 1 --warmup 2 --only-resident > gpurun_out/r02u_ncu_dram.log 2>&1
python tools/ncu_dram_total.py /tmp/r02u_dram.csv 'gw_|tc_mix_kernel|tc_dP_kernel|bn_finalize|bn_bwd_finalize|tc_support_images' 3 "G2 GWNet stack forward + backward" > gpurun_out/r02u_ncu_gw_stack_total.txt 2>&1
python tools/ncu_dram_total.py /tmp/r02u_dram.csv 'tc_attn|tc_linear|tc_embed|tc_ffn|tc_gram|gram_normalize' 3 "TSFormer encoder (bf16 tcgen05 path) + Gram" > gpurun_out/r02u_ncu_encoder_total.txt 2>&1
python tools/ncu_dram_total.py /tmp/r02u_dram.csv 'trunk_|fc_' 3 "D1 trunk forward + backward" > gpurun_out/r02u_ncu_trunk_total.txt 2>&1
python tools/ncu_dram_total.py /tmp/r02u_dram.csv '.' 3 "whole step" > gpurun_out/r02u_ncu_step_total.txt 2>&1
head -4 gpurun_out/r02u_ncu_gw_stack_total.txt gpurun_out/r02u_ncu_encoder_total.txt gpurun_out/r02u_ncu_trunk_total.txt gpurun_out/r02u_ncu_step_total.txt
timeout 600 ncu --set full --clock-control none -k regex:'tc_attn_kernel' -s 4 -c 1 -o /tmp/r02u_attn -f \
  python bench.py --steps 1 --warmup 2 --only-resident > gpurun_out/r02u_ncu_attn.log 2>&1
python tools/ncu_summary.py /tmp/r02u_attn.ncu-rep > gpurun_out/r02u_ncu_attn.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'trunk_conv2_tc|gw_fused_fwd_kernel|tc_dP_kernel|tc_ffn|trunk_conv1_bwd' -s 12 -c 12 -o /tmp/r02u_new -f \
  python bench.py --steps 1 --warmup 1 --only-resident > gpurun_out/r02u_ncu_new.log 2>&1
python tools/ncu_summary.py /tmp/r02u_new.ncu-rep > gpurun_out/r02u_ncu_new_full.txt 2>&1
python tools/ncu_table.py gpurun_out/r02u_ncu_new_full.txt > gpurun_out/r02u_ncu_new_table.txt 2>&1
cat gpurun_out/r02u_ncu_new_table.txt
