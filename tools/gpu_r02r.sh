#!/bin/bash
# round-2 evidence run at HEAD (1 GPU): tests, full bench line, reference arm, launch list, ncu captures -> text summaries
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02r_gputest.log 2>&1
tail -3 gpurun_out/r02r_gputest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02r_bench_1gpu.json 2> gpurun_out/r02r_bench_1gpu.err
tail -c 300 gpurun_out/r02r_bench_1gpu.json; tail -2 gpurun_out/r02r_bench_1gpu.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02r_bench_reference.json 2> gpurun_out/r02r_bench_reference.err
tail -c 400 gpurun_out/r02r_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r_launches.csv \
  python bench.py --steps 1 --warmup 3 --only-resident > gpurun_out/r02r_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/r02r_launches.csv 4 60 > gpurun_out/r02r_launches_summary.txt 2>&1
head -8 gpurun_out/r02r_launches_summary.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
  --log-file /tmp/r02r_dram.csv python bench.py --steps 1 --warmup 1 --only-resident > gpurun_out/r02r_ncu_dram.log 2>&1
python tools/ncu_dram_total.py /tmp/r02r_dram.csv 'gw_|tc_mix_kernel|tc_dP_kernel|bn_finalize|bn_bwd_finalize|tc_support_images' 2 "G2 GWNet stack forward + backward" > gpurun_out/r02r_ncu_gw_stack_total.txt 2>&1
python tools/ncu_dram_total.py /tmp/r02r_dram.csv 'tc_attn|tc_linear|tc_embed|tc_ffn|tc_gram|gram_normalize' 2 "TSFormer encoder (bf16 tcgen05 path) + Gram" > gpurun_out/r02r_ncu_encoder_total.txt 2>&1
python tools/ncu_dram_total.py /tmp/r02r_dram.csv '.' 2 "whole step" > gpurun_out/r02r_ncu_step_total.txt 2>&1
head -5 gpurun_out/r02r_ncu_gw_stack_total.txt gpurun_out/r02r_ncu_encoder_total.txt gpurun_out/r02r_ncu_step_total.txt
timeout 600 ncu --set full --clock-control none -k regex:'tc_attn_kernel' -s 4 -c 1 -o /tmp/r02r_attn -f \
  python bench.py --steps 1 --warmup 2 --only-resident > gpurun_out/r02r_ncu_attn.log 2>&1
python tools/ncu_summary.py /tmp/r02r_attn.ncu-rep > gpurun_out/r02r_ncu_attn.txt 2>&1
ncu -i /tmp/r02r_attn.ncu-rep --page details > gpurun_out/r02r_ncu_attn_details.txt 2>&1
timeout 900 ncu --set full --clock-control none \
  -k regex:'gw_|tc_mix_kernel|tc_dP_kernel|trunk_|fc_fwd|fc_dx|fc_dw|fc_bn|tc_gemm_kernel|tc_linear_kernel|tc_embed|tc_gram|topk|edge_logits|adam_update|grad_sumsq' \
  -s 150 -c 160 -o /tmp/r02r_rest -f python bench.py --steps 1 --warmup 1 --only-resident > gpurun_out/r02r_ncu_rest.log 2>&1
python tools/ncu_summary.py /tmp/r02r_rest.ncu-rep > gpurun_out/r02r_ncu_rest_full.txt 2>&1
python tools/ncu_table.py gpurun_out/r02r_ncu_rest_full.txt > gpurun_out/r02r_ncu_rest_table.txt 2>&1
cat gpurun_out/r02r_ncu_rest_table.txt | head -50
ls -la gpurun_out | tail -14
