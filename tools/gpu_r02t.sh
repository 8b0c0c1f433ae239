#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/attn_ab.py ab_tmp/libA.so > gpurun_out/r02t_attn_ab.txt 2>&1
cat gpurun_out/r02t_attn_ab.txt | tail -5
STEP_B200_ATTN_KHALF=0 STEP_B200_ATTN_STAGGER=0 timeout 300 python tools/attn_ab.py ab_tmp/libA.so > gpurun_out/r02t_attn_ab_noswitch.txt 2>&1
cat gpurun_out/r02t_attn_ab_noswitch.txt | tail -5
