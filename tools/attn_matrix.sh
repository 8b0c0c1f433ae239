#!/bin/bash
# A/B matrix of the attention kernel's P=168 switches (kernel time from tools/kbench-style CUDA events inside attn_trace.py)
for kh in 0 6; do for sg in 0 1; do for mn in 1; do
  echo "khalf=$kh stagger=$sg manual=$mn"
  STEP_B200_ATTN_KHALF=$kh STEP_B200_ATTN_STAGGER=$sg STEP_B200_ATTN_MANUAL=$mn timeout 200 python tools/attn_trace.py 6624 32 2>&1 | tail -3
done; done; done
