#!/bin/bash
mkdir -p gpurun_out
bash tools/attn_matrix.sh > gpurun_out/r02p_attn_matrix.txt 2>&1
grep -E "khalf|kernel time|group" gpurun_out/r02p_attn_matrix.txt
