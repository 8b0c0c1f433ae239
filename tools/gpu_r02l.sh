#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/ffn_debug.py > gpurun_out/r02l_ffn_debug.txt 2>&1
tail -40 gpurun_out/r02l_ffn_debug.txt
