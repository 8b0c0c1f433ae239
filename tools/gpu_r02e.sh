#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/debug_gemm.py > gpurun_out/r02e_gemm_debug.log 2>&1
grep -A8 "GwEpilogue autograd" gpurun_out/r02e_gemm_debug.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "prologue_and_epilogue and 2-207" > gpurun_out/r02e_memcheck.log 2>&1
tail -30 gpurun_out/r02e_memcheck.log
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02e_pytest.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r02e_pytest.log | tail -25
