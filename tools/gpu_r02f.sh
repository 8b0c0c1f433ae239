#!/bin/bash
# 2-GPU call: whole GPU suite incl. the multi-GPU tests, 2-GPU bench lines (batch-parallel headline + secondary incl. node mode)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02f_pytest.log 2>&1
grep -E "passed|failed|FAILED|skipped" gpurun_out/r02f_pytest.log | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_2gpu.json 2> gpurun_out/r02f_bench_2gpu.err
tail -c 1500 gpurun_out/r02f_bench_2gpu.json; grep -E "NCCL INFO.*(nranks|NVLS|Connected|Channel 00)" gpurun_out/r02f_bench_2gpu.err | head -8; tail -3 gpurun_out/r02f_bench_2gpu.err
STEP_B200_OVERLAP_REDUCE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02f_bench_2gpu_sync.json 2> /dev/null
tail -c 300 gpurun_out/r02f_bench_2gpu_sync.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02f_bench_1gpu.json 2>/dev/null
timeout 600 python bench.py --workload PEMS07 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02f_bench_pems07_1gpu.json 2>/dev/null
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --workload PEMS07 --mode node --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02f_bench_pems07_node2.json 2>/dev/null
for f in gpurun_out/r02f_bench_1gpu.json gpurun_out/r02f_bench_pems07_1gpu.json gpurun_out/r02f_bench_pems07_node2.json; do python -c "import json,sys; d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['n_gpus'], d['config']['parallelism'])"; done
