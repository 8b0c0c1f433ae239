#!/bin/bash
# round-2 multi-GPU evidence (2 GPUs): multi-GPU tests, batch-parallel bench (+ secondary configs), PEMS07 node-parallel vs 1 GPU
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/r02v_gputest_multi.log 2>&1
tail -4 gpurun_out/r02v_gputest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02v_bench_2gpu.json 2> gpurun_out/r02v_bench_2gpu.err
tail -c 1200 gpurun_out/r02v_bench_2gpu.json
grep -E "NCCL INFO.*(nranks|NVLS|Connected all)" gpurun_out/r02v_bench_2gpu.err | head -6
timeout 600 python bench.py --workload PEMS07 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02v_bench_pems07_1gpu.json 2>/dev/null
timeout 600 python bench.py --workload PEMS-BAY --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary > gpurun_out/r02v_bench_pemsbay_1gpu.json 2>/dev/null
python - <<'P'
import json
for f in ["r02v_bench_2gpu","r02v_bench_pems07_1gpu","r02v_bench_pemsbay_1gpu"]:
    d=json.load(open("gpurun_out/"+f+".json")); print(f, d["value"], d["ms_per_step"], d["n_gpus"], d["config"].get("workload"))
    for s in d.get("secondary",[]): print("   ", s["workload"], s["mode"], s["value"], s["ms_per_step"])
P
