#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r02z_bench_2gpu.json 2> gpurun_out/r02z_bench_2gpu.err
python - <<'P'
import json
d=json.load(open("gpurun_out/r02z_bench_2gpu.json")); print(d["value"], d["ms_per_step"], d["n_gpus"])
for s in d.get("secondary",[]): print("   ", s["workload"], s["mode"], s["value"], s["ms_per_step"])
P
