#!/bin/bash
# final state: full GPU test suite, default bench line, launch list, G2 DRAM totals, ncu of the fused GWNet layer kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02y_gputest.log 2>&1
tail -6 gpurun_out/r02y_gputest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02y_bench_1gpu.json 2> gpurun_out/r02y_bench_1gpu.err
tail -c 200 gpurun_out/r02y_bench_1gpu.json; tail -2 gpurun_out/r02y_bench_1gpu.err
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
  --log-file /tmp/r02y_dram.csv python bench.py --steps 1 --warmup 2 --only-resident > gpurun_out/r02y_ncu_dram.log 2>&1
python tools/ncu_dram_total.py /tmp/r02y_dram.csv 'gw_|tc_mix_kernel|tc_dP_kernel|bn_finalize|bn_bwd_finalize|tc_support_images' 3 "G2 GWNet stack forward + backward" > gpurun_out/r02y_ncu_gw_stack_total.txt 2>&1
python tools/ncu_dram_total.py /tmp/r02y_dram.csv '.' 3 "whole step" > gpurun_out/r02y_ncu_step_total.txt 2>&1
grep -v "^==" /tmp/r02y_dram.csv | python -c "
import sys,csv
r=csv.reader(sys.stdin); hdr=next(r); ki,mi,vi=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value')
rows=[(x[ki],x[vi]) for x in r if len(x)>vi and x[mi]=='gpu__time_duration.sum']
w=csv.writer(open('gpurun_out/r02y_launches.csv','w')); w.writerow(['Kernel Name','Metric Value'])
for k,v in rows: w.writerow([k,v])
"
python tools/launch_summary.py gpurun_out/r02y_launches.csv 3 70 > gpurun_out/r02y_launches_summary.txt 2>&1
head -14 gpurun_out/r02y_launches_summary.txt
timeout 300 ncu --set full --clock-control none -k regex:'gw_fused_fwd_kernel' -s 7 -c 2 -o /tmp/r02y_gwf -f \
  python bench.py --steps 1 --warmup 1 --only-resident > gpurun_out/r02y_ncu_gwf.log 2>&1
python tools/ncu_summary.py /tmp/r02y_gwf.ncu-rep > gpurun_out/r02y_ncu_gwf.txt 2>&1
grep -n "gpu__time_duration\|issue_active\|tensor_cycles\|fma_cycles" gpurun_out/r02y_ncu_gwf.txt | head -8
