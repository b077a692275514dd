#!/bin/bash
# on the GPU box: ncu capture of the warp-specialised wh_multi_step kernel (4 096 Small envs, greedy) + size thresholds
mkdir -p gpurun_out
python tools/multi_small.py small 4096 200 ws1 > gpurun_out/r02t_ws1.json 2>&1 || exit 2
ncu --set full --clock-control none --import-source on -k regex:k_multi_ws -s 1 -c 1 -f -o gpurun_out/r02t_prof_ws1_small4096 python tools/multi_small.py small 4096 200 ws1 > gpurun_out/r02t_ncu.log 2>&1
out=gpurun_out/r02t_thresholds.jsonl; : > $out
python tools/multi_small.py small 32768 100 throughput low_occupancy >> $out 2>&1
python tools/multi_small.py small 65536 50 throughput low_occupancy >> $out 2>&1
python tools/multi_small.py medium 16384 100 throughput low_occupancy >> $out 2>&1
python tools/multi_small.py medium 32768 50 throughput low_occupancy >> $out 2>&1
python tools/multi_small.py large 8192 50 throughput low_occupancy ws1 >> $out 2>&1
python tools/multi_small.py large 65536 20 throughput >> $out 2>&1
cat gpurun_out/r02t_ws1.json $out
