#!/bin/bash
# on the GPU box: wh_multi_step kernels across tuning builds (tools/ab_build.py) and batch sizes
# usage: tools/multi_sweep.sh  -> gpurun_out/multi_sweep.jsonl
mkdir -p gpurun_out
out=gpurun_out/multi_sweep.jsonl; : > $out
run() { lib=$1; shift; echo "{\"lib\": \"$lib\"}" >> $out; WH_B200_LIB=$PWD/rllib_warehouse_b200/lib/ab/$lib.so python tools/multi_small.py "$@" >> $out 2>&1; }
for lib in base d1 d2; do run $lib small 4096 200 ws1 ws2; done
for lib in base mb4 mb3 mb2; do
  run $lib medium 65536 50 throughput
  run $lib small 65536 50 throughput
  run $lib small 262144 20 throughput
done
run base medium 65536 50 low_occupancy
run base small 8192 100 throughput low_occupancy ws1 ws2
run base small 32768 100 throughput low_occupancy ws2
run base medium 8192 100 throughput low_occupancy ws1 ws2
cat $out
