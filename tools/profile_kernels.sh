#!/bin/bash
# usage: tools/profile_kernels.sh <tag> "<variants>"   ncu --set full captures of k_step for the given variants
# (each only after the same command has exited 0 without ncu); reports land in gpurun_out/<tag>_prof_<variant>.ncu-rep
tag=${1:-rXX}; variants=${2:-"large medium small"}
out=gpurun_out; mkdir -p $out
CMD="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extras"
for v in $variants; do
  $CMD --variant $v > $out/${tag}_bench_$v.json 2>/dev/null || exit 2
  ncu --set full --clock-control none --import-source on -k regex:k_step -s 8 -c 2 -f -o $out/${tag}_prof_$v \
      $CMD --variant $v > $out/${tag}_ncu_$v.log 2>&1
done
ls -la $out | grep ${tag}_
