#!/bin/bash
# usage: tools/ab_run.sh "<variants: small medium large>" name1 name2 ...   (on the GPU box)
# benches every rllib_warehouse_b200/lib/ab/<name>.so on the given variants, kernel-only numbers
variants="$1"; shift
mkdir -p gpurun_out
for name in "$@"; do
  for v in $variants; do
    WH_B200_LIB=$PWD/rllib_warehouse_b200/lib/ab/$name.so python bench.py --variant $v --steps 200 --warmup 20 \
      --no-e2e --no-cpu-baseline --no-extras 2>gpurun_out/ab_$name.$v.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name', '$v', '%.4e' % d['value'], 'frac %.4f' % d['roofline']['frac'], 'kernel_ms %.4f' % d['roofline']['kernel_ms_avg'])
" | tee -a gpurun_out/ab_results.txt
  done
done
