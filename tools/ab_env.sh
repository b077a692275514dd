#!/bin/bash
# usage: tools/ab_env.sh <lib name under lib/ab> "<variants>" "<envs list>" VAR=val ...   one bench line per combination
lib=$1; variants="$2"; envs="$3"; shift 3
for v in $variants; do for n in $envs; do
  env "$@" WH_B200_LIB=$PWD/rllib_warehouse_b200/lib/ab/$lib.so python bench.py --variant $v --envs $n --steps 300 --warmup 30 \
    --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib $* $v $n', '%.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'kernel_ms %.4f' % d['roofline']['kernel_ms_avg'])
" | tee -a gpurun_out/ab_results.txt
done; done
