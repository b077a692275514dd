#!/bin/bash
# usage: tools/ab_policy.sh "<lib names under lib/ab>" "<variant:envs:policy ...>"   one kernel-only bench line per combination
libs="$1"; cases="$2"
mkdir -p gpurun_out
for c in $cases; do IFS=: read v n pol <<< "$c"
  for lib in $libs; do
    WH_B200_LIB=$PWD/rllib_warehouse_b200/lib/ab/$lib.so python bench.py --variant $v --envs $n --policy $pol --steps 300 --warmup 30 \
      --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib $v $n $pol', '%.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'iso %.4f' % d['roofline']['frac_isolated'])
" | tee -a gpurun_out/ab_results.txt
  done
done
