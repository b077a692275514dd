#!/bin/bash
# usage: tools/ab_pool.sh "<lib names under lib/ab>" "<variant:envs:policy:pool ...>"   like ab_policy.sh, with the size of the random-action pool
libs="$1"; cases="$2"
mkdir -p gpurun_out
for c in $cases; do IFS=: read v n pol pool <<< "$c"
  for lib in $libs; do
    WH_B200_LIB=$PWD/rllib_warehouse_b200/lib/ab/$lib.so python bench.py --variant $v --envs $n --policy $pol --action-pool $pool --steps 300 --warmup 30 \
      --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib $v $n $pol pool=$pool', '%.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'iso %.4f' % d['roofline']['frac_isolated'])
" | tee -a gpurun_out/ab_results.txt
  done
done
