#!/bin/bash
# usage: tools/ab_pf.sh "<WH_B200_PF values>" "<variant:envs:policy ...>"   prefetch-ahead distance sweep (default library)
for c in $2; do IFS=: read v n pol <<< "$c"
  for pf in $1; do
    WH_B200_PF=$pf python bench.py --variant $v --envs $n --policy $pol --steps 300 --warmup 30 \
      --no-e2e --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pf=$pf $v $n $pol', '%.4e' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'iso %.4f' % d['roofline']['frac_isolated'])
" | tee -a gpurun_out/ab_results.txt
  done
done
