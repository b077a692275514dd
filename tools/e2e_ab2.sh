#!/bin/bash
# usage: tools/e2e_ab2.sh "<lib names under lib/ab>" "<chunk counts>"
for rep in 1 2; do for lib in $1; do for ch in $2; do
WH_B200_LIB=$PWD/rllib_warehouse_b200/lib/ab/$lib.so python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-extras --e2e-steps 300 --e2e-chunks $ch 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib chunks=$ch', 'e2e %.4e' % d['e2e']['value'], 'ms %.4f' % d['e2e']['ms_per_step'], d['e2e']['api'][:28], '| alt %.4e' % d['e2e_alt']['value'], 'ms %.4f' % d['e2e_alt']['ms_per_step'])
"; done; done; done
