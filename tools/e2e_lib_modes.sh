#!/bin/bash
# usage: tools/e2e_lib_modes.sh "<lib names under lib/ab or 'default'>" "<n_chunks list>"
for lib in $1; do for ch in $2; do
L=""; [ "$lib" != default ] && L="WH_B200_LIB=$PWD/rllib_warehouse_b200/lib/ab/$lib.so"
env $L python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-extras --e2e-steps 200 --e2e-chunks=$ch 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib chunks=$ch', 'kernel ms %.4f' % d['ms_per_step'], 'e2e %.4e' % d['e2e']['value'], 'ms %.4f' % d['e2e']['ms_per_step'], '| alt %.4e' % d['e2e_alt']['value'], 'ms %.4f' % d['e2e_alt']['ms_per_step'])
"; done; done
