#!/bin/bash
# e2e (host-buffer C ABI) vs number of pipeline chunks: tools/e2e_ab.sh "8 12 16 24 32"
for ch in $1; do
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-extras --e2e-steps 200 --e2e-chunks $ch 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunks=$ch', 'e2e %.4e' % d['e2e']['value'], 'ms %.4f' % d['e2e']['ms_per_step'], d['e2e']['api'][:28], '| alt %.4e' % d['e2e_alt']['value'])
"; done
