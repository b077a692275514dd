#!/bin/bash
# e2e (host-buffer C ABI) vs transport mode: tools/e2e_modes.sh "8 16 0 -8 -16"   (see wh_env_create: n_chunks)
for ch in $1; do
python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-extras --e2e-steps 200 --e2e-chunks=$ch 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunks=$ch', 'e2e %.4e' % d['e2e']['value'], 'ms %.4f' % d['e2e']['ms_per_step'], '| alt %.4e' % d['e2e_alt']['value'], 'ms %.4f' % d['e2e_alt']['ms_per_step'], '| host_obs %.4e' % d['e2e_host_obs']['value'])
"; done
