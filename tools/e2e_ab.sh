for pdl in 0 1 0 1; do for ch in 4 8; do
WH_B200_PDL=$pdl python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-extras --e2e-steps 200 --e2e-chunks $ch 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pdl=$pdl chunks=$ch', 'e2e %.4e' % d['e2e']['value'], 'ms %.4f' % d['e2e']['ms_per_step'], d['e2e']['api'][:28], '| alt %.4e' % d['e2e_alt']['value'])
"; done; done
