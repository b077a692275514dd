for lib in base dyn0 dyn5 base dyn0 dyn5; do
WH_B200_LIB=$PWD/rllib_warehouse_b200/lib/ab/$lib.so python bench.py --steps 50 --warmup 10 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib', 'value %.4e frac %.4f' % (d['value'], d['roofline']['frac']), 'e2e %.4e ms %.4f' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'alt %.4e ms %.4f' % (d['e2e_alt']['value'], d['e2e_alt']['ms_per_step']))"
done
