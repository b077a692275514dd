#!/bin/bash
# e2e transport on N GPUs of one box: tools/e2e_multi.sh N "<n_chunks list>" [extra bench flags]
n=$1; chunks="$2"; shift 2
nvidia-smi topo -m 2>/dev/null | head -14; numactl -H 2>/dev/null | head -4; ls /sys/devices/system/node/ | head
for ch in $chunks; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
    bench.py --gpus $n --steps 20 --warmup 5 --e2e-chunks=$ch "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=$n chunks=$ch $*', 'value %.3e' % d['value'], 'e2e %.3e ms %.3f' % (d['e2e']['value'], d['e2e']['ms_per_step']), '| alt %.3e ms %.3f' % (d['e2e_alt']['value'], d['e2e_alt']['ms_per_step']), 'numa', d.get('numa_node_rank0'))
"
done
