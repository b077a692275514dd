#!/bin/bash
# The driver's 1 -> 8 GPU scaling protocol on ONE box: tools/scale_run.sh <tag> "1 2 4 8"   (outputs gpurun_out/<tag>_scale_N.json)
tag=${1:-rXX}; ns=${2:-"1 2 4 8"}
mkdir -p gpurun_out
for n in $ns; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/${tag}_scale_$n.json 2> gpurun_out/${tag}_scale_$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_scale_$n.json 2> gpurun_out/${tag}_scale_$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_scale_$n.json").read().strip().splitlines()[-1])
    print("N=$n value %.4e ms/step %.4f frac %.4f collective_ms %.4f e2e %.4e (%s)" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["collective_ms"], d["e2e"]["value"], d["collective_api"][:40]))
except Exception as e:
    print("N=$n FAILED", e)
PY
done
