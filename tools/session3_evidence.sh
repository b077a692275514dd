#!/bin/bash
# Round-2 session-3 evidence on one GPU: the GPU test suite, smoke, the driver's bench line, Medium / Small lines and
# ncu --set full captures of the Medium / Small step kernels on the build with the by-array keep policy and the
# action burst prefetch. Outputs: gpurun_out/r02s3_*
tag=r02s3; out=gpurun_out; mkdir -p $out
(timeout 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > $out/${tag}_pytest.log
(timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > $out/${tag}_smoke.log
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
: > $out/${tag}_bench_medium_small.jsonl
for v in medium small; do
  timeout 60 python bench.py --variant $v --steps 300 --warmup 30 --no-e2e --no-cpu-baseline --no-extras >> $out/${tag}_bench_medium_small.jsonl 2>/dev/null
done
timeout 60 python bench.py --variant medium --envs 65536 --policy greedy_fused --steps 300 --warmup 30 --no-e2e --no-cpu-baseline --no-extras >> $out/${tag}_bench_medium_small.jsonl 2>/dev/null
CMD="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extras"
for v in medium small; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_step -s 8 -c 2 -f -o $out/${tag}_prof_$v \
      $CMD --variant $v > $out/${tag}_ncu_$v.log 2>&1
done
summarise() {
  local rep=$out/${tag}_prof_$1.ncu-rep
  [ -f $rep ] || return
  timeout 120 python tools/ncu_summary.py $rep > $out/${tag}_ncu_$1.txt 2>&1
  timeout 120 python tools/ncu_by_line.py $rep "$2" > $out/${tag}_hotspots_$1.txt 2>&1
  rm -f $rep
}
summarise medium k_stepILi9ELi9ELb0ELb0ELb1ELi2
summarise small k_stepILi4ELi4ELb0ELb0ELb1ELi0
cat $out/${tag}_pytest.log $out/${tag}_smoke.log
python - <<'P'
import json
for l in open("gpurun_out/r02s3_bench_medium_small.jsonl"):
    d = json.loads(l); print(d["config"]["workload"], d["config"]["policy"], "%.4e" % d["value"], "frac %.4f" % d["roofline"]["frac"])
d = json.load(open("gpurun_out/r02s3_bench_1gpu.json")); print("large", "%.4e" % d["value"], d["roofline"]["frac"], "e2e %.4e" % d["e2e"]["value"])
P
head -12 $out/${tag}_ncu_small.txt
