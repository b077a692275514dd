#!/bin/bash
# Round-end evidence run on the GPU box (one GPU): bench (no profiler), reference arm, Medium / Small lines, then ncu
# captures of the SAME bench command lines. usage: tools/profile_round.sh <tag>   (outputs under gpurun_out/<tag>_*)
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err || exit 1
python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
: > $out/${tag}_bench_medium_small.jsonl
for v in medium small; do
  python bench.py --variant $v --steps 300 --warmup 30 --no-e2e --no-cpu-baseline --no-extras >> $out/${tag}_bench_medium_small.jsonl 2>/dev/null
done
python bench.py --variant medium --envs 65536 --policy greedy_fused --steps 300 --warmup 30 --no-e2e --no-cpu-baseline --no-extras >> $out/${tag}_bench_medium_small.jsonl 2>/dev/null
python tools/diag_split.py > $out/${tag}_diag_split.jsonl 2>/dev/null
CMD="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extras"
for v in large medium small; do
  $CMD --variant $v > /dev/null 2>&1 || exit 2      # the command exits 0 without ncu first
  ncu --set full --clock-control none --import-source on -k regex:k_step -s 8 -c 2 -f -o $out/${tag}_prof_$v \
      $CMD --variant $v > $out/${tag}_ncu_$v.log 2>&1
done
# flat-observation step kernel (Medium): the 4th block of 220 k_step launches of diag_split.py
ncu --set full --clock-control none --import-source on -k k_step -s 700 -c 1 -f -o $out/${tag}_prof_medium_flat \
    python tools/diag_split.py medium > $out/${tag}_ncu_medium_flat.log 2>&1
LCMD="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extras --e2e-steps 10"
$LCMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file $out/${tag}_launches.csv $LCMD > $out/${tag}_launches.log 2>&1
ls -la $out | grep ${tag}_
