#!/bin/bash
# Round-end evidence run on the GPU box (one GPU): bench (no profiler), reference arm, Medium / Small lines, then ncu
# captures of the SAME command lines. The .ncu-rep files are summarised ON THE BOX (ncu_summary.py, ncu_by_line.py)
# and deleted — six of them exceed gpurun's 64 MiB return limit. KEEP_REPS="large ..." keeps some.
# usage: tools/profile_round.sh <tag> [parts]   parts: bench ncu multi launches (default: all)
# outputs under gpurun_out/<tag>_*
tag=${1:-rXX}
parts=${2:-"bench ncu multi launches"}
out=gpurun_out
mkdir -p $out
has() { [[ " $parts " == *" $1 "* ]]; }
summarise() {   # <name> <kernel-substring>: text summary + per-line hotspots from $out/${tag}_prof_<name>.ncu-rep
  local rep=$out/${tag}_prof_$1.ncu-rep
  [ -f $rep ] || return
  python tools/ncu_summary.py $rep > $out/${tag}_ncu_$1.txt 2>&1
  python tools/ncu_by_line.py $rep "$2" > $out/${tag}_hotspots_$1.txt 2>&1
  [[ " $KEEP_REPS " == *" $1 "* ]] || rm -f $rep
}
if has bench; then
  python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err || exit 1
  python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
  : > $out/${tag}_bench_medium_small.jsonl
  for v in medium small; do
    python bench.py --variant $v --steps 300 --warmup 30 --no-e2e --no-cpu-baseline --no-extras >> $out/${tag}_bench_medium_small.jsonl 2>/dev/null
  done
  python bench.py --variant medium --envs 65536 --policy greedy_fused --steps 300 --warmup 30 --no-e2e --no-cpu-baseline --no-extras >> $out/${tag}_bench_medium_small.jsonl 2>/dev/null
  python tools/diag_split.py > $out/${tag}_diag_split.jsonl 2>/dev/null
fi
if has ncu; then
  CMD="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extras"
  for v in large medium small; do
    $CMD --variant $v > /dev/null 2>&1 || exit 2      # the command exits 0 without ncu first
    ncu --set full --clock-control none --import-source on -k regex:k_step -s 8 -c 2 -f -o $out/${tag}_prof_$v \
        $CMD --variant $v > $out/${tag}_ncu_$v.log 2>&1
  done
  summarise large k_stepILi16ELi16ELb0ELb0ELb1ELi0
  summarise medium k_stepILi9ELi9ELb0ELb0ELb1ELi2
  summarise small k_stepILi4ELi4ELb0ELb0ELb1ELi0
  # flat-observation step kernel (Medium): the 4th block of 220 k_step launches of diag_split.py
  ncu --set full --clock-control none --import-source on -k k_step -s 700 -c 1 -f -o $out/${tag}_prof_medium_flat \
      python tools/diag_split.py medium > $out/${tag}_ncu_medium_flat.log 2>&1
  summarise medium_flat k_stepILi9ELi9ELb0ELb1ELb1ELi2
fi
if has multi; then
  # wh_multi_step: launch-sized (4 096 Small envs -> k_multi_ws) and BASELINE configs[2]-sized (Medium 65 536 -> k_multi)
  python tools/multi_small.py small 4096 200 throughput low_occupancy ws1 ws2 auto > $out/${tag}_multi_step.jsonl 2>&1
  python tools/multi_small.py medium 65536 50 auto low_occupancy >> $out/${tag}_multi_step.jsonl 2>&1
  python tools/multi_small.py small 262144 20 auto >> $out/${tag}_multi_step.jsonl 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_multi -s 1 -c 1 -f -o $out/${tag}_prof_multi_ws_small4096 \
      python tools/multi_small.py small 4096 200 auto > $out/${tag}_ncu_multi_ws_small4096.log 2>&1
  summarise multi_ws_small4096 k_multi_wsILi4ELi4ELb1ELi2
  ncu --set full --clock-control none --import-source on -k regex:k_multi -s 1 -c 1 -f -o $out/${tag}_prof_multi_medium65536 \
      python tools/multi_small.py medium 65536 50 auto > $out/${tag}_ncu_multi_medium65536.log 2>&1
  summarise multi_medium65536 k_multiILi9ELi9ELb1ELb0
fi
if has launches; then
  LCMD="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extras --e2e-steps 10"
  $LCMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
      --log-file $out/${tag}_launches.csv $LCMD > $out/${tag}_launches.log 2>&1
fi
ls -la $out | grep ${tag}_
du -sh $out
