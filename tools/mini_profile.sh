out=gpurun_out; tag=r01f
CMD="python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-extras"
$CMD --variant large > $out/${tag}_bench_large_short.json 2>/dev/null || exit 2
ncu --set full --clock-control none --import-source on -k regex:k_step -s 8 -c 2 -f -o $out/${tag}_prof_large $CMD --variant large > $out/${tag}_ncu_large.log 2>&1
LCMD="python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-extras --e2e-steps 10"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches.csv $LCMD > $out/${tag}_launches.log 2>&1
ls $out | grep -c $tag
