set -x
O=gpurun_out/r2; mkdir -p $O
BENCH_NCU_ROUND=1 timeout 1500 ncu --graph-profiling node --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches17_round_p4.csv python bench.py --pairs 4 --no-cpu --no-dec --steps 1 --warmup 1 > $O/ncu17.log 2>&1
gzip -f $O/launches17_round_p4.csv
