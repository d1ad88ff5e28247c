set -x
O=gpurun_out/r2; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 > $O/t33_all.log 2>&1; echo "rc=$?" >> $O/t33_all.log
timeout 900 python bench.py > $O/bench33.json 2> $O/bench33.err; echo "rc=$?" >> $O/bench33.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench33_ref.json 2> $O/bench33_ref.err; echo "rc=$?" >> $O/bench33_ref.err
timeout 600 python tools/phase_breakdown.py > $O/phases33.json 2> $O/phases33.err
timeout 600 python tools/batch_scaling.py > $O/bs33.json 2> $O/bs33.err
BENCH_NCU_ROUND=1 timeout 1500 ncu --graph-profiling node --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches33_round_p4.csv python bench.py --pairs 4 --no-cpu --no-dec --steps 1 --warmup 1 > $O/ncu33.log 2>&1
gzip -f $O/launches33_round_p4.csv
