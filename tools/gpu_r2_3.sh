set -x
mkdir -p gpurun_out/r2
BENCH_NCU_ROUND=1 timeout 1700 ncu --graph-profiling node --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_round_p4.csv python bench.py --pairs 4 --steps 1 --warmup 1 --no-cpu --no-dec > gpurun_out/r2/ncu_round_p4.log 2>&1; echo "rc=$?" >> gpurun_out/r2/ncu_round_p4.log
gzip -f gpurun_out/r2/launches_round_p4.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_ks_b8.csv python tools/ks_batch_once.py 8 > gpurun_out/r2/ncu_ks_b8.log 2>&1; echo "rc=$?" >> gpurun_out/r2/ncu_ks_b8.log
timeout 2400 python -m pytest tests/test_reference_on_engine.py -m gpu -x -q -s > gpurun_out/r2/t3_reference_unchanged.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t3_reference_unchanged.log
