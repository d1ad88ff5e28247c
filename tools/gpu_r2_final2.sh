# closing evidence of round 2 (one B200): GPU tests, A/B of the level-alignment epilogue, default bench, reference arm,
# phases, kernel sizes, launch list of the graph replay of a middle round
set -x
O=gpurun_out/r2final2; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
python -m pytest tests -m gpu -q --durations=8 > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["s_per_round_per_pair"], d["roofline"]["frac"], d["bytes_exact_vs_fips197"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["rotations_per_s_n16"]["level_14_batch_4"])'
for f in 0 1; do
  echo "== align_fuse $f"
  CKKS_ALIGN_FUSE=$f python bench.py --no-cpu --no-dec --steps 3 --warmup 3 > $O/ab_al_$f.json 2> $O/ab_al_$f.err
  python -c "$pick" < $O/ab_al_$f.json
done
python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
python -c "$pick" < $O/bench.json
python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python tools/phase_breakdown.py > $O/phases.json 2> $O/phases.err
python tools/batch_scaling.py > $O/batch_scaling.json 2> $O/batch_scaling.err
python tools/ntt_sizes.py > $O/ntt_sizes.json 2> $O/ntt_sizes.err
BENCH_NCU_ROUND=1 timeout 900 ncu --graph-profiling node --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_graph_round.csv python bench.py --no-cpu --no-dec --steps 1 --warmup 1 > $O/ncu_round.log 2>&1; gzip -f $O/launches_graph_round.csv
ls -la $O
