# EvalMod: the quotient's constant term joins the quotient before the product (CKKS_CHEB_C0_FOLD): A/B, GPU tests, default
# bench, phases and the launch list of the graph replay on the resulting tree
O=gpurun_out/r2j; mkdir -p $O
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["s_per_round_per_pair"], d["roofline"]["frac"], d["bytes_exact_vs_fips197"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["rotations_per_s_n16"]["level_14_batch_4"])'
for rep in 1 2; do for f in 0 1; do
  echo "== cheb_c0_fold $f"
  CKKS_CHEB_C0_FOLD=$f python bench.py --no-cpu --no-dec --steps 3 --warmup 3 > $O/ab_c0_$f.json 2> $O/ab_c0_$f.err
  python -c "$pick" < $O/ab_c0_$f.json
done; done
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
python -m pytest tests -m gpu -q --durations=5 > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
python -c "$pick" < $O/bench.json
python tools/phase_breakdown.py > $O/phases.json 2> $O/phases.err
python tools/boot_precision.py > $O/boot_precision.json 2> $O/boot_precision.err
BENCH_NCU_ROUND=1 timeout 600 ncu --graph-profiling node --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_graph_round.csv python bench.py --no-cpu --no-dec --steps 1 --warmup 1 > $O/ncu_round.log 2>&1; gzip -f $O/launches_graph_round.csv
