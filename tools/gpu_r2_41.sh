# level alignment folded into the rescale (CKKS_ALIGN_FUSE): GPU tests, A/B, default bench
O=gpurun_out/r2i; mkdir -p $O
python -m pytest tests -m gpu -q --durations=5 > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["s_per_round_per_pair"], d["roofline"]["frac"], d["bytes_exact_vs_fips197"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["rotations_per_s_n16"]["level_14_batch_4"])'
for rep in 1 2; do for f in 0 1; do
  echo "== align_fuse $f"
  CKKS_ALIGN_FUSE=$f python bench.py --no-cpu --no-dec --steps 3 --warmup 3 > $O/ab_al_$f.json 2> $O/ab_al_$f.err
  python -c "$pick" < $O/ab_al_$f.json
done; done
python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
python -c "$pick" < $O/bench.json
