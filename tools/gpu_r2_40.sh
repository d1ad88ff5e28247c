# 2 a b of the Chebyshev recurrences folded into the division's scalar (Engine::mul(a, b, factor)): GPU tests, A/B, default bench
O=gpurun_out/r2h; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
python -m pytest tests -m gpu -q --durations=5 > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["s_per_round_per_pair"], d["roofline"]["frac"], d["bytes_exact_vs_fips197"], d["gpu_launches"], d["clocks"]["sm_mhz"])'
for rep in 1 2; do for f in 0 1; do
  echo "== mul_factor_fuse $f"
  CKKS_MUL_FACTOR_FUSE=$f python bench.py --no-cpu --no-dec --steps 3 --warmup 3 > $O/ab_mf_$f.json 2> $O/ab_mf_$f.err
  python -c "$pick" < $O/ab_mf_$f.json
done; done
python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
python -c "$pick" < $O/bench.json
python tools/phase_breakdown.py > $O/phases.json 2> $O/phases.err
