# final evidence of round 2 (one B200): GPU tests, default bench, reference arm, graph-round launch list, phases, kernel sizes,
# ncu --set full of the FP64 multiply-accumulate kernels
set -x
O=gpurun_out/r2final; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
python -m pytest tests -m gpu -q --durations=8 > $O/gpu_tests.log 2>&1; echo "rc=$?" >> $O/gpu_tests.log
python bench.py > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
python bench.py --impl reference > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python tools/phase_breakdown.py > $O/phases.json 2> $O/phases.err
python tools/batch_scaling.py > $O/batch_scaling.json 2> $O/batch_scaling.err
python tools/ntt_sizes.py > $O/ntt_sizes.json 2> $O/ntt_sizes.err
python tools/lut_once.py > $O/lut_once.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lut2|k_lincomb|k_diag_mac_rows" -c 12 -o $O/ncu_lut python tools/lut_once.py > $O/ncu_lut.log 2>&1
BENCH_NCU_ROUND=1 timeout 1200 ncu --graph-profiling node --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_graph_round.csv python bench.py --no-cpu --no-dec --steps 1 --warmup 1 > $O/ncu_round.log 2>&1; gzip -f $O/launches_graph_round.csv
ls -la $O
