set -x
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests/test_engine_parity.py -m gpu -x -q > gpurun_out/r2/t6_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t6_parity.log
for P in 6 8; do
  timeout 900 python bench.py --pairs $P --steps 2 --warmup 1 --no-cpu --no-dec > gpurun_out/r2/bench6_p$P.json 2> gpurun_out/r2/bench6_p$P.err; echo "rc=$?" >> gpurun_out/r2/bench6_p$P.err
  timeout 900 python bench.py --pairs $P --steps 2 --warmup 1 --no-cpu --no-dec --serial-graphs > gpurun_out/r2/bench6_p${P}_serial.json 2> gpurun_out/r2/bench6_p${P}_serial.err; echo "rc=$?" >> gpurun_out/r2/bench6_p${P}_serial.err
done
