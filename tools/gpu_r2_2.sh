set -x
mkdir -p gpurun_out/r2
for P in 1 2 4; do
  timeout 900 python bench.py --pairs $P --steps 2 --warmup 1 --no-cpu --no-dec > gpurun_out/r2/bench_p$P.json 2> gpurun_out/r2/bench_p$P.err; echo "rc=$?" >> gpurun_out/r2/bench_p$P.err
done
timeout 900 python bench.py --pairs 4 --steps 2 --warmup 1 --no-cpu --no-dec --serial-graphs > gpurun_out/r2/bench_p4_serial.json 2> gpurun_out/r2/bench_p4_serial.err; echo "rc=$?" >> gpurun_out/r2/bench_p4_serial.err
timeout 1500 python bench.py --pairs 4 --steps 3 --warmup 3 > gpurun_out/r2/bench_full.json 2> gpurun_out/r2/bench_full.err; echo "rc=$?" >> gpurun_out/r2/bench_full.err
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 --config0 > gpurun_out/r2/bench_ref.json 2> gpurun_out/r2/bench_ref.err; echo "rc=$?" >> gpurun_out/r2/bench_ref.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
