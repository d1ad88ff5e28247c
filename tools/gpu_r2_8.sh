set -x
mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2/t8_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t8_all.log
timeout 900 python bench.py > gpurun_out/r2/bench8.json 2> gpurun_out/r2/bench8.err; echo "rc=$?" >> gpurun_out/r2/bench8.err
