set -x
mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2/t9_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t9_all.log
