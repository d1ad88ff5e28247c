set -x
mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests/test_reference_on_engine.py tests/test_lazy_engine.py -m gpu -x -q -s > gpurun_out/r2/t4_reference_unchanged.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t4_reference_unchanged.log
