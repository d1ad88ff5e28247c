set -x
mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests/test_reference_on_engine.py tests/test_lazy_engine.py tests/test_cabi.py -m gpu -x -q -s > gpurun_out/r2/t5_reference.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t5_reference.log
timeout 2400 python -m pytest tests/test_engine_parity.py -m gpu -x -q --durations=8 > gpurun_out/r2/t5_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t5_parity.log
timeout 2400 python -m pytest tests/test_aes_engine.py -m gpu -x -q --durations=8 > gpurun_out/r2/t5_aes.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t5_aes.log
timeout 1200 python -m pytest tests/test_snap.py tests/test_multi_rank.py tests/test_batch.py -m gpu -x -q --durations=5 > gpurun_out/r2/t5_rest.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t5_rest.log
