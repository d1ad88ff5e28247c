set -x
mkdir -p gpurun_out/r2
nvidia-smi -L
timeout 1500 python -m pytest tests/test_batch.py tests/test_engine_parity.py tests/test_cabi.py -m gpu -x -q > gpurun_out/r2/t1_parity.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t1_parity.log
CKKS_NTT_CLUSTER=2 timeout 1500 python -m pytest tests/test_batch.py tests/test_engine_parity.py -m gpu -x -q > gpurun_out/r2/t1_parity_cluster2.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t1_parity_cluster2.log
timeout 900 python tools/batch_scaling.py > gpurun_out/r2/batch_scaling_c0.json 2> gpurun_out/r2/batch_scaling_c0.err
CKKS_NTT_CLUSTER=1 timeout 900 python tools/batch_scaling.py > gpurun_out/r2/batch_scaling_c1.json 2> gpurun_out/r2/batch_scaling_c1.err
CKKS_NTT_CLUSTER=2 timeout 900 python tools/batch_scaling.py > gpurun_out/r2/batch_scaling_c2.json 2> gpurun_out/r2/batch_scaling_c2.err
timeout 1200 python -m pytest tests/test_aes_engine.py -m gpu -x -q -k "many_pairs or captured_round_replays" > gpurun_out/r2/t1_aes.log 2>&1; echo "rc=$?" >> gpurun_out/r2/t1_aes.log
