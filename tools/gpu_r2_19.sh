set -x
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_batch.py -m gpu -x -q > $O/t19_parity.log 2>&1; echo "rc=$?" >> $O/t19_parity.log
timeout 300 python tools/ntt_sizes.py > $O/ntt19.json 2> $O/ntt19.err
timeout 600 python tools/batch_scaling.py > $O/bs19.json 2> $O/bs19.err
timeout 600 python bench.py --no-cpu --no-dec --steps 2 --warmup 1 > $O/bench19.json 2> $O/bench19.err
