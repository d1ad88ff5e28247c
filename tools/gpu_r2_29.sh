set -x
O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_batch.py -m gpu -x -q > $O/t29_parity.log 2>&1; echo "rc=$?" >> $O/t29_parity.log
for v in base fold0; do
  lib=$PWD/$V/libckks_$v.so; [ $v = base ] && lib=$PWD/aes-implementation-fhe_b200/lib/libckks_b200.so
  CKKS_B200_LIB=$lib timeout 300 python tools/ntt_sizes.py > $O/ntt29_$v.json 2> $O/ntt29_$v.err
  CKKS_B200_LIB=$lib timeout 600 python tools/batch_scaling.py > $O/bs29_$v.json 2> $O/bs29_$v.err
  CKKS_B200_LIB=$lib timeout 600 python bench.py --no-cpu --no-dec --steps 2 --warmup 1 > $O/bench29_$v.json 2> $O/bench29_$v.err
done
