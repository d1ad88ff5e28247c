O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
for v in base abulkst; do
  lib=$PWD/$V/libckks_$v.so; [ $v = base ] && lib=$PWD/aes-implementation-fhe_b200/lib/libckks_b200.so
  CKKS_B200_LIB=$lib timeout 300 python tools/ntt_sizes.py > $O/ntt39_$v.json 2> $O/ntt39_$v.err
done
CKKS_B200_LIB=$PWD/$V/libckks_abulkst.so timeout 900 python -m pytest tests/test_engine_parity.py -m gpu -x -q > $O/t39_parity.log 2>&1; echo "rc=$?" >> $O/t39_parity.log
