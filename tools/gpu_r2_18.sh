set -x
O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
for v in base faketw1 faketw2; do
  lib=$PWD/$V/libckks_$v.so; [ $v = base ] && lib=$PWD/aes-implementation-fhe_b200/lib/libckks_b200.so
  CKKS_B200_LIB=$lib timeout 300 python tools/ntt_sizes.py > $O/ntt18_$v.json 2> $O/ntt18_$v.err
done
cat $O/ntt18_*.json
