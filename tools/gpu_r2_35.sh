O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
for v in base nostore noload; do
  lib=$PWD/$V/libckks_$v.so; [ $v = base ] && lib=$PWD/aes-implementation-fhe_b200/lib/libckks_b200.so
  CKKS_B200_LIB=$lib timeout 300 python tools/ntt_sizes.py > $O/ntt35_$v.json 2> $O/ntt35_$v.err
done
