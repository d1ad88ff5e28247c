set -x
O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
export CKKS_B200_LIB=$PWD/$V/libckks_bcfp1.so
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_batch.py -m gpu -x -q > $O/t10_parity_bcfp.log 2>&1; echo "rc=$?" >> $O/t10_parity_bcfp.log
run() {  # name, lib, extra env
  name=$1; lib=$2; shift 2
  env CKKS_B200_LIB=$PWD/$V/$lib "$@" timeout 600 python tools/batch_scaling.py > $O/bs10_$name.json 2> $O/bs10_$name.err
  env CKKS_B200_LIB=$PWD/$V/$lib "$@" timeout 600 python bench.py --no-cpu --no-dec --steps 2 --warmup 1 > $O/bench10_$name.json 2> $O/bench10_$name.err
}
run int libckks_bcfp1.so CKKS_BC_FP=0
run fp1 libckks_bcfp1.so
run fp3 libckks_bcfp3.so
run fp4 libckks_bcfp4.so
run fp1_int3 libckks_bcfp1.so CKKS_BC_FP_INT_EVERY=3
run fp1_int2 libckks_bcfp1.so CKKS_BC_FP_INT_EVERY=2
