set -x
O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
run() {  # name, lib, extra env
  name=$1; lib=$2; shift 2
  env CKKS_B200_LIB=$PWD/$V/$lib "$@" timeout 600 python tools/batch_scaling.py > $O/bs11_$name.json 2> $O/bs11_$name.err
  env CKKS_B200_LIB=$PWD/$V/$lib "$@" timeout 600 python bench.py --no-cpu --no-dec --steps 2 --warmup 1 > $O/bench11_$name.json 2> $O/bench11_$name.err
}
run p50_fp4 libckks_bcfp4.so 'CKKS_B200_ENGINE_OVERRIDES={"p_bits":50}'
run p50_int libckks_bcfp4.so 'CKKS_B200_ENGINE_OVERRIDES={"p_bits":50}' CKKS_BC_FP=0
