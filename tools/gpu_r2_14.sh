set -x
O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
export CKKS_B200_LIB=$PWD/$V/libckks_v2b4.so
export CKKS_B200_ENGINE_OVERRIDES='{"p_bits":50}'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_base_convert_fp -c 6 -o $O/ncu14_bcfp python tools/ks_batch_once.py 4 > $O/ncu14.log 2>&1
