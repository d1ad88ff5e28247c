set -x
O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
export CKKS_B200_LIB=$PWD/$V/libckks_bcfp4.so
export CKKS_B200_ENGINE_OVERRIDES='{"p_bits":50}'
timeout 300 python tools/ks_batch_once.py 4 > $O/ks12.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches12_ks_b4_p50.csv python tools/ks_batch_once.py 4 > $O/ncu12a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_base_convert_fp -c 6 -o $O/ncu12_bcfp python tools/ks_batch_once.py 4 > $O/ncu12b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ks_inner -c 3 -o $O/ncu12_inner python tools/ks_batch_once.py 4 > $O/ncu12c.log 2>&1
ls -la $O/*.ncu-rep
