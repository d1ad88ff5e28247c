set -x
O=gpurun_out/r2; mkdir -p $O
export CKKS_B200_LIB=$PWD/aes-implementation-fhe_b200/lib/variants/libckks_pipe1.so
NTT_ONCE=24,6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_fwd_passA -c 3 -o $O/ncu22_pipe python tools/ntt_once.py > $O/ncu22.log 2>&1
