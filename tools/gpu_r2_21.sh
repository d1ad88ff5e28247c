set -x
O=gpurun_out/r2; mkdir -p $O
NTT_ONCE=15,4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_ -c 8 -o $O/ncu21_ntt60 python tools/ntt_once.py > $O/ncu21.log 2>&1
