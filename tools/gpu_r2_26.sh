set -x
O=gpurun_out/r2; mkdir -p $O
NTT_ONCE=21,6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_ -c 8 -o $O/ncu26_ntt126 python tools/ntt_once.py > $O/ncu26a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_base_convert_fp|k_ks_inner" -c 8 -o $O/ncu26_ks python tools/ks_batch_once.py 4 > $O/ncu26b.log 2>&1
NTT_ONCE=21,6 timeout 300 python tools/ntt_once.py > $O/ntt26_once.txt 2>&1
