O=gpurun_out/r2; mkdir -p $O
for c in 0 1 2; do
  CKKS_NTT_CLUSTER=$c timeout 300 python tools/ntt_sizes.py > $O/ntt37_c$c.json 2> $O/ntt37_c$c.err
done
CKKS_NTT_CLUSTER=2 timeout 900 python -m pytest tests/test_engine_parity.py -m gpu -x -q > $O/t37_parity_c2.log 2>&1; echo "rc=$?" >> $O/t37_parity_c2.log
