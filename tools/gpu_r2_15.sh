set -x
O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_batch.py -m gpu -x -q > $O/t15_parity.log 2>&1; echo "rc=$?" >> $O/t15_parity.log
run() {  # name, lib, extra env
  name=$1; lib=$2; shift 2
  env CKKS_B200_LIB=$PWD/$lib "$@" timeout 600 python tools/batch_scaling.py > $O/bs15_$name.json 2> $O/bs15_$name.err
  env CKKS_B200_LIB=$PWD/$lib "$@" timeout 600 python bench.py --no-cpu --no-dec --steps 2 --warmup 1 > $O/bench15_$name.json 2> $O/bench15_$name.err
}
run b4 aes-implementation-fhe_b200/lib/libckks_b200.so
run b3 $V/libckks_v3b3.so
run b2 $V/libckks_v3b2.so
run p61_b4 aes-implementation-fhe_b200/lib/libckks_b200.so 'CKKS_B200_ENGINE_OVERRIDES={"p_bits":61}'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches15_ks_b4.csv python tools/ks_batch_once.py 4 > $O/ncu15a.log 2>&1
