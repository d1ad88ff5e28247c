set -x
O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
P50='CKKS_B200_ENGINE_OVERRIDES={"p_bits":50}'
env CKKS_B200_LIB=$PWD/$V/libckks_v2b4.so timeout 900 python -m pytest tests/test_engine_parity.py tests/test_batch.py -m gpu -x -q > $O/t13_parity_p61.log 2>&1; echo "rc=$?" >> $O/t13_parity_p61.log
run() {  # name, lib, extra env
  name=$1; lib=$2; shift 2
  env CKKS_B200_LIB=$PWD/$V/$lib "$@" timeout 600 python tools/batch_scaling.py > $O/bs13_$name.json 2> $O/bs13_$name.err
  env CKKS_B200_LIB=$PWD/$V/$lib "$@" timeout 600 python bench.py --no-cpu --no-dec --steps 2 --warmup 1 > $O/bench13_$name.json 2> $O/bench13_$name.err
}
run p50_b4 libckks_v2b4.so "$P50"
run p50_b3 libckks_v2b3.so "$P50"
run p50_b2 libckks_v2b2.so "$P50"
run p61_b4 libckks_v2b4.so
env CKKS_B200_LIB=$PWD/$V/libckks_v2b4.so "$P50" timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches13_ks_b4_p50.csv python tools/ks_batch_once.py 4 > $O/ncu13a.log 2>&1
