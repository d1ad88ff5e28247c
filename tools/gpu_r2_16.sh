set -x
O=gpurun_out/r2; mkdir -p $O
V=aes-implementation-fhe_b200/lib/variants
timeout 900 python -m pytest tests/test_engine_parity.py tests/test_batch.py -m gpu -x -q > $O/t16_parity.log 2>&1; echo "rc=$?" >> $O/t16_parity.log
run() {  # name, lib, extra env
  name=$1; lib=$2; shift 2
  env CKKS_B200_LIB=$PWD/$lib "$@" timeout 600 python tools/batch_scaling.py > $O/bs16_$name.json 2> $O/bs16_$name.err
}
run nb4 aes-implementation-fhe_b200/lib/libckks_b200.so
run nb2 $V/libckks_ksi2.so
run nb1 $V/libckks_ksi1.so
run old $V/libckks_ksi0.so
timeout 600 python bench.py --no-cpu --no-dec --steps 2 --warmup 1 > $O/bench16_nb4.json 2> $O/bench16_nb4.err
env CKKS_B200_LIB=$PWD/$V/libckks_ksi2.so timeout 600 python bench.py --no-cpu --no-dec --steps 2 --warmup 1 > $O/bench16_nb2.json 2> $O/bench16_nb2.err
