#!/bin/bash
# run bench.py (graph mode, no CPU leg) on every library variant built by tools/build_variants.sh; one line each
cd "$(dirname "$0")/.."
for lib in base aes-implementation-fhe_b200/lib/variants/*.so; do
  if [ "$lib" = base ]; then unset CKKS_B200_LIB; else export CKKS_B200_LIB=$PWD/$lib; fi
  timeout 200 python bench.py --no-cpu --steps 3 --warmup 2 ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'],2), round(d['value'],1), round(d['s_per_round']*1e3,2), d['bytes_exact_vs_fips197_round'])"
done
