#!/bin/bash
# First GPU session of the next round: everything that was added after round 1's last GPU minute and
# still needs a B200 number.  Usage (from the dev container):
#   python tools/build_variants.py && gpurun --timeout 900 -- 'bash tools/gpu_ab_next.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/next; mkdir -p $O
B="python bench.py --no-cpu-baseline"
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err ); }
# 1. parity with the final sources (new tests: random scenarios, patch paths, save_data, ...)
( timeout 300 python -m pytest tests -m gpu -x -q --timeout 150 --durations=8 -p no:cacheprovider > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
# 2. the official line, and where the e2e time goes with the one-pass patch
run target_default 120 python bench.py
SFB_DEBUG_TIMING=1 run target_e2e_debug 90 $B --e2e-steps 40
# 3. k_rows variants (tools/build_variants.py leaves one library per macro next to the default one)
for v in redux padvec v2; do
  SFB_LIB=$PWD/simfire_b200/libsimfire_b200_$v.so run target_rows_$v 90 $B
  SFB_LIB=$PWD/simfire_b200/libsimfire_b200_$v.so run cfg3_rows_$v 90 $B --workload cfg3
done
run cfg3 90 $B --workload cfg3
# 4. full burns and the small configurations
run cfg1_full_burn 120 python bench.py --workload cfg1 --full-burn
run cfg2_full_burn 200 python bench.py --workload cfg2 --full-burn
run cfg2 60 $B --workload cfg2
# 5. launch list + full capture of the default build
( timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 --burn-in 60 --roofline-steps 2 --e2e-steps 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1 )
( timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_row_list|k_rows|k_eval" -s 996 -c 6 -f -o $O/rows \
    python bench.py --steps 20 --warmup 3 --burn-in 60 --roofline-steps 2 --e2e-steps 3 --no-cpu-baseline > $O/ncu_full.log 2>&1 )
tail -2 $O/pytest.log; for f in $O/target_*.json; do echo $f; cut -c1-220 $f; done
