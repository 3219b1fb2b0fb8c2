#!/bin/bash
# Final short GPU session of the round, most important first (the budget may cut it short).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/shot3; mkdir -p $O
B="python bench.py --no-cpu-baseline"
run() { name=$1; t=$2; shift 2; ( timeout $t "$@" > $O/$name.json 2> $O/$name.err; echo "rc=$?" >> $O/$name.err ); date +%s >> $O/times; }
date +%s > $O/times
( timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_spread_graph.py tests/test_gpu_scale.py \
    -m gpu -x -q --timeout 100 --durations=5 -p no:cacheprovider > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log ); date +%s >> $O/times
run target_default 90 python bench.py
run target_chunks 60 $B --units chunks
( timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 --burn-in 60 --roofline-steps 2 --e2e-steps 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "rc=$?" >> $O/ncu_bench.log ); date +%s >> $O/times
SFB_GROUP_GRAPH=0 run target_nograph 60 $B
SFB_LIB=$PWD/simfire_b200/libsimfire_b200_shfl6.so run target_shfl6 60 $B
( timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_row_list|k_rows|k_eval" -s 996 -c 6 -f -o $O/r01_rowunits \
    python bench.py --steps 20 --warmup 3 --burn-in 60 --roofline-steps 2 --e2e-steps 3 --no-cpu-baseline > $O/ncu_full.log 2>&1; echo "rc=$?" >> $O/ncu_full.log ); date +%s >> $O/times
run cfg3 60 $B --workload cfg3
run cfg2 60 $B --workload cfg2
run target_groups2 60 $B --env-groups 2
( SFB_UNIT_SKIP=1 timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 100 -p no:cacheprovider -k "not unit_skipping_changes_nothing" > $O/pytest_forced_rows.log 2>&1; echo "rc=$?" >> $O/pytest_forced_rows.log ); date +%s >> $O/times
run target_off 60 $B --unit-skip off
tail -2 $O/pytest.log; cut -c1-200 $O/target_default.json
run target_thp 60 $B --mirror thp
SFB_HOST_THREADS=8 run target_ht8 60 $B
run target_thp_groups2 60 $B --mirror thp --env-groups 2
