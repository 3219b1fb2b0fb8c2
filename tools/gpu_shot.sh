#!/bin/bash
# One short GPU-box session: parity first, then the bench with and without unit skipping, then
# an ncu launch list.  Every stage has its own timeout and writes into gpurun_out/ as it goes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out/shot
O=gpurun_out/shot
date +%s > $O/t0
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
( timeout 330 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_spread_graph.py tests/test_gpu_scale.py \
    -m gpu -x -q --timeout 150 --durations=10 -p no:cacheprovider > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log )
date +%s > $O/t1
( timeout 100 python bench.py --no-cpu-baseline > $O/bench_skip_auto.json 2> $O/bench_skip_auto.err; echo "rc=$?" >> $O/bench_skip_auto.err )
date +%s > $O/t2
( timeout 100 python bench.py --no-cpu-baseline --unit-skip off > $O/bench_skip_off.json 2> $O/bench_skip_off.err; echo "rc=$?" >> $O/bench_skip_off.err )
date +%s > $O/t3
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log )
( timeout 100 python bench.py --no-cpu-baseline --workload cfg3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "rc=$?" >> $O/bench_cfg3.err )
date +%s > $O/t4
( timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 --burn-in 60 --roofline-steps 2 --e2e-steps 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1; echo "rc=$?" >> $O/ncu_bench.log )
date +%s > $O/t5
tail -3 $O/pytest.log; cat $O/bench_skip_auto.json | cut -c1-400; cat $O/bench_skip_off.json | cut -c1-400
