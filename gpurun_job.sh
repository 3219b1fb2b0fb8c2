# launch list of the default bench command (short step counts; shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 20 --warmup 3 --burn-in 60 --roofline-steps 2 --e2e-steps 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log | cut -c1-300
# full capture of the two hot-path kernels at the target workload (launches after the burn-in)
ncu --set full --clock-control none --import-source on -k regex:"k_sweep|k_eval" -s 150 -c 2 -o gpurun_out/prof_r01_target python bench.py --steps 20 --warmup 3 --burn-in 60 --roofline-steps 2 --e2e-steps 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-200
python bench.py > gpurun_out/bench_r01_target.json 2> gpurun_out/bench_r01_target.err; tail -c 600 gpurun_out/bench_r01_target.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.err; tail -c 1200 gpurun_out/bench_r01_reference.json; tail -3 gpurun_out/bench_r01_reference.err; nproc; lscpu | grep "Model name"
