timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # workload sweep extra
python bench.py --workload $1 --sweep $2 $3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/b.json 2>&1; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/b.json").read().strip().splitlines()[-1])
    r=j["roofline"]; print("$1 $2 $3 value %.3e ms/step %.3f sweep_ms %.3f eval_ms %.3f frac %.3f" % (j["value"], j["ms_per_step"], r["ms_per_launch"], r["k_eval_ms_per_launch"], r["frac"]))
except Exception as e: print("$1 $2 FAILED", e, open("gpurun_out/b.json").read()[-600:])
PY
}
run target tma ""; run cfg3 tma ""; run cfg3 tma "--rows-per-chunk 14"; run cfg3 ldg ""; run cfg3 ldg "--rows-per-chunk 16"
python bench.py --workload cfg5 --steps 200 --burn-in 200 > gpurun_out/bench_cfg5_n1.json 2>gpurun_out/bench_cfg5_n1.err; tail -c 900 gpurun_out/bench_cfg5_n1.json; tail -3 gpurun_out/bench_cfg5_n1.err
python bench.py --workload cfg2 --steps 200 --burn-in 200 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2>gpurun_out/bench_cfg2.err; tail -c 1500 gpurun_out/bench_cfg2.json | cut -c1-700; tail -3 gpurun_out/bench_cfg2.err
