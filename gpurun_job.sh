timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # workload sweep extra
python bench.py --workload $1 --sweep $2 $3 --no-cpu-baseline > gpurun_out/b.json 2>&1; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/b.json").read().strip().splitlines()[-1])
    r=j["roofline"]; print("$1 $2 $3 value %.3e ms/step %.3f sweep_ms %.3f eval_ms %.3f frac %.3f e2e %.3e" % (j["value"], j["ms_per_step"], r["ms_per_launch"], r["k_eval_ms_per_launch"], r["frac"], j["e2e"]["value"]))
except Exception as e: print("$1 $2 FAILED", e, open("gpurun_out/b.json").read()[-600:])
PY
}
run target tma ""; run cfg3 tma ""
