timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for w in target cfg3; do for sw in tma ldg; do python bench.py --workload $w --sweep $sw --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_${w}_${sw}.json 2>&1; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_${w}_${sw}.json").read().strip().splitlines()[-1])
    r=j["roofline"]; print("$w $sw value %.3e ms/step %.3f sweep_ms %.3f eval_ms %.3f frac %.3f items %d" % (j["value"], j["ms_per_step"], r["ms_per_launch"], r["k_eval_ms_per_launch"], r["frac"], r["work_items_per_step"]))
except Exception as e: print("$w $sw FAILED", e, open("gpurun_out/bench_${w}_${sw}.json").read()[-800:])
PY
done; done
