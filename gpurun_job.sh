timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "mirror" 2>&1 | tail -3
for v in host device; do if [ $v = device ]; then export SFB_LOG_ON_DEVICE=1; fi
SFB_DEBUG_TIMING=1 SFB_HOST_THREADS=8 python bench.py --no-cpu-baseline --steps 100 --e2e-steps 20 > gpurun_out/b.json 2> gpurun_out/b.err; grep sfb_sync gpurun_out/b.err | tail -3; python - <<PY
import json
j=json.loads(open("gpurun_out/b.json").read().strip().splitlines()[-1])
print("$v value %.3e ms/step %.3f e2e %.3e eval_ms %.3f" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["k_eval_ms_per_launch"]), j["e2e"]["mirror_matches_download"])
PY
done
