timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --no-cpu-baseline > gpurun_out/bench_track.json 2> gpurun_out/bench_track.err; python - <<PY
import json
j=json.loads(open("gpurun_out/bench_track.json").read().strip().splitlines()[-1])
print("value %.3e ms/step %.3f" % (j["value"], j["ms_per_step"])); print(j["e2e"]); print({k:v for k,v in j["roofline"].items() if k!="survey_model"})
PY
tail -3 gpurun_out/bench_track.err
