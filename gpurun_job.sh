timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { # lib workload sweep extra
python bench.py --workload $2 --sweep $3 $4 --no-cpu-baseline --e2e-steps 3 > gpurun_out/b.json 2>&1; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/b.json").read().strip().splitlines()[-1])
    r=j["roofline"]; print("$1 $2 $3 $4 value %.3e ms/step %.3f sweep_ms %.3f eval_ms %.3f frac %.3f" % (j["value"], j["ms_per_step"], r["ms_per_launch"], r["k_eval_ms_per_launch"], r["frac"]))
except Exception as e: print("$1 $2 $3 FAILED", e, open("gpurun_out/b.json").read()[-600:])
PY
}
for lib in libsimfire_b200 libsfb_w4_mb8 libsfb_w2_mb6 libsfb_w2_mb8; do export SFB_LIB=$PWD/simfire_b200/$lib.so; for w in target cfg3; do run $lib $w ldg ""; done; done
export SFB_LIB=$PWD/simfire_b200/libsimfire_b200.so; run default cfg3 ldg "--rows-per-chunk 16"; run default cfg3 ldg "--rows-per-chunk 8"
for lib in libsimfire_b200 libsfb_st3 libsfb_w2_mb6; do export SFB_LIB=$PWD/simfire_b200/$lib.so; for w in target cfg3; do run $lib $w tma ""; done; done
