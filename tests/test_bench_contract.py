"""bench.py prints one JSON line with the keys the driver reads: CPU arm here, GPU arm on a B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}  # fmt: skip


def _run(args, timeout=600):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)  # fmt: skip
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    j = _run(["--impl", "reference", "--workload", "small", "--steps", "3", "--warmup", "1"])
    assert j["impl"] == "reference" and BASE_KEYS <= set(j)
    assert j["metric"] == "cell_updates_per_s" and j["unit"] == "cell-updates/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["gpu_launches"] == 0 and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["workload"] == "small"


def test_reference_arm_other_ranks_stay_silent():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))  # fmt: skip
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    j = _run(["--workload", "small", "--steps", "20", "--warmup", "3", "--cpu-budget", "2", "--parity-updates", "40"])
    assert BASE_KEYS | {"roofline", "clocks", "cpu_baseline"} <= set(j)
    assert j["n_gpus"] == 1 and j["steps"] == 20 and j["value"] > 0 and j["scaling"] == "weak"
    r = j["roofline"]
    assert r["bound"] in ("hbm", "issue", "latency") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["dense_sweep"]["bound"] == "hbm"
    assert j["gpu_launches"] > 0 and j["gpu_launches"] % 20 == 0
    e = j["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["mirror_matches_download"]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1
    assert cb["parity"]["full_grid"]["fire_map_equal"] is True and cb["parity"]["windows"]["all_equal"] is True
    assert j["fire_age"]["updates"] == 2000
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
