"""
Golden files for `FireSimulation._save_data` (simfire/sim/simulation.py:887-959): runs the
UNMODIFIED reference (headless, under ref_shim's stand-ins) with `save_data: true`,
`data_type: npy`, and records what it wrote under <sf_home>/data/<start_time>/.  Dev container
only (needs /root/reference).  h5 / jsonl need h5py / jsonlines, which this image lacks; the
jsonl line format is the one `jsonlines.Writer.write` produces (`json.dumps(obj) + "\\n"`).

    python tests/golden/gen_savedata_golden.py   # rewrites tests/golden/savedata_npy.npz
"""
from __future__ import annotations

import copy
import json
import os
import shutil
import sys

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
os.environ.setdefault("LOGLEVEL", "ERROR")
import ref_shim  # noqa: E402

ref_shim.install()
from gen_api_golden import make_config_dict  # noqa: E402
from simfire.sim.simulation import FireSimulation  # noqa: E402
from simfire.utils.config import Config  # noqa: E402


def main():
    home = "/tmp/sf_home_savedata_golden"
    shutil.rmtree(home, ignore_errors=True)
    cfg = make_config_dict((48, 48), "flat", (20, 24), max_dur=3)  # square: the reference fuel image assumes it
    cfg["simulation"]["save_data"] = True
    cfg["simulation"]["data_type"] = "npy"
    cfg["simulation"]["sf_home"] = home
    sim = FireSimulation(Config(config_dict=copy.deepcopy(cfg)))
    sim.run(3)
    sim.update_mitigation([(x, 30, 3) for x in range(5, 40)])
    sim.run(4)
    datapath = os.path.join(home, "data", sim.start_time)
    files = sorted(os.listdir(datapath))
    meta = json.load(open(os.path.join(datapath, "metadata.json")))
    out = {"files": np.array(files), "metadata_json": json.dumps(meta), "config_yaml": yaml.safe_dump(cfg),
           "fire_map": np.load(os.path.join(datapath, "fire_map.npy"))}
    for f in files:
        if f.endswith(".npy") and f != "fire_map.npy":
            out["static_" + f[:-4]] = np.load(os.path.join(datapath, f))
    np.savez_compressed(os.path.join(HERE, "savedata_npy.npz"), **out)
    print("files", files, "fire_map history", out["fire_map"].shape, out["fire_map"].dtype, "elapsed_steps", sim.elapsed_steps)
    print("metadata keys", sorted(meta), "shape", meta["shape"], "static", meta["static_data"])


if __name__ == "__main__":
    main()
