#!/usr/bin/env python
"""
Turn the ncu artefacts a gpurun call brought back into the small, committed summaries under
profiles/ (the .ncu-rep files themselves are scratch and stay in gpurun_out/).

    python tools/summarize_ncu.py --rep gpurun_out/prof_r01_target.ncu-rep \
        --launches gpurun_out/launches_r01.csv --round r01 --workload target

Writes profiles/<round>_kernels.json (per-launch metrics of the captured kernels),
profiles/<round>_launch_shares.json (+ a copy of the launch list) and refreshes
profiles/ncu_sweep_summary.json, which bench.py reads for `roofline.traffic`.
"""
import argparse
import collections
import csv
import io
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]  # fmt: skip
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
              "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}  # fmt: skip


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    ap.add_argument("--round", default="r01")
    ap.add_argument("--workload", default="target")
    ap.add_argument("--command", default="")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

    if a.rep:
        hdr, units, rows = raw_page(a.rep)
        idx = {h: i for i, h in enumerate(hdr)}
        kernels = []
        for r in rows:
            k = {"kernel": r[idx["Kernel Name"]]}
            for m in METRICS:
                if m in idx:
                    try:
                        v = float(r[idx[m]].replace(",", ""))
                    except ValueError:
                        continue
                    u = units[idx[m]]
                    k[m] = v * UNIT_SCALE[u] if (m.endswith("bytes_read.sum") or m.endswith("bytes_write.sum") or
                                                 m.endswith("t_bytes.sum") or m.startswith("gpu__time")) and u in UNIT_SCALE else v  # fmt: skip
            if "dram__bytes_read.sum" in k:
                k["dram_bytes"] = k["dram__bytes_read.sum"] + k.get("dram__bytes_write.sum", 0.0)
                k["dram_GBps"] = k["dram_bytes"] / k["gpu__time_duration.sum"] / 1e9
            kernels.append(k)
        doc = {"source": os.path.basename(a.rep), "workload": a.workload, "command": a.command,
               "note": "ncu --set full --clock-control none: each launch replayed ~40x, cold caches; durations are "
                       "not bench numbers", "kernels": kernels}  # fmt: skip
        with open(os.path.join(ROOT, "profiles", f"{a.round}_kernels.json"), "w") as f:
            json.dump(doc, f, indent=1)
        sweeps = [k for k in kernels if "k_sweep" in k["kernel"] and "dram_bytes" in k]
        if sweeps:
            s = {"workload": a.workload, "round": a.round, "kernel": sweeps[0]["kernel"],
                 "dram_bytes_per_launch": sum(k["dram_bytes"] for k in sweeps) / len(sweeps),
                 "ncu_duration_s": sum(k["gpu__time_duration.sum"] for k in sweeps) / len(sweeps),
                 "launches_averaged": len(sweeps), "source": os.path.basename(a.rep)}  # fmt: skip
            with open(os.path.join(ROOT, "profiles", "ncu_sweep_summary.json"), "w") as f:
                json.dump(s, f, indent=1)
            print("k_sweep dram bytes/launch", s["dram_bytes_per_launch"], "duration", s["ncu_duration_s"])

    if a.launches:
        dst = os.path.join(ROOT, "profiles", f"{a.round}_launches.csv")
        shutil.copy(a.launches, dst)
        rows = [r for r in csv.reader(open(a.launches)) if len(r) > 5 and not r[0].startswith("==")]
        hdr = rows[0]
        ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        tot, cnt = collections.Counter(), collections.Counter()
        for r in rows[1:]:
            v = float(r[iv].replace(",", "")) * UNIT_SCALE.get(r[iu], 1.0)
            name = r[ik].split("(")[0].replace("void ", "").strip()
            tot[name] += v
            cnt[name] += 1
        total = sum(tot.values())
        shares = {k: {"launches": cnt[k], "total_ms": 1e3 * tot[k], "avg_us": 1e6 * tot[k] / cnt[k], "share": tot[k] / total}
                  for k in sorted(tot, key=lambda k: -tot[k])}  # fmt: skip
        step = {k: v for k, v in tot.items() if any(t in k for t in ("k_sweep", "k_rows", "k_eval", "k_row_list", "k_units"))}
        st = sum(step.values())
        doc = {"source": os.path.basename(a.launches), "workload": a.workload, "command": a.command,
               "note": "ncu --metrics gpu__time_duration.sum: serialised, cold-cache launch times; compare shares",
               "all_kernels": shares, "share_within_step": {k: v / st for k, v in step.items()} if st else {}}  # fmt: skip
        with open(os.path.join(ROOT, "profiles", f"{a.round}_launch_shares.json"), "w") as f:
            json.dump(doc, f, indent=1)
        print(json.dumps(doc["share_within_step"], indent=1))


if __name__ == "__main__":
    main()
