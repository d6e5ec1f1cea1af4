"""profiles/traffic.json from an `ncu --set full` capture of the benched build: DRAM bytes per launch of each kernel
(dram__bytes_read.sum + dram__bytes_write.sum), stored with the commit, workload and env count so that bench.py
reports `roofline.traffic` only for the configuration the capture was taken on.

    python tools/ncu_traffic.py gpurun_out/<capture>.ncu-rep C3 4096 [commit]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, workload, envs = sys.argv[1], sys.argv[2], int(sys.argv[3])
    commit = sys.argv[4] if len(sys.argv) > 4 else subprocess.run(
        ["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].split("::")[-1].replace("void ", "").strip()
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[col[m]].replace(",", "")) * scale.get(units[col[m]], 1.0)
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        du = units[col["gpu__time_duration.sum"]]
        dur_us = dur * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(du, 1.0)
        per.setdefault(name, []).append((tot, dur_us))
    kernels = {k: {"dram_bytes_per_launch": sum(t for t, _ in v) / len(v), "us_per_launch_under_ncu": sum(d for _, d in v) / len(v),
                   "launches": len(v)} for k, v in per.items()}
    pair = sum(v["dram_bytes_per_launch"] for k, v in kernels.items() if k.startswith("gnn_layers") or k.startswith("head"))
    rec = {"commit": commit, "workload": workload, "envs": envs, "source": os.path.basename(rep), "kernels": kernels,
           "policy_pair_bytes": pair}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
