"""DRAM traffic of the timed kernel of a bench.py step, from ncu (run on the GPU box, AFTER the plain bench run exited 0):

    python scripts/ncu_traffic.py <git-sha> [workload ...]        -> gpurun_out/r2_traffic.json (copy to profiles/)

For every workload it runs `bench.py --workload W --steps 1 --warmup 1 --no-cpu --no-parity` under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:admm_kernel`
and records bytes per launch of the persistent kernel (one launch = one step).  bench.py reads the file back for
`roofline.traffic`; numbers printed by the bench run under ncu are discarded."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sha = sys.argv[1] if len(sys.argv) > 1 else "unknown"
workloads = sys.argv[2:] or ["cfg5"]
out_path = os.path.join(ROOT, "gpurun_out", "r2_traffic.json")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
rec = json.load(open(out_path)) if os.path.exists(out_path) else {}
for w in workloads:
    log = os.path.join(ROOT, "gpurun_out", f"r2_ncu_traffic_{w}.csv")
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "-k", "regex:admm_kernel", "--csv", "--log-file", log, sys.executable, os.path.join(ROOT, "bench.py"),
           "--workload", w, "--steps", "1", "--warmup", "1", "--no-cpu", "--no-parity"]
    r = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        print(f"ncu failed for {w}: {r.stderr[-400:]}", file=sys.stderr)
        continue
    text = open(log).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per = {}
    for row in rows:
        per.setdefault(row["ID"], {"kernel": row["Kernel Name"]})[row["Metric Name"]] = (
            float(row["Metric Value"].replace(",", "")), row["Metric Unit"])

    def to_bytes(v, unit):
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]

    launches = []
    for k in sorted(per, key=int):
        p = per[k]
        rd = to_bytes(*p["dram__bytes_read.sum"]); wr = to_bytes(*p["dram__bytes_write.sum"])
        t, tu = p["gpu__time_duration.sum"]
        ms = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(tu, 1e-6)
        launches.append({"kernel": p["kernel"], "dram_read": rd, "dram_write": wr, "ms_under_ncu": ms})
    if not launches:
        continue
    last = launches[-1]            # the timed step (the first launch is the warm-up step)
    rec[f"{w}_n1"] = {"dram_bytes_per_launch": last["dram_read"] + last["dram_write"], "dram_read": last["dram_read"],
                      "dram_write": last["dram_write"], "ms_under_ncu": last["ms_under_ncu"], "kernel": last["kernel"],
                      "launches_seen": len(launches), "git": sha,
                      "what": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the step's admm_kernel launch, "
                              "bench.py --steps 1 --warmup 1 --no-cpu --no-parity"}
    json.dump(rec, open(out_path, "w"), indent=1)
    print(w, rec[f"{w}_n1"])
