"""profiles/traffic.json from the committed ncu launch lists (profiles/r2f_<config>_launches.csv): DRAM bytes and duration
of every kernel of ONE full-batch launch set (zero_kernel + chdb_jit_select + chdb_jit_gather), serialised, cold L2.
bench.py reads dram_bytes_per_launch as roofline.traffic.  Usage: python profiles/make_traffic.py"""
import csv
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
FULL_BATCH = {"C2": 1 << 22, "C3": 1 << 22, "C4": 1 << 21, "C5": 1 << 22}
out = {}
for cfg, rows in FULL_BATCH.items():
    path = os.path.join(HERE, f"r2g_{cfg}_launches.csv")   # (re-captured after the last change to that config's kernels)
    if not os.path.exists(path):
        path = os.path.join(HERE, f"r2f_{cfg}_launches.csv")
    if not os.path.exists(path):
        continue
    lines = [l for l in open(path) if l.startswith('"')]
    launches = {}
    for r in csv.DictReader(lines):
        k = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"].split("(")[0].replace("chdb::", ""), "grid": r["Grid Size"]})
        k[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    ids = sorted(launches)
    # launch sets = zero, select, gather; take the LAST complete set whose gather grid is a full batch's
    sets = [ids[i:i + 3] for i in range(0, len(ids) - 2, 3)]
    full = [s for s in sets if launches[s[2]]["name"] == "chdb_jit_gather"
            and int(launches[s[2]]["grid"].strip("()").split(",")[0]) == rows // 512]
    s = full[-1]
    kern = {launches[i]["name"]: {"us": round(launches[i]["gpu__time_duration.sum"] / 1e3, 2),
                                  "dram_read": int(launches[i]["dram__bytes_read.sum"]),
                                  "dram_write": int(launches[i]["dram__bytes_write.sum"])} for i in s}
    rd = sum(k["dram_read"] for k in kern.values())
    wr = sum(k["dram_write"] for k in kern.values())
    out[cfg] = {"batch_rows": rows, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "kernels": kern,
                "duration_us_under_ncu": round(sum(k["us"] for k in kern.values()), 2),
                "capture": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                           "(serialised, cold L2), the last full-batch launch set (zero_kernel + chdb_jit_select + chdb_jit_gather) "
                           f"of profiles/{os.path.basename(path)}; the gather kernel's re-read of the predicate columns counts in "
                           "full here, and part of the output is still dirty in L2 when the last kernel ends"}
json.dump(out, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
for c, v in out.items():
    print(c, v["batch_rows"], v["dram_bytes_per_launch"], v["duration_us_under_ncu"], {k: x["us"] for k, x in v["kernels"].items()})
