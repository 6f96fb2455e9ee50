"""Condense ncu --csv launch lists into the tables committed under profiles/.
usage: python profiles/summarize_launches.py launches gpurun_out/launches.csv   > profiles/<name>.txt
       python profiles/summarize_launches.py traffic  gpurun_out/gemm_traffic.csv > profiles/<name>.txt  (also writes gemm_traffic.json)"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict


def rows(path):
    lines = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    m = re.match(r"([\w:]+(?:<[^>]*>)?)", name)
    return (m.group(1) if m else name)[:64]


def launches(path):
    agg = OrderedDict()
    for r in rows(path):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print("(%d launches ~ one training step; cold-cache serialised times: shares matter, not absolutes)" % n)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-66s %3d launches %9.1f us %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    print("total us %.1f launches %d" % (tot, n))
    gemm = sum(v[1] for k, v in agg.items() if k.startswith("gemm_kernel"))
    print("gemm_kernel share of the step: %.1f%%" % (100 * gemm / tot))


def traffic(path):
    per = OrderedDict()
    for r in rows(path):
        d = per.setdefault(r["ID"], dict(name=short(r["Kernel Name"])))
        d[r["Metric Name"]] = float(r["Metric Value"]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3,
                                                          "usecond": 1e3, "nsecond": 1}.get(r["Metric Unit"], 1)
    agg = OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("dram__bytes_read.sum", 0.0)
        a[2] += d.get("dram__bytes_write.sum", 0.0)
        a[3] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    n = sum(v[0] for v in agg.values())
    rd, wr, us = sum(v[1] for v in agg.values()), sum(v[2] for v in agg.values()), sum(v[3] for v in agg.values())
    print("all %d GEMM launches of one B=256 training step: DRAM read %.1f MB + write %.1f MB = %.1f MB per step; summed "
          "(cold, serialised) duration %.1f us" % (n, rd / 1e6, wr / 1e6, (rd + wr) / 1e6, us))
    for k, v in agg.items():
        print("%-44s launches %3d  dram_rd %9.1f MB  dram_wr %9.1f MB  time %8.1f us" % (k, v[0], v[1] / 1e6, v[2] / 1e6, v[3]))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_traffic.json")
    json.dump(dict(gemm_launches_per_step=n, dram_bytes_per_step=rd + wr, dram_bytes_per_launch=(rd + wr) / n,
                   source="profiles/%s (ncu, B=256 training step)" % os.environ.get("TRAFFIC_TXT", "r1_final_gemm_dram_traffic.txt")),
              open(out, "w"))


if __name__ == "__main__":
    {"launches": launches, "traffic": traffic}[sys.argv[1]](sys.argv[2])
