"""Per-kernel table from an ncu metrics CSV of one eager step (long format: one row per launch and metric).

``python tools/kernel_table.py gpurun_out/step_metrics.csv [--peak-gbs 6458.4] > profiles/...md``

The CSV comes from
``ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
--clock-control none -s <skip> -c <n> --csv --log-file <csv> python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline``.
Times are cold-cache (ncu flushes L2 between kernels) and serialised: DRAM GB/s here is the traffic ncu counted divided
by that time, i.e. what the kernel pulls from HBM when nothing is L2-resident.
"""
import argparse
import csv
import json
import os
import re
import sys
from collections import OrderedDict


def to_float(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--peak-gbs", type=float, default=None)
    args = ap.parse_args()
    peak = args.peak_gbs
    if peak is None:
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        peak = json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0
    rows = [l for l in open(args.csv) if l.startswith('"')]
    launches = OrderedDict()
    for r in csv.DictReader(rows):
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"]})
        val, unit = to_float(r["Metric Value"]), r["Metric Unit"]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}.get(unit, 1.0)
        d[r["Metric Name"]] = val * scale
    agg = OrderedDict()
    for d in launches.values():
        name = re.sub(r"^void ", "", d["name"])
        name = re.sub(r"\(.*$", "", name)
        name = name[:60]
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tc": 0.0})
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tc"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * d.get("gpu__time_duration.sum", 0.0)
    total = sum(a["us"] for a in agg.values())
    print(f"| kernel | launches | us/launch | share | DRAM read MB/launch | DRAM write MB/launch | DRAM GB/s | of {peak:.0f} GB/s | tensor pipe active |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
        print(f"| `{name}` | {a['n']} | {a['us'] / a['n']:.2f} | {100 * a['us'] / total:.1f}% | {a['rd'] / a['n'] / 1e6:.2f} | "
              f"{a['wr'] / a['n'] / 1e6:.2f} | {gbs:.0f} | {100 * gbs / peak:.0f}% | {a['tc'] / a['us'] if a['us'] else 0:.1f}% |")
    print(f"\n{sum(a['n'] for a in agg.values())} launches, {total / 1e3:.2f} ms in total (cold-cache, serialised).", file=sys.stdout)


if __name__ == "__main__":
    main()
