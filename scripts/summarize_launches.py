"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel totals and the conv_tc sequence.

    python scripts/summarize_launches.py gpurun_out/launches.csv
"""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")
        rows.append((name.split("(")[0].split("<")[0], ms))
    tot = sum(ms for _, ms in rows)
    agg = OrderedDict()
    for k, ms in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    print(f"total {tot:.2f} ms over {len(rows)} launches")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {ms:8.3f} ms  {n:4d}x  {k}")
    tc = [ms for k, ms in rows if "conv_tc" in k]
    print("conv_tc", f"{sum(tc):.2f}", "::", " ".join(f"{x:.2f}" for x in tc))


if __name__ == "__main__":
    main(sys.argv[1])
