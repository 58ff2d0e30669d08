"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
   python tools/ncu_summary.py gpurun_out/launches.csv > profiles/rNN_<what>_launches.md"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("vbnn::<unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        a = agg.setdefault(name, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"launches: {n}   total device time: {tot / 1e3:.3f} ms (cold-cache, serialised under ncu: compare SHARES)\n")
    print("| share | total us | launches | avg us | grid | block | kernel |")
    print("|---:|---:|---:|---:|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * a[1] / tot:.1f}% | {a[1]:.1f} | {a[0]} | {a[1] / a[0]:.1f} | {a[2]} | {a[3]} | `{k}` |")


if __name__ == "__main__":
    main(sys.argv[1])
