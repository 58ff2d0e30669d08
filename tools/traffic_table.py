"""Tabulate tools/traffic_sweep.sh output: per GEMM instantiation, mean DRAM MB and time per launch."""
import collections, csv, glob, re, sys
for path in sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/traffic_gm*_h*.csv")):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    agg = collections.OrderedDict()
    for r in rows:
        m = re.search(r"gemm_tc_kernel<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+)", r["Kernel Name"])
        k = f"<{m.group(1)},{m.group(2)},{m.group(3)}>" if m else r["Kernel Name"][:30]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"].startswith("dram"):
            v *= {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}[u]
            agg.setdefault(k, [0.0, 0.0, set()])[0] += v
        else:
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
            agg.setdefault(k, [0.0, 0.0, set()])[1] += v
        agg[k][2].add(r["ID"])
    print(path.split("/")[-1], "  ".join(f"{k}: {a[0] / len(a[2]):.0f} MB {a[1] / len(a[2]):.0f} us" for k, a in agg.items()))
