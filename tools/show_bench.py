"""Pretty-print the JSON line(s) bench.py / tools/c5_sweep.py wrote to a file:  python tools/show_bench.py FILE"""
import json
import sys


def show(d, ind=0):
    for k, v in (d or {}).items():
        if isinstance(v, dict) and any(isinstance(x, dict) for x in v.values()) or k in (
                "roofline", "roofline_hbm", "e2e", "e2e_u8", "cpu_baseline", "phases_ms_per_step", "clocks", "dp_parity"):
            print(" " * ind + k + ":")
            show(v, ind + 2)
        else:
            s = json.dumps(v)
            print(" " * ind + f"{k}: {s[:170]}")


for line in open(sys.argv[1]):
    if line.startswith("{"):
        show(json.loads(line))
        print()
