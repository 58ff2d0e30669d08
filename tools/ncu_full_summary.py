"""Summarise an `ncu --set full` report: per-launch table (markdown) + per-kernel DRAM traffic (JSON).

   python tools/ncu_full_summary.py gpurun_out/prof.ncu-rep profiles/rNN_x_ncu_full.md profiles/rNN_x_traffic.json "title"
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def main(rep, md_out, json_out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    col = h.index
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

    def val(r, n):
        i = col(n)
        return float(r[i].replace(",", "")) * scale.get(u[i], 1.0)

    out = []
    for r in rows[2:]:
        name = r[col("Kernel Name")]
        m = re.search(r"gemm_tc_kernel<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+)", name) or re.search(r"gemm_tc_kernel<(\d+), (\d+), (\d+)", name)
        short = f"gemm_tc_kernel<{m.group(1)},{m.group(2)},{m.group(3)}>" if m else re.sub(r"\(.*", "", name).split("::")[-1]
        us = val(r, "gpu__time_duration.sum")
        us = us / 1e3 if u[col("gpu__time_duration.sum")] in ("ns", "nsecond") else us
        out.append(dict(kernel=short, us=us, dram_read=val(r, "dram__bytes_read.sum"), dram_write=val(r, "dram__bytes_write.sum"),
                        tensor_pct=val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        dram_pct=val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        l2_hit=val(r, "lts__t_sector_hit_rate.pct"), regs=int(float(r[col("launch__registers_per_thread")]))))
    per = collections.OrderedDict()
    for o in out:
        per.setdefault(o["kernel"], []).append(o)
    summary = {}
    for k, lst in per.items():
        n = len(lst)
        summary[k] = dict(launches_captured=n, us=sum(o["us"] for o in lst) / n,
                          traffic_bytes=sum(o["dram_read"] + o["dram_write"] for o in lst) / n,
                          tensor_pct=sum(o["tensor_pct"] for o in lst) / n, dram_pct=sum(o["dram_pct"] for o in lst) / n)
    gemm = [o for o in out if o["kernel"].startswith("gemm_tc_kernel")]
    res = dict(source=md_out, gemm_launches_captured=len(gemm),
               gemm_traffic_bytes_per_launch=sum(o["dram_read"] + o["dram_write"] for o in gemm) / max(len(gemm), 1),
               per_kernel=summary)
    # per bench.py epilogue class (what `roofline.traffic` / per_class.traffic_bytes_per_launch report), with the
    # algorithmic bytes of the same launch for workload C3 (bench.gemm_algorithmic_bytes)
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    classes = {"fwd_lrt": ("0", "7"), "fwd": ("1",), "dx": ("3",), "dx_lrt": ("4",), "dw": ("5",), "dw_lrt": ("6",)}
    w = bench.WORKLOADS["c3"]
    per_class = {}
    for cls, modes in classes.items():
        sel = [o for o in gemm if re.match(r"gemm_tc_kernel<(\d+),", o["kernel"]).group(1) in modes]
        if not sel:
            continue
        per_class[cls] = dict(dram_bytes_per_launch=sum(o["dram_read"] + o["dram_write"] for o in sel) / len(sel),
                              algorithmic_bytes_per_launch=bench.gemm_algorithmic_bytes(cls, w, w["N"]) if cls.endswith("lrt") else None,
                              tensor_pipe_pct=sum(o["tensor_pct"] * o["us"] for o in sel) / sum(o["us"] for o in sel),
                              launches_captured=len(sel))
    res["per_class"] = per_class
    json.dump(res, open(json_out, "w"), indent=1)
    with open(md_out, "w") as f:
        f.write(f"# {title}\n\n")
        f.write("Template args of gemm_tc_kernel: <epilogue class, BLOCK_N, CTA-group size>; classes 0 store (first half of the "
                "split LRT forward), 1 fwd, 2 fwd-lrt (dual), 3 dx, 4 dx-lrt, 5 dw, 6 dw-lrt, 7 fwd-lrt second half "
                "(variance GEMM + join).  Per-launch times are cold-cache and serialised under ncu.\n\n")
        f.write("| kernel | time us | tensor pipe % | DRAM read MB | DRAM write MB | DRAM % | L2 hit % | regs |\n|---|---:|---:|---:|---:|---:|---:|---:|\n")
        for o in out:
            f.write(f"| `{o['kernel']}` | {o['us']:.1f} | {o['tensor_pct']:.1f} | {o['dram_read'] / 1e6:.1f} | {o['dram_write'] / 1e6:.1f} | "
                    f"{o['dram_pct']:.1f} | {o['l2_hit']:.1f} | {o['regs']} |\n")
    print(json.dumps(res, indent=1)[:2500])


if __name__ == "__main__":
    main(*sys.argv[1:5])
