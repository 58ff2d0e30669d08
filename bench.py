#!/usr/bin/env python
"""bench.py -- VB-MLP train samples/sec on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--workload c3|c2|c1] [--impl reference]

One "step" = one minibatch of main.lua:28-40: S x (sample, forward, loss, backward), the gradient
allreduce when N > 1, and the fused KL + Adam update.  Default workload (N=1): BASELINE configs[2],
the wide VB-MLP 4096-4096x4-1000, batch 8192 per GPU, local reparameterisation, bf16 GEMMs with
fp32 accumulate -- the config the metric's "tensor-pipe % of peak" is quoted on.  Weak scaling:
per-GPU batch fixed, rows of the global minibatch sharded over ranks, NCCL sum-allreduce of the
{gradWeight, gradSum, gradBias} arena.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput over exactly K timed
minibatches; `e2e` is the same metric through the public host-buffer API (H2D of every minibatch +
D2H of its loss inside the timed region); `roofline` / `roofline_hbm` / `phases_ms_per_step` come
from CUDA events around every tensor-core GEMM / fused-update launch and phase marks inside the
minibatch during an instrumented repeat of the same K minibatches (instrumenting disables graph
replay, so `value` is timed without it); `cpu_baseline` is the oracle ("port" of the reference's Torch7 CPU path, which cannot run
here) on a bounded sample.  The oracle is only ever the checker / CPU baseline, never the product.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: sizes, per-GPU batch, S, reparam, precision, vb_output, B (trainSize/N)
    "c1": dict(sizes=[784, 100, 10], N=100, S=1, reparam="weight", precision="fp32", B=600.0,
               desc="mlp.lua VBLinear MLP 784-100-10, batch 100, 1 weight sample (BASELINE configs[0])"),
    "c2": dict(sizes=[784, 1200, 1200, 10], N=1024, S=10, reparam="weight", precision="bf16", B=58.59,
               desc="VBLinear MLP 784-1200-1200-10, batch 1024, 10 MC weight samples (BASELINE configs[1])"),
    "c3": dict(sizes=[4096, 4096, 4096, 4096, 4096, 1000], N=8192, S=1, reparam="local", precision="bf16",
               B=100.0,
               desc="wide VB-MLP 4096-4096x4-1000, local reparameterization, batch 8192/GPU, bf16 GEMMs fp32 "
                    "accumulate (BASELINE configs[2])"),
}


def flops_per_sample(w):
    """weight sampling c=4 (first layer) / 6; LRT c=8 / 12; plain nn.Linear output c=6."""
    s = w["sizes"]
    lrt = w["reparam"] == "local"
    total = 0
    for j in range(len(s) - 1):
        last = j == len(s) - 2
        if last:
            c = 6
        elif lrt:
            c = 8 if j == 0 else 12
        else:
            c = 4 if j == 0 else 6
        total += c * s[j] * s[j + 1]
    return w["S"] * total


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d["bf16_tflops_sustained"]), tflops_burst=float(d["bf16_tflops"]),
                    hbm=float(d["hbm_gbs"]), source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def summary(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons, pw = [], 0.0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                c, m = float(f[0]), float(f[1])
            except ValueError:
                continue
            mx = max(mx, m)
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(c)
                try:
                    pw.append(float(f[2]))
                except ValueError:
                    pass
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        if not sm:
            sm = [float(l.split(",")[0]) for _, l in self.rows[-3:] if l and l.split(",")[0].strip().replace(".", "").isdigit()] or [0.0]
        return dict(sm_mhz=statistics.median(sm), sm_min_mhz=min(sm), sm_max_mhz=mx, reasons=sorted(reasons),
                    samples=len(sm), power_w_max=max(pw) if pw else None)


def bind_to_gpu_numa_node(index):
    """Host side of the end-to-end path: run this rank's host thread (and so first-touch its pinned staging
    buffers) on the NUMA node the GPU hangs off, so that 8 ranks x 134 MB per minibatch do not cross the
    socket interconnect.  Best effort; returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, z = part.partition("-")
            cpus.update(range(int(a), int(z or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:                                            # noqa: BLE001
        pass
    return None


def build_net(w, ctx, N):
    import vbnn_b200
    s = w["sizes"]
    opt = vbnn_b200.default_opt(input_size=s[0], hidden=s[1:-1], classes=[str(i) for i in range(s[-1])],
                                S=w["S"], B=w["B"], batchSize=N, testBatchSize=N, mu_init=1, msr_init=False,
                                var_init=0.001,
                                reparam=w["reparam"], precision=w["precision"], strict_reference=False, log=False,
                                seed=5)
    net = vbnn_b200.MLP(opt, ctx, max_batch=N)
    net.init_params(seed=4, he_means=True)
    return net, opt


def cpu_baseline(w, threads, budget_rows=None):
    """The 'reference Torch7 CPU path' as restated by the oracle, fp32, `threads` host threads, on a
    bounded sample: the full-size network, a row subsample of the minibatch through S x (sample,
    forward, backward) plus ONE full-size update; extrapolated to the full batch."""
    import numpy as np
    import torch
    from oracle import vbnn_oracle as O
    torch.set_num_threads(threads)
    s = w["sizes"]
    opt = O.default_opt(input_size=s[0], hidden=s[1:-1], classes=[str(i) for i in range(s[-1])], S=w["S"],
                        B=w["B"], batchSize=w["N"], mu_init=1, msr_init=False, var_init=0.001, reparam=w["reparam"],
                        strict_reference=False)
    net = O.MLPOracle(opt, torch.float32, seed=3)
    rows = budget_rows or max(16, min(w["N"], int(2.0e11 / max(flops_per_sample(w), 1))))
    rows = min(rows, w["N"])
    rng = np.random.RandomState(0)
    X = torch.from_numpy(rng.randn(rows, s[0]).astype(np.float32))
    T = torch.from_numpy(rng.randint(1, s[-1] + 1, rows).astype(np.float32))
    net.resetGradients()
    t0 = time.perf_counter()
    for _ in range(opt["S"]):
        net.sample()                                   # host randomkit-style Gaussian fill, as the reference
        net.run(X, T)
    t_fb = time.perf_counter() - t0
    t0 = time.perf_counter()
    net.update(opt)
    t_up = time.perf_counter() - t0
    full = t_fb * (w["N"] / rows) + t_up
    return dict(value=w["N"] / full, unit="samples/s", cores=threads, kind="port",
                sample=f"{rows} of {w['N']} rows x S={w['S']} through the full-size net ({t_fb:.2f} s) + one full update "
                       f"({t_up:.2f} s), fp32 torch-CPU oracle, extrapolated to the full minibatch",
                host_cpus=os.cpu_count())


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 8
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(w, threads)
        if i >= args.warmup:
            vals.append(r)
    v = statistics.mean(x["value"] for x in vals)
    r = vals[-1]
    r["value"] = v
    line = dict(metric="VB-MLP train samples/sec", value=v, unit="samples/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * w["N"] / v, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=w["desc"], global_batch=w["N"], S=w["S"], reparam=w["reparam"]),
                cpu_baseline=r, e2e=dict(value=v, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="vbnn", choices=["vbnn", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.steps is None:
        args.steps = 3 if args.impl == "reference" else (200 if args.workload == "c3" else 500)
    if args.warmup is None:
        args.warmup = 1 if args.impl == "reference" else (10 if args.workload == "c3" else 20)
    if args.impl == "reference":
        return run_reference(args, w)
    args.warmup = max(args.warmup, 3)

    import torch
    import vbnn_b200
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local_rank}"))
    numa = bind_to_gpu_numa_node(local_rank)
    ctx = vbnn_b200.Context(local_rank, seed=5)
    if world > 1:
        def bcast(buf):
            t = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local_rank}")
            if rank == 0:
                t.copy_(torch.tensor(list(buf), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        ctx.init_comm(rank, world, bcast)

    N = w["N"]
    net, opt = build_net(w, ctx, N)
    dp_mode = "single"
    if world > 1:
        dp_mode = os.environ.get("VBNN_DP", "peer")
        if dp_mode == "peer":
            def gather_bytes(blob):
                t = torch.tensor(list(blob), dtype=torch.uint8, device=f"cuda:{local_rank}")
                out = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(out, t)
                return [bytes(o.cpu().tolist()) for o in out]
            # every rank must agree: if CUDA IPC is unavailable anywhere, all fall back to the NCCL exchange
            ok = 1
            try:
                net.enable_peer(gather_bytes)
            except Exception as e:                                   # noqa: BLE001
                ok = 0
                print(f"[bench rank {rank}] peer mode unavailable ({e}); falling back to NCCL allreduce", file=sys.stderr)
            flag = torch.tensor([ok], device=f"cuda:{local_rank}")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag[0]) == 0:
                if net.peer_active:
                    raise SystemExit("peer mode is active on this rank but not on all ranks")
                dp_mode = "nccl (peer mode unavailable)"
    g = torch.Generator(device="cpu").manual_seed(3 + rank)
    nbuf = 2
    Xd = [torch.randn(N, w["sizes"][0], generator=g).cuda() for _ in range(nbuf)]
    Td = [torch.randint(1, w["sizes"][-1] + 1, (N,), generator=g).float().cuda() for _ in range(nbuf)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_region():
        """EXACTLY K minibatches between barrier + synchronize on both sides; device time, max over ranks."""
        barrier()
        t0 = sampler.mark() if sampler else 0
        ev0.record()
        for i in range(args.steps):
            net.train_step(Xd[i % nbuf], Td[i % nbuf], sync=False)
        ev1.record()
        barrier()
        t1 = sampler.mark() if sampler else 0
        t_ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([t_ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t[0])
        return t_ms, t0, t1

    # ---------------- device-resident throughput (value): the path a user gets -------------------
    for i in range(args.warmup):
        net.train_step(Xd[i % nbuf], Td[i % nbuf], sync=False)
    l0 = net.launch_count()
    ms, t_mark0, t_mark1 = timed_region()
    launches = net.launch_count() - l0
    value = args.steps * N * world / (ms / 1e3)
    # ---------------- the same K minibatches again with CUDA events around every GEMM / update launch
    # and phase marks inside the minibatch (roofline, phases).  Instrumenting turns graph replay off and
    # adds two event records per launch, so it is a separate pass: `value` carries no instrumentation.
    ms_prof, prof, phases = None, {}, {}
    if w["precision"] == "bf16":
        ctx.profile(True)
        for i in range(3):
            net.train_step(Xd[i % nbuf], Td[i % nbuf], sync=False)
        ctx.profile(False); ctx.profile(True)                       # drop the warm-up records
        ms_prof, _, _ = timed_region()
        prof = ctx.profile_read()
        phases = ctx.phase_read()
        ctx.profile(False)
    err_last = float(net._res.cpu()[0])

    # ---------------- end to end through the host-buffer API (e2e) ------------------------------
    e2e = None
    if not args.no_e2e:
        Xh = [x.cpu().pin_memory() for x in Xd]
        Th = [t.cpu().pin_memory() for t in Td]
        for i in range(3):
            net.submit_host(Xh[i % nbuf], Th[i % nbuf]); net.collect()
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        net.submit_host(Xh[0], Th[0])
        for i in range(1, args.steps):
            net.submit_host(Xh[i % nbuf], Th[i % nbuf])        # copy of minibatch i overlaps compute of i-1
            net.collect()
        net.collect()
        ev1.record()
        barrier()
        ms_e = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms_e], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t[0])
        e2e = dict(value=args.steps * N * world / (ms_e / 1e3), unit="samples/s",
                   h2d_bytes_per_step=N * w["sizes"][0] * 4 + N * 4, d2h_bytes_per_step=8,
                   api="vbnn_mlp_submit_host/vbnn_mlp_collect (pinned fp32 host minibatch, double-buffered)")

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    clocks = sampler.summary(t_mark0, t_mark1)
    peaks = load_peaks()
    fps = flops_per_sample(w)
    roof = None
    upd = prof.pop("update", None)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01e_c3_traffic.json")
    if args.workload == "c3" and os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = dict(gemm_bytes_per_launch=tj["gemm_traffic_bytes_per_launch"], source=tj["source"])
    if prof:
        tot_ms = sum(v[0] for v in prof.values())
        tot_fl = sum(v[2] for v in prof.values())
        tot_n = sum(v[1] for v in prof.values())
        achieved = tot_fl / (tot_ms / 1e3) / 1e12
        roof = dict(bound="tensor", kernel="gemm_tc_kernel (tcgen05 bf16, all epilogue classes)",
                    achieved=achieved, peak=peaks["tflops"], unit="TFLOP/s", frac=achieved / peaks["tflops"],
                    peak_source=peaks["source"], peak_burst=peaks["tflops_burst"],
                    traffic=traffic["gemm_bytes_per_launch"] if traffic else None,
                    traffic_source=traffic["source"] if traffic else None,
                    launches=tot_n, avg_launch_ms=tot_ms / max(tot_n, 1), share_of_step=tot_ms / ms_prof,
                    timed="CUDA events around every launch during an instrumented repeat of the K timed minibatches "
                          "(graph replay off); `value` is the un-instrumented pass",
                    instrumented_ms_per_step=ms_prof / args.steps,
                    per_class={k: dict(tflops=v[2] / (v[0] / 1e3) / 1e12, ms_per_step=v[0] / args.steps, launches=v[1])
                               for k, v in prof.items()})
    line = dict(metric="VB-MLP train samples/sec", value=value, unit="samples/s", n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16" if w["precision"] == "bf16" else "f32", data="synthetic",
                config=dict(workload=w["desc"], sizes=w["sizes"], global_batch=N * world, per_gpu_batch=N, S=w["S"],
                            reparam=w["reparam"], parallelism=f"dp{world}", dp_exchange=dp_mode, host_numa_node=numa, output_layer="nn.Linear (mlp.lua:29)",
                            l2="working set (parameters + optimizer state + activations) >> 126 MB L2; two "
                               "alternating input minibatches",
                            flops_per_sample=fps),
                step_tflops=fps * N / (ms / args.steps / 1e3) / 1e12,
                step_frac_of_peak=fps * N / (ms / args.steps / 1e3) / 1e12 / peaks["tflops"],
                gpu_launches=launches, clocks=clocks, last_error=err_last,
                phases_ms_per_step={k: v[0] / args.steps for k, v in phases.items()})
    if e2e:
        line["e2e"] = e2e
    if roof:
        line["roofline"] = roof
    if upd:
        gbs = upd[2] / (upd[0] / 1e3) / 1e9
        line["roofline_hbm"] = dict(bound="hbm", kernel="k_update (fused KL + 2x Adam, 56 B/weight algorithmic)", achieved=gbs,
                                    peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"], launches=upd[1],
                                    avg_launch_ms=upd[0] / max(upd[1], 1), share_of_step=upd[0] / ms_prof,
                                    peak_source=peaks["source"])
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(w, min(os.cpu_count() or 8, 64))
    elif not args.no_cpu_baseline:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
