"""BASELINE configs[4]: the data-parallel scaling sweep -- VB-MLP hidden width H in {1024, 2048, 4096, 8192} (C3's
topology: input = H, 4 hidden layers, 1000 classes, local reparameterisation, bf16 GEMMs), GLOBAL batch in
{4k, 16k, 64k}, on G GPUs (strong scaling inside each (H, batch) cell: the global batch is split over the ranks).

  python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/c5_sweep.py [--steps K]
  python tools/c5_sweep.py                      # G = 1: the per-cell single-GPU reference the efficiencies divide by

Rank 0 prints one JSON line per cell: samples/s (device time, max over ranks, auxiliary streams joined), ms per
minibatch, GEMM TFLOP/s per epilogue class and the phase marks -- `wait_params` is what the exchange exposes on the main
stream (the next forward waiting for the owners' refreshed operands)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--widths", default="1024,2048,4096,8192")
    ap.add_argument("--batches", default="4096,16384,65536")
    args = ap.parse_args()
    env = bench.Env(args)
    peaks = bench.load_peaks()
    for H in [int(x) for x in args.widths.split(",")]:
        for Ng in [int(x) for x in args.batches.split(",")]:
            N = Ng // env.world
            w = dict(sizes=[H, H, H, H, H, 1000], N=N, S=1, reparam="local", precision="bf16", B=100.0,
                     l2="inputs alternate between two minibatches", desc=f"C5 cell H={H}, global batch {Ng}")
            # activations: ~5 bf16 tensors + one fp32 per layer; skip cells that cannot fit next to the parameters
            need = 5 * (5 * 2 + 4) * N * H + 5 * 40 * H * H
            if need > 150e9:
                if env.rank == 0:
                    print(json.dumps(dict(H=H, global_batch=Ng, n_gpus=env.world, skipped="does not fit 180 GB")), flush=True)
                continue
            m = bench.measure(env, args, w, N, args.steps, args.warmup, None, e2e=False)
            if env.rank != 0:
                continue
            ro = bench.roofline_objects(m, w, N, args.steps, peaks)
            fps = bench.flops_per_sample(w)
            line = dict(H=H, global_batch=Ng, per_gpu_batch=N, n_gpus=env.world, dp_exchange=env.dp_mode,
                        value=m["value"], unit="samples/s", ms_per_step=m["ms_per_step"],
                        step_tflops_per_gpu=fps * N / (m["ms_per_step"] / 1e3) / 1e12,
                        gemm_tflops=ro.get("roofline_all_gemms", {}).get("achieved"),
                        per_class={k: round(v["tflops"], 1) for k, v in ro.get("roofline_all_gemms", {}).get("per_class", {}).items()},
                        phases_ms_per_step=ro.get("phases_ms_per_step"))
            print(json.dumps(line), flush=True)
    if env.dist is not None:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
