"""Generates the golden fixtures of tests/golden/*.npz from the CPU oracle (fp64).

The reference itself cannot run in this image (no Torch7/LuaJIT) and ships
no vectors of its own, so these fixtures pin the ORACLE's restatement of VBLinear.lua / mlp.lua /
main.lua:19-51 on seeded inputs.  They travel to the GPU box (where /root/reference and this
script's fp64 run are not needed) and are checked twice: against the oracle on the CPU
(tests/test_golden_cpu.py) and against the CUDA path through the C ABI (tests/test_gpu_*.py).

Run from the repo root:  python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vbnn_oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
T64 = torch.float64


def t2n(t):
    return t.detach().cpu().numpy().astype(np.float64)


def layer_case(reparam, I=20, Ol=12, N=16, S=3, seed=11):
    rng = np.random.RandomState(seed)
    opt = O.default_opt(B=40.0, S=S, mu_init=1, var_init=0.01, reparam=reparam)
    lyr = O.VBLinearOracle(I, Ol, opt, T64, np.random.RandomState(seed + 1))
    lyr.means.copy_(torch.from_numpy(rng.randn(Ol, I) * math.sqrt(2.0 / I)))
    lyr.lvars.copy_(torch.from_numpy(rng.uniform(math.log(1e-4), math.log(1e-2), (Ol, I))))
    lyr.bias.copy_(torch.from_numpy(rng.randn(Ol) * 0.1))
    lyr.compute_prior()
    out = dict(I=I, O=Ol, N=N, S=S, B=opt["B"], means0=t2n(lyr.means), lvars0=t2n(lyr.lvars),
               bias0=t2n(lyr.bias), var_hat0=lyr.var_hat)
    lyr.resetAcc(); lyr.gradWeight.zero_(); lyr.gradBias.zero_()
    X = torch.from_numpy(rng.randn(N, I))
    out["X"] = t2n(X)
    for s in range(S):
        G = torch.from_numpy(rng.randn(N, Ol) / N)
        if reparam == "local":
            noise = torch.from_numpy(rng.randn(N, Ol))
            Y = lyr.updateOutput(X, noise)
        else:
            noise = torch.from_numpy(rng.randn(Ol, I))
            lyr.sample(noise)
            Y = lyr.updateOutput(X)
        dX = lyr.updateGradInput(X, G)
        lyr.accGradParameters(X, G, 1.0)
        out[f"G{s}"], out[f"noise{s}"], out[f"Y{s}"], out[f"dX{s}"] = t2n(G), t2n(noise), t2n(Y), t2n(dX)
    out["gradWeight"], out["gradSum"], out["gradBias"] = t2n(lyr.gradWeight), t2n(lyr.gradSum), t2n(lyr.gradBias)
    stats = lyr.update(opt)
    out["means1"], out["lvars1"], out["bias1"] = t2n(lyr.means), t2n(lyr.lvars), t2n(lyr.bias)
    out["var_hat1"] = lyr.var_hat
    out["stat_names"] = np.array(list(stats.keys()))
    out["stat_values"] = np.array(list(stats.values()), dtype=np.float64)
    out["lc_sum_cached"] = float(lyr.calc_lc(opt).sum())      # quirk Q6: pre-step tensors
    # a second update (Adam t=2) with fresh accumulators
    lyr.resetAcc(); lyr.gradWeight.zero_(); lyr.gradBias.zero_()
    if reparam == "local":
        lyr.updateOutput(X, torch.from_numpy(t2n(torch.zeros(N, Ol)) + out["noise0"]))
    else:
        lyr.sample(torch.from_numpy(out["noise0"]))
        out["W_step2"] = t2n(lyr.weight)                        # quirk Q1: mu_new + sigma_old*eps
        lyr.updateOutput(X)
    lyr.accGradParameters(X, torch.from_numpy(out["G0"]), 1.0)
    lyr.update(opt)
    out["means2"], out["lvars2"] = t2n(lyr.means), t2n(lyr.lvars)
    return out


def mlp_case(seed=21, steps=2):
    rng = np.random.RandomState(seed)
    opt = O.default_opt(input_size=24, hidden=[16, 12], classes=list("abcde"), S=2, B=25.0,
                        batchSize=8, mu_init=1, var_init=0.01)
    net = O.MLPOracle(opt, T64, seed=3)
    out = dict(sizes=np.array([24, 16, 12, 5]), S=2, B=25.0, N=8, steps=steps)
    for k, lyr in enumerate(net.vb):
        lyr.means.copy_(torch.from_numpy(rng.randn(*lyr.means.shape) * math.sqrt(2.0 / lyr.I)))
        lyr.lvars.copy_(torch.from_numpy(rng.uniform(math.log(1e-4), math.log(1e-2), tuple(lyr.lvars.shape))))
        lyr.compute_prior()
        out[f"means0_{k}"], out[f"lvars0_{k}"] = t2n(lyr.means), t2n(lyr.lvars)
    out["wout0"], out["bout0"] = t2n(net.out.weight), t2n(net.out.bias)
    for it in range(steps):
        X = torch.from_numpy(rng.randn(8, 24))
        T = torch.from_numpy(rng.randint(1, 6, 8).astype(np.float64))
        eps = [[torch.from_numpy(rng.randn(l.O, l.I)) for l in net.vb] for _ in range(opt["S"])]
        err, acc = O.train_minibatch(net, X, T, opt, eps=eps)
        out[f"X_{it}"], out[f"T_{it}"] = t2n(X), t2n(T)
        for s in range(opt["S"]):
            for k in range(len(net.vb)):
                out[f"eps_{it}_{s}_{k}"] = t2n(eps[s][k])
        out[f"err_{it}"], out[f"acc_{it}"] = err, acc
    for k, lyr in enumerate(net.vb):
        out[f"means1_{k}"], out[f"lvars1_{k}"], out[f"bias1_{k}"] = t2n(lyr.means), t2n(lyr.lvars), t2n(lyr.bias)
    out["wout1"], out["bout1"] = t2n(net.out.weight), t2n(net.out.bias)
    out["lc"] = net.calc_lc(opt)
    return out


def write_replay_bin():
    """mlp_weight.npz as a flat little-endian file the C++ replay host (tools/replay.cu) can read without a zip /
    npy parser: "VBRP", u32 count, then per array: u32 name_len, name, u32 ndim, u32 dims[ndim], f32 data."""
    import struct
    g = np.load(os.path.join(HERE, "mlp_weight.npz"))
    with open(os.path.join(HERE, "mlp_weight.replay.bin"), "wb") as f:
        f.write(b"VBRP" + struct.pack("<I", len(g.files)))
        for k in g.files:
            a = np.atleast_1d(np.asarray(g[k], dtype=np.float64)).astype("<f4")
            f.write(struct.pack("<I", len(k)) + k.encode() + struct.pack("<I", a.ndim))
            f.write(struct.pack(f"<{a.ndim}I", *a.shape))
            f.write(a.tobytes())


if __name__ == "__main__":
    if "--replay-only" in sys.argv:
        write_replay_bin()
        sys.exit(0)
    np.savez_compressed(os.path.join(HERE, "vblinear_weight.npz"), **layer_case("weight"))
    np.savez_compressed(os.path.join(HERE, "vblinear_local.npz"), **layer_case("local"))
    np.savez_compressed(os.path.join(HERE, "mlp_weight.npz"), **mlp_case())
    write_replay_bin()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz") or f.endswith(".bin"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
