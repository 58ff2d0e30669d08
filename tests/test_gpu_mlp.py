"""GPU parity tests of the mlp.lua / main.lua:28-40 path through the C ABI (pytest -m gpu).

The fused minibatch draws its noise on the device (Philox); the tests dump exactly that noise
through vbnn_layer_draw_noise and inject it into the CPU oracle, so both sides see identical
epsilon / zeta.  Tolerances (relative Frobenius error per tensor):
  * fp32 mode vs the fp64 oracle: <= 1e-5 on accumulators, <= 2e-4 after optimiser steps;
  * bf16-operand mode vs the oracle restating the same bf16 operand rounding
    (operand_round=round_bf16, accumulation in fp64): <= 5e-3 -- the residual is one-ulp bf16
    rounding flips where the fp32 and fp64 pre-rounding values straddle a tie;
  * bf16-operand mode vs the un-rounded fp64 oracle: <= 0.12, stated separately -- this is the
    inherent cost of bf16 operands through three layers (measured 2.5-6 % on first-layer
    gradients, identical in a pure-torch emulation; see tools/diag_bf16.py)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import vbnn_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def cpu(t):
    return t.detach().float().cpu().numpy()


def build_pair(ctx, sizes, N, S, B, precision, reparam, seed=0, strict=True, vb_output=False, lr_mu=None,
               operand_round=None, dtype=torch.float64):
    import vbnn_b200
    from vbnn_b200 import _lib as L
    over = dict(input_size=sizes[0], hidden=list(sizes[1:-1]), classes=[str(i) for i in range(sizes[-1])],
                S=S, B=B, batchSize=N, testBatchSize=N, mu_init=1, var_init=0.01, reparam=reparam,
                strict_reference=strict, vb_output=vb_output, log=False)
    if lr_mu:
        over["meanState"] = dict(learningRate=lr_mu)
    gopt = vbnn_b200.default_opt(precision=precision, **over)
    net = vbnn_b200.MLP(gopt, ctx, max_batch=N)
    oopt = O.default_opt(**over)
    ref = O.MLPOracle(oopt, dtype, seed=3, operand_round=operand_round)
    rng = np.random.RandomState(seed)
    layers = ref.vb + [ref.out]
    for k, (gl, ol) in enumerate(zip(net.model, layers)):
        if isinstance(ol, O.VBLinearOracle):
            mu = rng.randn(ol.O, ol.I) * math.sqrt(2.0 / ol.I)
            lv = rng.uniform(math.log(1e-4), math.log(1e-2), (ol.O, ol.I))
            ol.means.copy_(torch.from_numpy(mu)); ol.lvars.copy_(torch.from_numpy(lv))
            ol.compute_prior()
            gl.set(L.BUF_MEANS, mu); gl.set(L.BUF_LVARS, lv)
            gl.compute_prior()
        else:
            w = rng.randn(ol.weight.shape[0], ol.weight.shape[1]) * math.sqrt(2.0 / ol.weight.shape[1])
            ol.weight.copy_(torch.from_numpy(w)); ol.bias.zero_()
            gl.set(L.BUF_WEIGHT, w)
    return net, ref, gopt, oopt


def oracle_step(ref, oopt, X, T, eps=None, zeta=None):
    """train_minibatch with the raw accumulators captured before update() rescales them in place."""
    ref.resetGradients()
    serr = sacc = 0.0
    for s in range(oopt["S"]):
        ref.sample(None if eps is None else eps[s])
        e, a = ref.run(X, T, None if zeta is None else zeta[s])
        serr += e; sacc += a
    acc = [(l.gradWeight.clone(), l.gradSum.clone(), l.gradBias.clone()) for l in ref.vb_all]
    ref.update(oopt)
    return serr / oopt["S"], sacc / oopt["S"], acc


def test_mlp_piecewise_against_golden(ctx):
    """resetGradients / sample / run / update with the fixture's epsilon injected (fp32 mode)."""
    import vbnn_b200
    from vbnn_b200 import _lib as L
    g = np.load(os.path.join(GOLD, "mlp_weight.npz"))
    sizes = [int(v) for v in g["sizes"]]
    opt = vbnn_b200.default_opt(input_size=sizes[0], hidden=sizes[1:-1], classes=list("abcde"), S=int(g["S"]),
                                B=float(g["B"]), batchSize=int(g["N"]), mu_init=1, var_init=0.01, log=False)
    net = vbnn_b200.MLP(opt, ctx, max_batch=int(g["N"]))
    for k in range(2):
        net.model[k].set(L.BUF_MEANS, g[f"means0_{k}"]); net.model[k].set(L.BUF_LVARS, g[f"lvars0_{k}"])
        net.model[k].compute_prior()
    net.model[2].set(L.BUF_WEIGHT, g["wout0"]); net.model[2].set(L.BUF_BIAS, g["bout0"])
    for it in range(int(g["steps"])):
        X = torch.from_numpy(g[f"X_{it}"]).float().cuda()
        T = torch.from_numpy(g[f"T_{it}"]).float().cuda()
        net.resetGradients()
        serr = sacc = 0.0
        for s in range(opt["S"]):
            for k in range(2):
                net.model[k].sample(eps=torch.from_numpy(g[f"eps_{it}_{s}_{k}"]).float().cuda(), sample_idx=s)
            err, acc = net.run(X, T)
            serr += err; sacc += acc
        net.update(opt)
        assert abs(serr / opt["S"] - float(g[f"err_{it}"])) < 1e-4 * abs(float(g[f"err_{it}"]))
        assert abs(sacc / opt["S"] - float(g[f"acc_{it}"])) < 1e-3
    for k in range(2):
        assert rel(cpu(net.model[k].means), g[f"means1_{k}"]) < 2e-4
        assert rel(cpu(net.model[k].lvars), g[f"lvars1_{k}"]) < 2e-4
        assert rel(cpu(net.model[k].bias), g[f"bias1_{k}"]) < 1e-4
    assert rel(cpu(net.model[2].weight), g["wout1"]) < 1e-5
    assert rel(cpu(net.model[2].bias), g["bout1"]) < 1e-4
    assert abs(net.calc_lc() - float(g["lc"])) < 1e-3 * abs(float(g["lc"]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("reparam,vb_output", [("weight", False), ("local", False), ("local", True), ("weight", True)])
def test_fused_step_vs_oracle(ctx, precision, reparam, vb_output):
    """vbnn_mlp_step (S samples batched per launch; CUDA graph from the 2nd minibatch on) against
    the oracle fed with the device-drawn noise."""
    sizes, N, S = [40, 48, 36, 6], 24, 3
    bf = precision == "bf16"
    net, ref, gopt, oopt = build_pair(ctx, sizes, N, S, 30.0, precision, reparam, vb_output=vb_output,
                                      operand_round=O.round_bf16 if bf else None)
    ref64 = build_pair(ctx, sizes, N, S, 30.0, precision, reparam, vb_output=vb_output)[1] if bf else None
    rng = np.random.RandomState(5)
    tol_acc = 5e-3 if bf else 1e-5
    tol_par = 5e-3 if bf else 2e-4
    nvb = len(ref.vb_all)
    gidx = lambda k: k if k < len(ref.vb) else len(net.model) - 1
    for it in range(4):
        Xn = rng.randn(N, sizes[0]); Tn = rng.randint(1, sizes[-1] + 1, N).astype(np.float64)
        step = ctx.get_step()
        noise = [[torch.from_numpy(cpu(net.model[gidx(k)].draw_noise(step, s, rows=N)).astype(np.float64))
                  for k in range(nvb)] for s in range(S)]
        err, acc = net.train_step(torch.from_numpy(Xn).float().cuda(), torch.from_numpy(Tn).float().cuda())
        kw = dict(zeta=noise) if reparam == "local" else dict(eps=noise)
        rerr, racc, accs = oracle_step(ref, oopt, torch.from_numpy(Xn), torch.from_numpy(Tn), **kw)
        assert ctx.get_step() == step + 1
        assert abs(err - rerr) < (2e-3 if bf else 1e-4) * abs(rerr), (it, err, rerr)
        if not bf:
            assert abs(acc - racc) < 1e-3
        for k, ol in enumerate(ref.vb_all):
            gl = net.model[gidx(k)]
            gw, gs, gb = accs[k]
            assert rel(cpu(gl.gradWeight), gw.numpy()) < tol_acc, (it, k, "gW")
            assert rel(cpu(gl.gradSum), gs.numpy()) < tol_acc * 2, (it, k, "gS")
            assert rel(cpu(gl.gradBias), gb.numpy()) < tol_acc, (it, k, "gb")
            assert rel(cpu(gl.means), ol.means.numpy()) < tol_par, (it, k, "means")
            assert rel(cpu(gl.lvars), ol.lvars.numpy()) < tol_par, (it, k, "lvars")
            assert rel(cpu(gl.bias), ol.bias.numpy()) < max(tol_par, 1e-4), (it, k, "bias")
        if not vb_output:
            assert rel(cpu(net.model[-1].weight), ref.out.weight.numpy()) < tol_par
        if bf and it == 0:
            # the bf16-vs-fp64 gap, stated separately (inherent to bf16 operands)
            e64, _, accs64 = oracle_step(ref64, oopt, torch.from_numpy(Xn), torch.from_numpy(Tn), **kw)
            assert abs(err - e64) < 2e-2 * abs(e64)
            for k in range(nvb):
                assert rel(cpu(net.model[gidx(k)].gradWeight), accs64[k][0].numpy()) < 0.12, (k, "gW vs fp64")


def test_graph_replay_equals_eager(ctx):
    rng = np.random.RandomState(1)
    X = torch.from_numpy(rng.randn(32, 40)).float().cuda()
    T = torch.from_numpy(rng.randint(1, 7, 32).astype(np.float32)).cuda()
    outs = []
    import vbnn_b200
    for no_graph in (1, 0):
        vbnn_b200.knob("no_graph", no_graph)
        ctx.set_step(100)
        net, _, _, _ = build_pair(ctx, [40, 48, 36, 6], 32, 2, 30.0, "fp32", "weight", seed=9)
        res = [net.train_step(X, T) for _ in range(4)]
        outs.append((res, cpu(net.model[0].means).copy(), cpu(net.model[1].lvars).copy()))
    vbnn_b200.knob("no_graph")
    (r0, m0, l0), (r1, m1, l1) = outs
    assert np.allclose(np.array(r0), np.array(r1), rtol=1e-5, atol=1e-6)
    assert rel(m1, m0) < 1e-6 and rel(l1, l0) < 1e-6


def test_step_host_equals_step_device(ctx):
    rng = np.random.RandomState(2)
    Xh = torch.from_numpy(rng.randn(16, 40)).float().pin_memory()
    Th = torch.from_numpy(rng.randint(1, 7, 16).astype(np.float32)).pin_memory()
    res = []
    for mode in ("dev", "host", "pipe"):
        ctx.set_step(7)
        net, _, _, _ = build_pair(ctx, [40, 48, 36, 6], 16, 2, 30.0, "fp32", "weight", seed=4)
        if mode == "dev":
            r = [net.train_step(Xh.cuda(), Th.cuda()) for _ in range(3)]
        elif mode == "host":
            r = [net.train_step_host(Xh, Th) for _ in range(3)]
        else:
            net.submit_host(Xh, Th); net.submit_host(Xh, Th)
            r = [net.collect()]
            net.submit_host(Xh, Th)
            r += [net.collect(), net.collect()]
        res.append((r, cpu(net.model[0].means).copy()))
    for r, m in res[1:]:
        assert np.allclose(np.array(r), np.array(res[0][0]), rtol=1e-5, atol=1e-6)
        assert rel(m, res[0][1]) < 1e-6


def test_mlp_test_map_and_sampled(ctx):
    sizes, N = [40, 48, 36, 6], 20
    net, ref, gopt, oopt = build_pair(ctx, sizes, N, 2, 30.0, "fp32", "weight", seed=3)
    rng = np.random.RandomState(8)
    Xn = rng.randn(N, 40); Tn = rng.randint(1, 7, N).astype(np.float64)
    X, T = torch.from_numpy(Xn).float().cuda(), torch.from_numpy(Tn).float().cuda()
    net.opt["quicktest"] = True; oopt["quicktest"] = True                 # mlp.lua:87-91
    e, a = net.test(X, T)
    re_, ra = ref.test(torch.from_numpy(Xn), torch.from_numpy(Tn))
    assert abs(e - re_) < 1e-4 * abs(re_) and abs(a - ra) < 1e-3
    net.opt["quicktest"] = False; oopt["quicktest"] = False               # mlp.lua:93-102
    net.opt["testSamples"] = oopt["testSamples"] = 5
    step = ctx.get_step()
    eps = [[torch.from_numpy(cpu(net.model[k].draw_noise(step, (1 << 20) + s)).astype(np.float64)) for k in range(2)]
           for s in range(5)]
    e, a = net.test(X, T)
    re_, ra = ref.test(torch.from_numpy(Xn), torch.from_numpy(Tn), eps_lists=eps)
    assert abs(e - re_) < 1e-4 * abs(re_) and abs(a - ra) < 1e-3


def test_c1_config_learns(ctx):
    """BASELINE configs[0] shape (784-100-10, batch 100, S=1): the loss must go down."""
    import vbnn_b200
    opt = vbnn_b200.default_opt(hidden=[100], S=1, B=600.0, batchSize=100, mu_init=1, msr_init=True, log=False,
                                meanState=dict(learningRate=0.003))
    net = vbnn_b200.MLP(opt, ctx, max_batch=100)
    net.init_params(seed=4, he_means=True)
    rng = np.random.RandomState(0)
    Xn = rng.randn(100, 784).astype(np.float32)
    Tn = ((Xn @ rng.randn(784, 10)).argmax(1) + 1).astype(np.float32)
    X, T = torch.from_numpy(Xn).cuda(), torch.from_numpy(Tn).cuda()
    first = net.train_step(X, T)[0]
    for _ in range(150):
        last = net.train_step(X, T)
    assert last[0] < 0.7 * first and last[1] > 50.0


@pytest.mark.parametrize("sizes,N,S,reparam", [([784, 1200, 1200, 10], 1024, 10, "weight"),
                                              ([1024, 1024, 1024, 1000], 2048, 1, "local")])
def test_full_size_bf16_tensor_core_path(ctx, sizes, N, S, reparam):
    """BASELINE-size minibatch (C2 exactly; a C3-shaped net) on the tcgen05 path, checked two ways:
    (1) against the oracle restating the bf16 operand rounding, fed the device-drawn noise
    (fp32 accumulate on the CPU to keep it to seconds): <= 5e-3;
    (2) against this library's own fp32 CUDA-core path (oracle-checked above) with identical Philox
    noise: the inherent bf16-operand gap, <= 0.12 on gradients and <= 2e-2 on the loss."""
    rng = np.random.RandomState(3)
    Xn = rng.randn(N, sizes[0]).astype(np.float32)
    Tn = rng.randint(1, sizes[-1] + 1, N).astype(np.float32)
    X, T = torch.from_numpy(Xn).cuda(), torch.from_numpy(Tn).cuda()
    out = {}
    for precision in ("fp32", "bf16"):
        ctx.set_step(11)
        bf = precision == "bf16"
        net, ref, _, oopt = build_pair(ctx, sizes, N, S, 58.6, precision, reparam, seed=6, strict=False,
                                       operand_round=O.round_bf16 if bf else None, dtype=torch.float32)
        if bf:
            nvb = len(ref.vb)
            noise = [[net.model[k].draw_noise(11, s, rows=N).cpu() for k in range(nvb)] for s in range(S)]
        err, acc = net.train_step(X, T)
        assert math.isfinite(err)
        out[precision] = (err, [cpu(m.gradWeight).copy() for m in net.model],
                          [cpu(m.gradSum).copy() for m in net.model[:-1]], [cpu(m.means).copy() for m in net.model[:-1]])
        if bf:
            kw = dict(zeta=noise) if reparam == "local" else dict(eps=noise)
            rerr, _, accs = oracle_step(ref, oopt, torch.from_numpy(Xn), torch.from_numpy(Tn), **kw)
            assert abs(err - rerr) < 2e-3 * abs(rerr)
            for k in range(nvb):
                assert rel(out["bf16"][1][k], accs[k][0].numpy()) < 5e-3, (k, "gW vs bf16 oracle")
                assert rel(out["bf16"][2][k], accs[k][1].numpy()) < 1e-2, (k, "gS vs bf16 oracle")
                assert rel(out["bf16"][3][k], ref.vb[k].means.numpy()) < 5e-3, (k, "means vs bf16 oracle")
        del net
    assert abs(out["bf16"][0] - out["fp32"][0]) < 2e-2 * abs(out["fp32"][0])
    for a, b in zip(out["bf16"][1], out["fp32"][1]):
        assert rel(a, b) < 0.12
    for a, b in zip(out["bf16"][2], out["fp32"][2]):
        assert rel(a, b) < 0.15
    for a, b in zip(out["bf16"][3], out["fp32"][3]):
        assert rel(a, b) < 2e-2


def test_checkpoint_resume_is_exact(ctx):
    """export after 2 minibatches, load into a fresh net, continue -- identical to the
    uninterrupted run (parameters, Adam state, step counters, quirk-Q1 sigma cache all restored)."""
    rng = np.random.RandomState(12)
    X = torch.from_numpy(rng.randn(32, 40)).float().cuda()
    T = torch.from_numpy(rng.randint(1, 7, 32).astype(np.float32)).cuda()
    ctx.set_step(0)
    a, _, _, _ = build_pair(ctx, [40, 48, 36, 6], 32, 2, 30.0, "fp32", "weight", seed=2)
    for _ in range(2):
        a.train_step(X, T)
    sd = a.state_dict()
    ra = [a.train_step(X, T) for _ in range(3)]
    ma = cpu(a.model[0].means).copy()
    b, _, _, _ = build_pair(ctx, [40, 48, 36, 6], 32, 2, 30.0, "fp32", "weight", seed=99)
    b.load_state_dict(sd)
    rb = [b.train_step(X, T) for _ in range(3)]
    assert np.allclose(np.array(ra), np.array(rb), rtol=1e-6, atol=1e-7)
    assert rel(cpu(b.model[0].means), ma) < 1e-7
    assert b.model[0].t == a.model[0].t == 5


def test_epoch_loops_device_resident(ctx):
    """main:train / main:test (main.lua:13-74) over a device-resident synthetic
    MNIST-shaped dataset; the fused loop and the step-by-step reference-order loop agree."""
    import vbnn_b200
    from vbnn_b200 import train as tr
    res = []
    for fused in (True, False):
        ctx.set_step(0)
        opt = vbnn_b200.default_opt(hidden=[32], S=2, B=8.0, batchSize=50, testBatchSize=100, trainSize=400,
                                    testSize=200, mu_init=1, msr_init=True, log=False, testSamples=3,
                                    meanState=dict(learningRate=0.002))
        net = vbnn_b200.MLP(opt, ctx, max_batch=100)
        net.init_params(seed=4, he_means=True)
        ds = tr.synthetic_dataset(400, 784, 10, seed=3, geometry=(28, 28))
        acc, err = tr.train(net, ds, opt, fused=fused, shuffle_seed=1)
        tacc, terr = tr.test(net, dict(inputs=ds["inputs"][:200], targets=ds["targets"][:200]), opt)
        assert math.isfinite(err) and 0.0 <= acc <= 100.0 and math.isfinite(terr) and 0.0 <= tacc <= 100.0
        res.append((acc, err, tacc, terr))
    assert np.allclose(res[0], res[1], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("sizes,N,S,reparam", [([13, 7, 2], 1, 1, "weight"),        # one row, odd widths, 2 classes
                                              ([13, 7, 2], 3, 3, "local"),
                                              ([784, 100, 10], 100, 1, "weight"),  # BASELINE C1 exactly
                                              ([33, 129, 65, 3], 130, 2, "weight"),  # every dim ragged vs 32/64/128 tiles
                                              ([33, 129, 65, 3], 257, 2, "local")])
def test_ragged_and_tiny_shapes_vs_oracle(ctx, precision, sizes, N, S, reparam):
    """Edge shapes (rows / features that are not multiples of the 8-element TMA pitch, the 32-row
    epilogue chunk or the 128 / 256 tile; a single-row minibatch; two classes) through vbnn_mlp_step."""
    bf = precision == "bf16"
    net, ref, gopt, oopt = build_pair(ctx, sizes, N, S, 30.0, precision, reparam, seed=3,
                                      operand_round=O.round_bf16 if bf else None)
    rng = np.random.RandomState(8)
    nvb = len(ref.vb_all)
    for it in range(2):
        Xn = rng.randn(N, sizes[0]); Tn = rng.randint(1, sizes[-1] + 1, N).astype(np.float64)
        step = ctx.get_step()
        noise = [[torch.from_numpy(cpu(net.model[k].draw_noise(step, s, rows=N)).astype(np.float64))
                  for k in range(nvb)] for s in range(S)]
        err, acc = net.train_step(torch.from_numpy(Xn).float().cuda(), torch.from_numpy(Tn).float().cuda())
        kw = dict(zeta=noise) if reparam == "local" else dict(eps=noise)
        rerr, racc, accs = oracle_step(ref, oopt, torch.from_numpy(Xn), torch.from_numpy(Tn), **kw)
        assert abs(err - rerr) < (5e-3 if bf else 1e-4) * abs(rerr), (it, err, rerr)
        for k, ol in enumerate(ref.vb_all):
            gl = net.model[k]
            assert rel(cpu(gl.gradWeight), accs[k][0].numpy()) < (1e-2 if bf else 1e-5), (it, k)
            assert rel(cpu(gl.gradSum), accs[k][1].numpy()) < (2e-2 if bf else 2e-5), (it, k)
            assert rel(cpu(gl.means), ol.means.numpy()) < (5e-3 if bf else 2e-4), (it, k)
            assert rel(cpu(gl.lvars), ol.lvars.numpy()) < (5e-3 if bf else 2e-4), (it, k)
        assert rel(cpu(net.model[-1].weight), ref.out.weight.numpy()) < (5e-3 if bf else 2e-4)


def test_batch_larger_than_max_batch_is_rejected(ctx):
    from vbnn_b200 import _lib as L
    net, _, _, _ = build_pair(ctx, [13, 7, 2], 4, 1, 30.0, "fp32", "weight")
    X = torch.zeros(5, 13).cuda(); T = torch.ones(5).cuda()
    with pytest.raises(L.VbnnError) as e:
        net.train_step(X, T)
    assert e.value.code == L.E_INVALID


def test_convnet_head_behind_the_same_abi(ctx):
    """BASELINE configs[3] (convnet.lua:14-33): the conv feature extractor is stock library code
    (torch / cuDNN here, cunn in the reference); its 2-layer VBLinear head -- VBLinear(256,100) + ReLU
    + VBLinear(100,C), both sampled (vb_indices = {8,10}, convnet.lua:9) -- runs through the layer-level
    C ABI on the extractor's output, and the gradient w.r.t. the features flows back into it."""
    import vbnn_b200
    from vbnn_b200 import _lib as L
    torch.manual_seed(0)
    N, C = 64, 10
    conv = torch.nn.Sequential(torch.nn.Conv2d(3, 16, 5), torch.nn.ReLU(), torch.nn.MaxPool2d(2),      # convnet.lua:14-20
                               torch.nn.Conv2d(16, 16, 5), torch.nn.ReLU(), torch.nn.MaxPool2d(2),
                               torch.nn.Flatten(), torch.nn.Linear(16 * 5 * 5, 256), torch.nn.ReLU()).cuda()
    opt = vbnn_b200.default_opt(S=1, B=50.0, batchSize=N, mu_init=1, var_init=0.01, strict_reference=False, log=False)
    h1 = vbnn_b200.VBLinear(256, 100, opt, ctx)
    h2 = vbnn_b200.VBLinear(100, C, opt, ctx)
    x = torch.randn(N, 3, 32, 32, device="cuda")
    t = torch.randint(0, C, (N,), device="cuda")
    feats = conv(x)
    f = feats.detach().contiguous()
    for h in (h1, h2):
        h.resetAcc(); h.sample(sample_idx=0)
    a1 = h1.updateOutput(f)
    r1 = torch.clamp(a1, min=0)
    logits = h2.updateOutput(r1).clone().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(logits, t)
    loss.backward()
    g2 = logits.grad.contiguous()
    gr1 = h2.updateGradInput(r1, g2); h2.accGradParameters(r1, g2)
    g1 = (gr1 * (a1 > 0)).contiguous()
    gf = h1.updateGradInput(f, g1); h1.accGradParameters(f, g1)
    feats.backward(gf)                                     # into the stock extractor
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in conv.parameters())
    # the head's data gradient equals autograd through the same sampled weights
    W1 = h1.get(L.BUF_WEIGHT).cuda(); W2 = h2.get(L.BUF_WEIGHT).cuda()
    f2 = f.clone().requires_grad_(True)
    l2 = torch.nn.functional.cross_entropy(torch.clamp(f2 @ W1.t() + h1.bias, min=0) @ W2.t() + h2.bias, t)
    l2.backward()
    assert rel(cpu(gf), cpu(f2.grad)) < 1e-4
    assert abs(l2.item() - loss.item()) < 1e-4 * abs(loss.item())
    before = cpu(h1.means).copy()
    h1.update(opt); h2.update(opt)                         # convnet.lua:105-117
    assert not np.array_equal(cpu(h1.means), before)


def test_c3_exact_size_properties(ctx):
    """BASELINE configs[2] at its exact size (4096-4096x4-1000, batch 8192, local reparameterisation): too big
    for the CPU oracle in seconds, so size-independent properties through the same C ABI --
    (1) the bf16 tcgen05 path against this library's fp32 CUDA-core path (itself oracle-checked at small
        sizes above) with identical Philox noise: loss <= 2e-2, gradients <= 0.12 / 0.15, means <= 2e-2;
    (2) counter-based noise makes the minibatch reproducible: re-running it from the same parameters and
        step counter gives bit-identical gradWeight / gradSum (no atomics on that path);
    (3) the loss of a freshly initialised net on random labels is close to log(1000) + the LRT noise term
        and finite everywhere."""
    import vbnn_b200
    sizes, N = [4096, 4096, 4096, 4096, 4096, 1000], 8192
    g = torch.Generator().manual_seed(3)
    X = torch.randn(N, sizes[0], generator=g).cuda()
    T = torch.randint(1, sizes[-1] + 1, (N,), generator=g).float().cuda()
    res = {}
    for precision in ("fp32", "bf16"):
        opt = vbnn_b200.default_opt(input_size=sizes[0], hidden=sizes[1:-1], classes=[str(i) for i in range(sizes[-1])],
                                    S=1, B=100.0, batchSize=N, testBatchSize=N, mu_init=1, var_init=0.001,
                                    reparam="local", precision=precision, strict_reference=False, log=False, seed=5)
        net = vbnn_b200.MLP(opt, ctx, max_batch=N)
        net.init_params(seed=4, he_means=True)
        ctx.set_step(21)
        err, acc = net.train_step(X, T)
        assert math.isfinite(err) and 0.0 <= acc <= 100.0
        gw = [m.gradWeight.clone() for m in net.model]
        gs = [m.gradSum.clone() for m in net.model[:-1]]
        mu = [m.means.clone() for m in net.model[:-1]]
        for t in gw + gs + mu:
            assert bool(torch.isfinite(t).all())
        res[precision] = (err, gw, gs, mu)
        if precision == "bf16":
            # (2) same parameters + same step counter -> same minibatch, bit for bit
            net2 = vbnn_b200.MLP(opt, ctx, max_batch=N)
            net2.init_params(seed=4, he_means=True)
            ctx.set_step(21)
            err2, _ = net2.train_step(X, T)
            assert abs(err2 - err) < 1e-5 * abs(err)       # the loss sum itself is an atomic accumulation
            for a, m in zip(gw, net2.model):
                assert torch.equal(a, m.gradWeight)
            for a, m in zip(gs, net2.model[:-1]):
                assert torch.equal(a, m.gradSum)
            del net2
        del net
        torch.cuda.empty_cache()
    e32, e16 = res["fp32"][0], res["bf16"][0]
    assert abs(e16 - e32) < 2e-2 * abs(e32), (e16, e32)
    assert e32 > 0.8 * math.log(1000.0)
    relt = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    for a, b in zip(res["bf16"][1], res["fp32"][1]):
        assert relt(a, b) < 0.12
    for a, b in zip(res["bf16"][2], res["fp32"][2]):
        assert relt(a, b) < 0.15
    for a, b in zip(res["bf16"][3], res["fp32"][3]):
        assert relt(a, b) < 2e-2


# ---------------------------------------------------------------------------------------------
# The CTA-pair (cta_group::2, 256 x 256) tiles are what bench.py times on C3; at oracle-sized problems the
# cost model picks 128 x 128 single-CTA tiles, so these tests FORCE the pair tiles (vbnn_debug_knob) at a
# size with >= 2 pair tiles per dimension and ragged edges in every dimension, and compare every GEMM form
# (dual / split forward, dual / split backward-data, dual / split dW, weight-mode fwd / dx / multi-sample dW)
# with the bf16-rounding oracle at the tight tolerance.
@pytest.mark.parametrize("reparam,S,lrt_split,dw_split", [("local", 1, 1, 0),    # bench default: split fwd, dual dx, dual dW
                                                          ("local", 1, 0, 1),    # dual fwd, split dW (the peer-mode form)
                                                          ("local", 2, 3, 0),    # split fwd + split dx, 2 samples
                                                          ("weight", 1, 1, 0),   # EPI_FWD / EPI_DX / EPI_DW pair tiles
                                                          ("weight", 3, 1, 0)])  # multi-sample dW, global accumulate
def test_pair_tile_epilogues_vs_oracle(ctx, reparam, S, lrt_split, dw_split):
    """Two checks per tensor: (1) against the bf16-rounding oracle at the tolerance of the other bf16 tests;
    (2) against THIS library's default tiling (128 x 128 single-CTA tiles at this size, oracle-checked by the tests
    above) on the same inputs and noise -- identical rounding points, so only fp32 summation order differs: <= 1e-3,
    whole tensor and per 64-row / 64-column band (a wrong 32-column chunk or a row offset in ONE CTA of the pair would
    be diluted in the Frobenius norm of the whole matrix, not in its band)."""
    import vbnn_b200
    sizes, N = [600, 552, 530, 10], 540
    names = ("tc_bn", "tc_cg", "lrt_split", "dw_split")
    lrt = reparam == "local"
    rng = np.random.RandomState(17)
    data = [(rng.randn(N, sizes[0]), rng.randint(1, sizes[-1] + 1, N).astype(np.float64)) for _ in range(2)]
    runs = {}
    for forced in (False, True):
        old = [vbnn_b200.knob(n) for n in names]
        if forced:
            for n, v in zip(names, (256, 2, lrt_split, dw_split)):
                vbnn_b200.knob(n, v)
        try:
            ctx.set_step(40)
            net, ref, gopt, oopt = build_pair(ctx, sizes, N, S, 30.0, "bf16", reparam, seed=7, strict=False,
                                              operand_round=O.round_bf16)
            nvb = len(ref.vb_all)
            trace = []
            for it, (Xn, Tn) in enumerate(data):      # second minibatch: graph replay + Adam state carried over
                step = ctx.get_step()
                noise = [[torch.from_numpy(cpu(net.model[k].draw_noise(step, s, rows=N)).astype(np.float64))
                          for k in range(nvb)] for s in range(S)]
                err, acc = net.train_step(torch.from_numpy(Xn).float().cuda(), torch.from_numpy(Tn).float().cuda())
                kw = dict(zeta=noise) if lrt else dict(eps=noise)
                rerr, racc, accs = oracle_step(ref, oopt, torch.from_numpy(Xn), torch.from_numpy(Tn), **kw)
                assert abs(err - rerr) < 5e-3 * abs(rerr), (forced, it, err, rerr)
                snap = dict(err=err)
                for k, ol in enumerate(ref.vb_all):
                    gl = net.model[k]
                    gw, gs, gb = accs[k]
                    snap[k] = dict(gW=cpu(gl.gradWeight).copy(), gS=cpu(gl.gradSum).copy(), gb=cpu(gl.gradBias).copy(),
                                   mu=cpu(gl.means).copy(), lv=cpu(gl.lvars).copy())
                    assert rel(snap[k]["gW"], gw.numpy()) < (5e-3 if lrt else 1e-2), (forced, it, k, "gW vs oracle")
                    assert rel(snap[k]["gS"], gs.numpy()) < (1e-2 if lrt else 2e-2), (forced, it, k, "gS vs oracle")
                    assert rel(snap[k]["gb"], gb.numpy()) < 1e-2, (forced, it, k, "gb vs oracle")
                    assert rel(snap[k]["mu"], ol.means.numpy()) < 5e-3, (forced, it, k, "means vs oracle")
                    assert rel(snap[k]["lv"], ol.lvars.numpy()) < 5e-3, (forced, it, k, "lvars vs oracle")
                snap["out_w"] = cpu(net.model[-1].weight).copy(); snap["out_gw"] = cpu(net.model[-1].gradWeight).copy()
                assert rel(snap["out_w"], ref.out.weight.numpy()) < 5e-3
                assert rel(snap["out_gw"], ref.out.gradWeight.numpy()) < 1e-2
                trace.append(snap)
            runs[forced] = trace
            del net
        finally:
            for n, v in zip(names, old):
                vbnn_b200.knob(n, v)
    for it in range(2):
        a, b = runs[True][it], runs[False][it]
        assert abs(a["err"] - b["err"]) < 1e-4 * abs(b["err"]), (it, a["err"], b["err"])
        for k in range(2):
            for name in ("gW", "gS", "gb", "mu", "lv"):
                A, B = a[k][name], b[k][name]
                assert rel(A, B) < 1e-3, (it, k, name, rel(A, B))
                if A.ndim == 2 and name in ("gW", "gS"):
                    for r0 in range(0, A.shape[0], 64):
                        assert rel(A[r0:r0 + 64], B[r0:r0 + 64]) < 3e-3, (it, k, name, "rows", r0)
                    for c0 in range(0, A.shape[1], 64):
                        assert rel(A[:, c0:c0 + 64], B[:, c0:c0 + 64]) < 3e-3, (it, k, name, "cols", c0)
        assert rel(a["out_w"], b["out_w"]) < 1e-4 and rel(a["out_gw"], b["out_gw"]) < 1e-3
    # (3) a race in the mbarrier / TMEM / cluster pipelines shows up as run-to-run differences (compute-sanitizer's
    # racecheck is closed on this pool): the same minibatch from the same state, three times, must be BIT-identical
    import vbnn_b200 as vb
    old = [vb.knob(n) for n in names]
    for n, v in zip(names, (256, 2, lrt_split, dw_split)):
        vb.knob(n, v)
    try:
        Xd, Td = torch.from_numpy(data[0][0]).float().cuda(), torch.from_numpy(data[0][1]).float().cuda()
        outs = []
        for rep in range(3):
            ctx.set_step(40)
            net, _, _, _ = build_pair(ctx, sizes, N, S, 30.0, "bf16", reparam, seed=7, strict=False)
            net.train_step(Xd, Td)
            # (gradBias is a column sum finished with fp32 atomics: order-dependent by design, not compared)
            outs.append([m.gradWeight.clone() for m in net.model] +
                        [m.gradSum.clone() for m in net.model[:-1]] + [m.means.clone() for m in net.model[:-1]])
            del net
        for o in outs[1:]:
            for a, b in zip(outs[0], o):
                assert torch.equal(a, b), "forced pair tiles are not run-to-run deterministic"
    finally:
        for n, v in zip(names, old):
            vb.knob(n, v)


@pytest.mark.parametrize("M,N,K", [(540, 600, 552), (530, 552, 540)])
def test_pair_tile_layer_outputs_vs_torch(ctx, M, N, K):
    """The same forced pair tiles on the raw GEMM (EPI_STORE -> the `aux` product of the split forms): every
    element against fp64 on bf16-rounded operands (fp32 accumulate: <= 1e-5)."""
    import vbnn_b200
    from vbnn_b200 import _lib as L
    import ctypes as C
    old = [vbnn_b200.knob("tc_bn", 256), vbnn_b200.knob("tc_cg", 2)]
    try:
        g = torch.Generator().manual_seed(M + N)
        r8 = lambda v: (v + 7) // 8 * 8
        for ak, bk in ((1, 1), (1, 0), (0, 0)):                      # fwd, dx, dw operand majors
            A = torch.randn(M, K, generator=g).bfloat16(); B = torch.randn(N, K, generator=g).bfloat16()
            ref = (A.double() @ B.double().t()).numpy()
            if ak:
                lda = r8(K); Ad = torch.zeros(M, lda, dtype=torch.bfloat16); Ad[:, :K] = A
            else:
                lda = r8(M); Ad = torch.zeros(K, lda, dtype=torch.bfloat16); Ad[:, :M] = A.t()
            if bk:
                ldb = r8(K); Bd = torch.zeros(N, ldb, dtype=torch.bfloat16); Bd[:, :K] = B
            else:
                ldb = r8(N); Bd = torch.zeros(K, ldb, dtype=torch.bfloat16); Bd[:, :N] = B.t()
            Ad, Bd = Ad.cuda(), Bd.cuda()
            ldd = r8(N)
            D = torch.full((M, ldd), float("nan"), device="cuda")
            L.check(L.lib().vbnn_gemm_bf16(ctx.handle, C.c_void_p(Ad.data_ptr()), lda, ak, C.c_void_p(Bd.data_ptr()), ldb, bk,
                                           C.c_void_p(D.data_ptr()), ldd, M, N, K, 1, 0, 0, 0))
            ctx.synchronize()
            got = cpu(D)[:, :N]
            assert np.isfinite(got).all()
            assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), (ak, bk)
    finally:
        vbnn_b200.knob("tc_bn", old[0]); vbnn_b200.knob("tc_cg", old[1])


def test_snr_mask_vs_oracle(ctx):
    """the signal-to-noise pruning mask of mainviz.lua:20-24 on random parameters whose
    |mu|/sigma straddles the threshold, element for element against the oracle restatement."""
    import vbnn_b200
    from vbnn_b200 import _lib as L
    opt = vbnn_b200.default_opt(S=1, B=50.0, mu_init=1, var_init=0.01, strict_reference=False, log=False)
    lyr = vbnn_b200.VBLinear(203, 77, opt, ctx)
    rng = np.random.RandomState(4)
    lv = rng.uniform(math.log(1e-4), math.log(1e-1), (77, 203)).astype(np.float32)
    snr = np.abs(rng.standard_cauchy((77, 203))) * 0.005              # about half below the 0.005 threshold
    sign = rng.choice([-1.0, 1.0], (77, 203))
    mu = (sign * snr * np.exp(0.5 * lv.astype(np.float64))).astype(np.float32)
    lyr.set(L.BUF_MEANS, mu); lyr.set(L.BUF_LVARS, lv)
    for thresh in (0.005, 0.001, 0.05):
        mask, count = lyr.snr_prune_mask(thresh)
        rmask, rcount = O.snr_prune_mask(torch.from_numpy(mu).double(), torch.from_numpy(lv).double(), thresh)
        got = mask.cpu().numpy().astype(bool)
        want = rmask.numpy()
        # fp32 exp / divide on the GPU vs fp64: only elements within 1e-5 (relative) of the threshold may differ
        ratio = np.abs(mu.astype(np.float64)) / np.exp(0.5 * lv.astype(np.float64))
        near = np.abs(ratio - thresh) < 1e-5 * thresh
        assert np.array_equal(got[~near], want[~near])
        assert abs(count - rcount) <= int(near.sum())
        assert 0.05 * mu.size < count < 0.95 * mu.size                  # the threshold really splits the set
        assert lyr.snr_prune_count(thresh) == count


def test_epoch_loop_vs_oracle(ctx):
    """main:train (main.lua:13-53) over a shuffled epoch of a device-resident dataset --
    the fused loop (vbnn_b200.train.train) against the ORACLE's train_epoch visiting the same shuffled order
    with the device-drawn epsilon injected."""
    import vbnn_b200
    from vbnn_b200 import train as tr
    from vbnn_b200 import _lib as L
    over = dict(hidden=[32, 24], S=2, B=8.0, batchSize=50, testBatchSize=100, trainSize=400, testSize=200, mu_init=1,
                var_init=0.01, log=False, testSamples=3, strict_reference=True, meanState=dict(learningRate=0.002))
    opt = vbnn_b200.default_opt(**over)
    ctx.set_step(0)
    net = vbnn_b200.MLP(opt, ctx, max_batch=100)
    oopt = O.default_opt(**over)
    ref = O.MLPOracle(oopt, torch.float64, seed=3)
    rng = np.random.RandomState(21)
    for gl, ol in zip(net.model, ref.vb + [ref.out]):
        if isinstance(ol, O.VBLinearOracle):
            mu = rng.randn(ol.O, ol.I) * math.sqrt(2.0 / ol.I)
            lv = rng.uniform(math.log(1e-4), math.log(1e-2), (ol.O, ol.I))
            ol.means.copy_(torch.from_numpy(mu)); ol.lvars.copy_(torch.from_numpy(lv)); ol.compute_prior()
            gl.set(L.BUF_MEANS, mu); gl.set(L.BUF_LVARS, lv); gl.compute_prior()
        else:
            w = rng.randn(*ol.weight.shape) * math.sqrt(2.0 / ol.weight.shape[1])
            ol.weight.copy_(torch.from_numpy(w)); ol.bias.zero_()
            gl.set(L.BUF_WEIGHT, w)
    ds = tr.synthetic_dataset(400, 784, 10, seed=3, geometry=(28, 28))
    ods = dict(inputs=ds["inputs"].double().cpu(), targets=ds["targets"].double().cpu())
    # the same shuffled order main.lua:18 / utils.shuffle would visit
    starts = torch.arange(0, 400, 50)
    order = starts[torch.randperm(len(starts), generator=torch.Generator().manual_seed(1))].tolist()
    eps_fn = lambda i: [[torch.from_numpy(cpu(net.model[k].draw_noise(i, s)).astype(np.float64)) for k in range(2)]
                        for s in range(2)]                                # Philox step i = i-th visited minibatch
    noise = [eps_fn(i) for i in range(len(order))]
    acc, err = tr.train(net, ds, opt, fused=True, shuffle_seed=1)
    racc, rerr = O.train_epoch(ref, ods, oopt, order, eps_fn=lambda i: noise[i])
    assert abs(err - rerr) < 1e-4 * abs(rerr), (err, rerr)
    assert abs(acc - racc) < 1e-2, (acc, racc)
    for k in range(2):
        assert rel(cpu(net.model[k].means), ref.vb[k].means.numpy()) < 2e-4
        assert rel(cpu(net.model[k].lvars), ref.vb[k].lvars.numpy()) < 2e-4
    assert rel(cpu(net.model[2].weight), ref.out.weight.numpy()) < 2e-4
    # main:test (main.lua:55-74), MAP evaluation, against the oracle on the same rows
    net.opt["quicktest"] = True; oopt["quicktest"] = True
    tacc, terr = tr.test(net, dict(inputs=ds["inputs"][:200], targets=ds["targets"][:200]), opt)
    e = a = 0.0
    for t in range(0, 200, 100):
        e_, a_ = ref.test(ods["inputs"][t:t + 100], ods["targets"][t:t + 100])
        e += e_; a += a_
    assert abs(terr - e / 2) < 1e-4 * abs(e / 2) and abs(tacc - a / 2) < 1e-2


def test_checkpoint_files_and_metric_files(ctx, tmp_path):
    """the files the reference's scripts read -- `parameters/means`, `parameters/vars`, `opt`
    (mainviz.lua:12-15), `model` with `.old` rotation (utils.lua:73-80, main.lua:181) and the logger's per-id
    metric files (logger.lua:18-26) carrying the 14 diagnostics of VBLinear.lua:150-163 -- written from device
    state, read back, and resumed from (main.lua:146-148)."""
    import vbnn_b200
    from vbnn_b200 import checkpoint, logger, t7, train as tr
    from vbnn_b200 import _lib as L
    d = str(tmp_path / "exp")
    over = dict(hidden=[32], S=2, B=8.0, batchSize=50, testBatchSize=100, trainSize=200, testSize=100, mu_init=1,
                var_init=0.01, log=True, testSamples=2, network_name=d, meanState=dict(learningRate=0.002))
    opt = vbnn_b200.default_opt(**over)
    ctx.set_step(0)
    net = vbnn_b200.MLP(opt, ctx, max_batch=100)
    net.init_params(seed=4, he_means=True)
    oopt = O.default_opt(**over)
    ds = tr.synthetic_dataset(200, 784, 10, seed=3)
    logger.init(d)
    try:
        r = tr.epoch(net, ds, dict(inputs=ds["inputs"][:100], targets=ds["targets"][:100]), opt, shuffle_seed=2)
        # ---- metric files: one value per line; 14 ids x (4 minibatches x 1 VB layer) + 5 epoch metrics
        for name in L.STAT_NAMES:
            vals = logger.read_data(os.path.join(d, name))
            assert len(vals) == 4 and all(math.isfinite(v) for v in vals), name
        assert logger.read_data(os.path.join(d, "trainerr")) == [pytest.approx(r[1], rel=1e-12)]
        assert logger.read_data(os.path.join(d, "devacc")) == [pytest.approx(r[2], rel=1e-12)]
        assert logger.read_data(os.path.join(d, "lc")) == [pytest.approx(r[4], rel=1e-12)]
        # 'var hat' of the last minibatch is compute_prior() of the parameters BEFORE its Adam step (VBLinear.lua:130,157)
        assert logger.read_data(os.path.join(d, "min variance"))[-1] > 0
    finally:
        logger.Log.close(); logger.Log = None
    # ---- parameters/means, parameters/vars, opt: what mainviz.lua:12-24 computes from them
    checkpoint.save_parameters(net, d)
    means = t7.load(os.path.join(d, "parameters", "means"))
    vars_ = t7.load(os.path.join(d, "parameters", "vars"))
    lopt = t7.load(os.path.join(d, "opt"))
    assert means.dtype == np.float32 and means.shape == (32 * 784,) and vars_.shape == means.shape
    assert np.array_equal(means, cpu(net.model[0].means).ravel())
    assert rel(vars_, np.exp(cpu(net.model[0].lvars).astype(np.float64)).ravel()) < 1e-6
    assert lopt["hidden"] == [32] and lopt["S"] == 2 and lopt["varState"]["learningRate"] == 0.05
    _, count, _, _ = checkpoint.snr_pruned(means, vars_, 0.005)
    assert abs(count - net.model[0].snr_prune_count(0.005)) <= 2
    # ---- model (+ .old): resume is exact
    assert os.path.isfile(os.path.join(d, "model"))
    checkpoint.save_net(net, d)
    assert os.path.isfile(os.path.join(d, "model.old"))
    X, T = ds["inputs"][:50], ds["targets"][:50]
    net.opt["log"] = False
    ra = [net.train_step(X, T) for _ in range(2)]
    net2 = vbnn_b200.MLP(vbnn_b200.default_opt(**dict(over, log=False)), ctx, max_batch=100)
    checkpoint.load_net(net2, d)
    rb = [net2.train_step(X, T) for _ in range(2)]
    assert np.allclose(np.array(ra), np.array(rb), rtol=1e-6, atol=1e-7)
    assert rel(cpu(net2.model[0].means), cpu(net.model[0].means)) < 1e-7


@pytest.mark.parametrize("precision,reparam", [("fp32", "weight"), ("bf16", "local")])
def test_submit_host_u8_equals_fp32_host_path(ctx, precision, reparam):
    """The uint8 host format (MNIST bytes before data.lua:25 u.normalize; normalisation fused into the operand
    staging on the device) gives the same minibatches as the fp32 host format fed (x - mean) * (1 / std)."""
    sizes, N = [52, 48, 36, 6], 40                                       # 52 % 8 != 0: the ragged byte-row path too
    rng = np.random.RandomState(31)
    px = torch.from_numpy(rng.randint(0, 256, (N, sizes[0])).astype(np.uint8))
    Th = torch.from_numpy(rng.randint(1, 7, N).astype(np.float32))
    mean, std = 33.3, 78.6
    inv = np.float32(1.0 / std)
    Xf = ((px.float() - np.float32(mean)) * inv).contiguous().pin_memory()
    res = []
    for fmt in ("f32", "u8"):
        ctx.set_step(3)
        net, _, _, _ = build_pair(ctx, sizes, N, 2, 30.0, precision, reparam, seed=4, strict=False)
        out = []
        for _ in range(3):
            if fmt == "f32":
                net.submit_host(Xf, Th)
            else:
                net.submit_host_u8(px.pin_memory(), Th, mean, std)
            out.append(net.collect())
        res.append((out, cpu(net.model[0].means).copy(), cpu(net.model[1].lvars).copy()))
    assert np.allclose(np.array(res[0][0]), np.array(res[1][0]), rtol=1e-6, atol=1e-7)
    assert rel(res[1][1], res[0][1]) < 1e-7 and rel(res[1][2], res[0][2]) < 1e-7
