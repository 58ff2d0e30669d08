"""CPU tests of the oracle itself (-m "not gpu").

The reference ships no tests or golden vectors, so the oracle is pinned by
the six analytic invariants derived from the reference's formulas, by Random123's published
known-answer vectors for Philox4x32-10, and by torch fp64 autograd for the local-reparameterisation
formulas (which have no reference code at all)."""
import math

import numpy as np
import pytest
import torch

from oracle import vbnn_oracle as O

torch.manual_seed(0)


def make_layer(I=7, O_=5, **kw):
    opt = O.default_opt(B=50.0, S=3, mu_init=1, var_init=0.01, **kw)
    rng = np.random.RandomState(1)
    lyr = O.VBLinearOracle(I, O_, opt, torch.float64, rng)
    lyr.lvars.copy_(torch.from_numpy(rng.uniform(math.log(1e-4), math.log(1e-2), (O_, I))))
    lyr.compute_prior()
    return lyr, opt


def lc_sum(lyr, opt):
    lyr.compute_prior()
    return float(lyr.calc_lc(opt).sum())


def fd(lyr, opt, tensor, h=1e-6):
    g = torch.zeros_like(tensor)
    flat, gf = tensor.view(-1), g.view(-1)
    for k in range(flat.numel()):
        old = float(flat[k])
        flat[k] = old + h
        f1 = lc_sum(lyr, opt)
        flat[k] = old - h
        f2 = lc_sum(lyr, opt)
        flat[k] = old
        gf[k] = (f1 - f2) / (2 * h)
    lyr.compute_prior()
    return g


def test_invariant1_fd_lc_wrt_lvars_is_vargrads_lcg():
    lyr, opt = make_layer()
    g = fd(lyr, opt, lyr.lvars)
    _, lcg = lyr.compute_vargrads(opt)
    assert torch.allclose(g, lcg, rtol=1e-5, atol=1e-10)


def test_invariant2_fd_lc_wrt_means_is_mugrads_lcg():
    lyr, opt = make_layer()
    g = fd(lyr, opt, lyr.means)
    _, lcg = lyr.compute_mugrads(opt)
    assert torch.allclose(g, lcg, rtol=1e-5, atol=1e-10)


def test_invariant3_lc_identity():
    lyr, opt = make_layer()
    total = lc_sum(lyr, opt)
    expect = float((0.5 * torch.log(lyr.var_hat / lyr.vars)).sum() / opt["B"])
    assert abs(total - expect) < 1e-12 * max(1.0, abs(expect))


def test_invariant4_likelihood_term_of_lvar_gradient():
    # L = <G, mu + sigma*eps>  =>  dL/dlvar = G*eps*sigma/2 == gradSum/(2S)*stdv with S samples
    lyr, opt = make_layer()
    rng = np.random.RandomState(7)
    N, S = 6, opt["S"]
    X = torch.from_numpy(rng.randn(N, lyr.I))
    lyr.resetAcc()
    lyr.gradWeight.zero_()
    total = torch.zeros_like(lyr.lvars)
    for s in range(S):
        eps = torch.from_numpy(rng.randn(lyr.O, lyr.I))
        lyr.sample(eps)
        G = torch.from_numpy(rng.randn(N, lyr.O))
        lyr.accGradParameters(X, G)
        lv = lyr.lvars.clone().requires_grad_(True)
        w = lyr.means + torch.exp(0.5 * lv) * eps
        ((X @ w.t()) * G).sum().backward()
        total += lv.grad
    leg, _ = lyr.compute_vargrads(opt)
    assert torch.allclose(leg, total / S, rtol=1e-10, atol=1e-12)


def test_invariant5_clamp_to_map_is_plain_linear():
    lyr, _ = make_layer()
    X = torch.randn(4, lyr.I, dtype=torch.float64)
    lyr.clamp_to_map()
    assert torch.allclose(lyr.updateOutput(X), X @ lyr.means.t() + lyr.bias)


def test_invariant6_zero_eps_gives_plain_backprop():
    lyr, _ = make_layer()
    X = torch.randn(4, lyr.I, dtype=torch.float64)
    G = torch.randn(4, lyr.O, dtype=torch.float64)
    lyr.sample(torch.zeros(lyr.O, lyr.I))
    lyr.resetAcc()
    lyr.gradWeight.zero_()
    lyr.accGradParameters(X, G)
    assert float(lyr.gradSum.abs().max()) == 0.0
    assert torch.allclose(lyr.gradWeight, G.t() @ X)
    assert torch.allclose(lyr.updateGradInput(X, G), G @ lyr.means)


def test_q1_sample_uses_stale_sigma():
    lyr, opt = make_layer()
    stale = lyr.stdv.clone()
    lyr.lvars.add_(1.0)                      # as an Adam step would, without compute_prior
    eps = torch.ones(lyr.O, lyr.I, dtype=torch.float64)
    lyr.sample(eps)
    assert torch.allclose(lyr.weight, lyr.means + stale)


def test_adam_matches_torch_optim():
    x = torch.randn(5, 3, dtype=torch.float64)
    ref = x.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.05, betas=(0.9, 0.999), eps=1e-8)
    st = dict(learningRate=0.05)
    for k in range(5):
        g = torch.randn(5, 3, dtype=torch.float64)
        ref.grad = g.clone()
        opt.step()
        O.optim_adam(x, g, st)
        # torch's Adam puts eps outside sqrt(v)/sqrt(bc2); the 2015 optim.adam puts it on sqrt(v):
        # identical up to O(eps)
        assert torch.allclose(x, ref.detach(), rtol=1e-6, atol=1e-7)


def test_lrt_formulas_match_autograd():
    opt = O.default_opt(B=50.0, S=1, mu_init=1, var_init=0.01, reparam="local")
    rng = np.random.RandomState(2)
    lyr = O.VBLinearOracle(6, 4, opt, torch.float64, rng)
    lyr.bias.copy_(torch.from_numpy(rng.randn(4)))
    N = 5
    X = torch.from_numpy(rng.randn(N, 6)).requires_grad_(True)
    zeta = torch.from_numpy(rng.randn(N, 4))
    G = torch.from_numpy(rng.randn(N, 4))
    mu = lyr.means.clone().requires_grad_(True)
    lv = lyr.lvars.clone().requires_grad_(True)
    Y = X @ mu.t() + lyr.bias + torch.sqrt((X * X) @ torch.exp(lv).t()) * zeta
    (Y * G).sum().backward()
    Xd = X.detach()
    out = lyr.updateOutput(Xd, zeta)
    assert torch.allclose(out, Y.detach())
    lyr.resetAcc(); lyr.gradWeight.zero_(); lyr.gradBias.zero_()
    dX = lyr.updateGradInput(Xd, G)
    lyr.accGradParameters(Xd, G)
    assert torch.allclose(dX, X.grad, rtol=1e-10, atol=1e-12)
    assert torch.allclose(lyr.gradWeight, mu.grad, rtol=1e-10, atol=1e-12)
    lyr.compute_prior()
    leg, _ = lyr.compute_vargrads(opt)
    assert torch.allclose(leg, lv.grad, rtol=1e-10, atol=1e-12)


def test_lrt_moments_match_weight_sampling():
    # one row: Y under weight sampling has mean X mu^T and variance X^2 s2^T -- what LRT draws
    opt = O.default_opt(mu_init=1, var_init=0.05, strict_reference=False)
    rng = np.random.RandomState(3)
    lyr = O.VBLinearOracle(8, 3, opt, torch.float64, rng)
    x = torch.from_numpy(rng.randn(1, 8))
    ys = []
    for _ in range(4000):
        lyr.sample()
        ys.append(lyr.updateOutput(x).clone())
    ys = torch.cat(ys)
    mean = x @ lyr.means.t()
    var = (x * x) @ torch.exp(lyr.lvars).t()
    assert torch.allclose(ys.mean(0, keepdim=True), mean, atol=4 * float(var.max().sqrt()) / math.sqrt(4000))
    assert torch.allclose(ys.var(0, keepdim=True), var, rtol=0.15)


def test_philox_known_answer_vectors():
    # Random123 kat_vectors, philox4x32 10 rounds
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, out in kats:
        got = O.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == out


def test_philox_normal_moments():
    z = O.philox_normal_matrix(seed=5, step=1, stream=7, sample=0, rows=400, cols=1001)
    assert z.shape == (400, 1001)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1.0) < 0.01
    kurt = ((z - z.mean()) ** 4).mean() / z.var() ** 2
    assert abs(kurt - 3.0) < 0.05
    a, b = z[:, :-1].ravel(), z[:, 1:].ravel()
    assert abs(np.corrcoef(a, b)[0, 1]) < 0.01                      # lag-1
    z2 = O.philox_normal_matrix(seed=5, step=1, stream=7, sample=1, rows=400, cols=1001)
    assert abs(np.corrcoef(z.ravel(), z2.ravel())[0, 1]) < 0.01     # across samples
    # row-offset invariance: a shard starting at row 100 sees the same numbers (data parallel)
    zs = O.philox_normal_matrix(seed=5, step=1, stream=7, sample=0, rows=50, cols=1001, row0=100)
    assert np.array_equal(zs, z[100:150])


def test_train_minibatch_runs_and_learns():
    opt = O.default_opt(hidden=[16], S=2, B=10.0, input_size=12, classes=list("abc"), mu_init=1,
                        var_init=0.01, batchSize=32)
    opt["meanState"] = dict(learningRate=0.01)
    net = O.MLPOracle(opt, torch.float64, seed=3)
    rng = np.random.RandomState(0)
    X = torch.from_numpy(rng.randn(32, 12))
    W = rng.randn(12, 3)
    T = torch.from_numpy((X.numpy() @ W).argmax(1) + 1.0)
    first = last = None
    for it in range(60):
        err, acc = O.train_minibatch(net, X, T, opt)
        first = first if first is not None else err
        last = err
    assert last < first


def test_inherited_torch7_pieces_match_torch_autograd():
    """The un-vendored Torch7 modules the path inherits (nn.Linear, nn.LogSoftMax, nn.ClassNLLCriterion with
    sizeAverage, nn.ReLU) are restated by hand in the oracle; pin those restatements to torch's own
    implementations and autograd (fp64)."""
    import torch.nn.functional as F
    rng = np.random.RandomState(2)
    N, I, C = 9, 7, 5
    lin = O.LinearOracle(I, C, torch.float64, rng=rng)
    x = torch.from_numpy(rng.randn(N, I))
    t1 = torch.from_numpy(rng.randint(1, C + 1, N).astype(np.float64))          # 1-based targets (data.lua:16)
    # forward
    y = lin.updateOutput(x)
    assert torch.allclose(y, F.linear(x, lin.weight, lin.bias), atol=1e-14)
    logp = O.log_softmax(y)
    assert torch.allclose(logp, F.log_softmax(y, dim=1), atol=1e-14)
    loss = O.class_nll_forward(logp, t1)
    assert abs(loss - float(F.nll_loss(logp, t1.long() - 1))) < 1e-14
    assert abs(O.get_accuracy(logp, t1) - 100.0 * float((logp.argmax(1) + 1 == t1.long()).double().mean())) < 1e-12
    # backward through criterion, LogSoftMax and Linear against autograd
    xa = x.clone().requires_grad_(True)
    wa = lin.weight.clone().requires_grad_(True)
    ba = lin.bias.clone().requires_grad_(True)
    F.nll_loss(F.log_softmax(F.linear(xa, wa, ba), dim=1), t1.long() - 1).backward()
    g = O.log_softmax_backward(logp, O.class_nll_backward(logp, t1))
    gx = lin.updateGradInput(x, g)
    lin.accGradParameters(x, g, 1.0)
    assert torch.allclose(gx, xa.grad, atol=1e-14)
    assert torch.allclose(lin.gradWeight, wa.grad, atol=1e-14)
    assert torch.allclose(lin.gradBias, ba.grad, atol=1e-14)
    # accGradParameters accumulates and honours `scale` (nn.Linear semantics)
    lin.accGradParameters(x, g, 0.5)
    assert torch.allclose(lin.gradWeight, 1.5 * wa.grad, atol=1e-14)


def test_sgd_matches_torch_optim():
    x = torch.arange(6, dtype=torch.float64).reshape(2, 3).clone()
    g = torch.ones_like(x) * 0.25
    p = x.clone().requires_grad_(True)
    opt = torch.optim.SGD([p], lr=1e-3)
    state = {"learningRate": 1e-3}
    for _ in range(3):
        p.grad = g.clone()
        opt.step()
        O.optim_sgd(x, g, state)
    assert torch.allclose(x, p.detach(), atol=1e-15)


def test_train_epoch_is_the_shuffled_sequence_of_minibatches():
    """main:train (main.lua:13-53): the oracle's epoch loop visits the given start indices in order, assembles rows
    [t, t + batchSize) (data.lua:9-20) and returns (accuracy / B, error / B) with B = trainSize / batchSize."""
    import copy
    opt = O.default_opt(input_size=12, hidden=[9], classes=list("abc"), S=2, B=6.0, batchSize=5, trainSize=20, mu_init=1,
                        var_init=0.01)
    net = O.MLPOracle(opt, torch.float64, seed=3)
    ref = copy.deepcopy(net)
    rng = np.random.RandomState(4)
    ds = dict(inputs=torch.from_numpy(rng.randn(20, 12)), targets=torch.from_numpy(rng.randint(1, 4, 20).astype(np.float64)))
    order = [10, 0, 15, 5]
    eps = [[[torch.from_numpy(rng.randn(9, 12))] for _ in range(2)] for _ in order]
    acc, err = O.train_epoch(net, ds, opt, order, eps_fn=lambda i: eps[i])
    a = e = 0.0
    for i, t in enumerate(order):
        e_, a_ = O.train_minibatch(ref, ds["inputs"][t:t + 5], ds["targets"][t:t + 5], opt, eps=eps[i])
        a += a_; e += e_
    assert abs(acc - a / 4) < 1e-12 and abs(err - e / 4) < 1e-12
    assert torch.equal(net.vb[0].means, ref.vb[0].means) and torch.equal(net.out.weight, ref.out.weight)


def test_snr_prune_mask_restates_mainviz():
    """mainviz.lua:20-22: pruned = lt(abs(means / sqrt(vars)), 0.005) with vars = exp(lvars)."""
    means = torch.tensor([[0.0, 1e-4, 0.5], [-1e-5, 4.9e-4, -5.1e-4]], dtype=torch.float64)
    lvars = torch.log(torch.tensor([[1e-2, 1e-2, 1e-2], [1e-6, 1e-2, 1e-2]], dtype=torch.float64))
    mask, count = O.snr_prune_mask(means, lvars, 0.005)
    assert mask.tolist() == [[True, True, False], [False, True, False]] and count == 3
