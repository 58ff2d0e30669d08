"""GPU parity tests of the kernels behind the layer-level C ABI (pytest -m gpu).

Every comparison is CUDA path (through libvbnn.so) vs the CPU oracle on identical inputs with
identical noise injected.  Tolerances: fp32 mode -- relative Frobenius error <= 1e-5 (the
north star asks <= 1e-3 "with fp32 accumulate"); bf16-operand mode -- <= 2e-2 against the fp64
fixture on outputs that are sums of bf16 products (single layer), stated per test; the tight
bf16 check against the oracle's bf16-operand restatement is in test_gpu_mlp.py."""
import ctypes as C
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


def test_library_loaded_is_in_tree():
    from vbnn_b200 import _lib
    assert os.path.dirname(_lib.LIB_PATH).endswith("vbnn_b200")
    assert _lib.lib().vbnn_abi_version() == 2


def test_philox_matches_cpu_restatement(ctx):
    from vbnn_b200 import _lib as L
    for rows, cols, row0 in [(33, 1001, 0), (7, 12, 100), (128, 784, 0)]:
        out = torch.empty(rows, cols, device="cuda")
        L.check(L.lib().vbnn_philox_normal(ctx.handle, 5, 3, (1 << 16) | 2, 1, rows, cols, row0,
                                           C.c_void_p(out.data_ptr())))
        ref = O.philox_normal_matrix(5, 3, (1 << 16) | 2, 1, rows, cols, row0)
        # identical Philox bits; Box-Muller uses fast intrinsics on the GPU
        assert np.abs(cpu(out) - ref).max() < 2e-4
    z = torch.empty(2000, 1000, device="cuda")
    L.check(L.lib().vbnn_philox_normal(ctx.handle, 9, 0, 0, 0, 2000, 1000, 0, C.c_void_p(z.data_ptr())))
    assert abs(float(z.mean())) < 3e-3 and abs(float(z.var()) - 1) < 5e-3
    kurt = float(((z - z.mean()) ** 4).mean() / z.var() ** 2)
    assert abs(kurt - 3) < 0.03


@pytest.mark.parametrize("ak,bk", [(1, 1), (1, 0), (0, 1), (0, 0)])
@pytest.mark.parametrize("M,N,K,batch", [(128, 128, 64, 1), (100, 10, 784, 1), (256, 384, 200, 1),
                                         (1024, 1200, 784, 2), (300, 520, 1000, 3)])
def test_tcgen05_gemm_all_layouts(ctx, ak, bk, M, N, K, batch):
    """Raw bf16 GEMM in every operand-major combination, ragged sizes (TMA zero fill), batched."""
    from vbnn_b200 import _lib as L
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + ak * 2 + bk)
    r8 = lambda v: (v + 7) // 8 * 8
    A = torch.randn(batch, M, K, generator=g)
    B = torch.randn(batch, N, K, generator=g)
    Ab, Bb = A.bfloat16(), B.bfloat16()
    ref = torch.matmul(Ab.float().double(), Bb.float().double().transpose(1, 2)).numpy()
    if ak:
        lda = r8(K); Ad = torch.zeros(batch, M, lda, dtype=torch.bfloat16); Ad[:, :, :K] = Ab; sA = M * lda
    else:
        lda = r8(M); Ad = torch.zeros(batch, K, lda, dtype=torch.bfloat16); Ad[:, :, :M] = Ab.transpose(1, 2); sA = K * lda
    if bk:
        ldb = r8(K); Bd = torch.zeros(batch, N, ldb, dtype=torch.bfloat16); Bd[:, :, :K] = Bb; sB = N * ldb
    else:
        ldb = r8(N); Bd = torch.zeros(batch, K, ldb, dtype=torch.bfloat16); Bd[:, :, :N] = Bb.transpose(1, 2); sB = K * ldb
    Ad, Bd = Ad.cuda(), Bd.cuda()
    D = torch.full((batch, M, N), float("nan"), device="cuda")
    L.check(L.lib().vbnn_gemm_bf16(ctx.handle, C.c_void_p(Ad.data_ptr()), lda, ak, C.c_void_p(Bd.data_ptr()), ldb, bk,
                                   C.c_void_p(D.data_ptr()), N, M, N, K, batch, sA, sB, M * N))
    ctx.synchronize()
    assert rel(cpu(D), ref) < 1e-5


def _make_layer(g, ctx, precision, reparam):
    import vbnn_b200
    opt = vbnn_b200.default_opt(B=float(g["B"]), S=int(g["S"]), mu_init=1, var_init=0.01,
                                reparam=reparam, precision=precision, log=True)
    lyr = vbnn_b200.VBLinear(int(g["I"]), int(g["O"]), opt, ctx)
    from vbnn_b200 import _lib as L
    lyr.set(L.BUF_MEANS, g["means0"]); lyr.set(L.BUF_LVARS, g["lvars0"]); lyr.set(L.BUF_BIAS, g["bias0"])
    return lyr, opt


@pytest.mark.parametrize("precision,tol_out,tol_acc", [("fp32", 1e-5, 1e-5), ("bf16", 1e-2, 2e-2)])
@pytest.mark.parametrize("reparam", ["weight", "local"])
def test_layer_api_against_golden(ctx, precision, tol_out, tol_acc, reparam):
    """VBLinear: sample -> updateOutput -> updateGradInput -> accGradParameters x S -> update,
    with the oracle's epsilon/zeta injected, against the committed fp64 fixture."""
    g = np.load(os.path.join(GOLD, f"vblinear_{reparam}.npz"))
    lyr, opt = _make_layer(g, ctx, precision, reparam)
    mu_hat, var_hat = lyr.compute_prior()
    assert mu_hat == 0.0 and abs(var_hat - float(g["var_hat0"])) < 1e-5 * float(g["var_hat0"])
    lyr.resetAcc()
    X = torch.from_numpy(g["X"]).float().cuda()
    for s in range(int(g["S"])):
        G = torch.from_numpy(g[f"G{s}"]).float().cuda()
        noise = torch.from_numpy(g[f"noise{s}"]).float().cuda()
        if reparam == "local":
            lyr.sample(sample_idx=s)
            Y = lyr.updateOutput(X, zeta=noise)
        else:
            lyr.sample(eps=noise, sample_idx=s)
            Y = lyr.updateOutput(X)
        dX = lyr.updateGradInput(X, G)
        lyr.accGradParameters(X, G, 1.0)
        assert rel(cpu(Y), g[f"Y{s}"]) < tol_out, ("Y", s)
        assert rel(cpu(dX), g[f"dX{s}"]) < tol_out * 2, ("dX", s)
    assert rel(cpu(lyr.gradWeight), g["gradWeight"]) < tol_acc
    assert rel(cpu(lyr.gradSum), g["gradSum"]) < tol_acc * 2
    assert rel(cpu(lyr.gradBias), g["gradBias"]) < 1e-5
    stats = lyr.update(opt)
    tol_p = 1e-5 if precision == "fp32" else 5e-3
    assert rel(cpu(lyr.means), g["means1"]) < tol_p
    assert rel(cpu(lyr.lvars), g["lvars1"]) < tol_p
    assert rel(cpu(lyr.bias), g["bias1"]) < 1e-5
    if precision == "fp32":
        for name, val in zip(g["stat_names"], g["stat_values"]):
            assert abs(stats[str(name)] - val) <= 2e-4 * abs(val) + 1e-9, (name, stats[str(name)], val)
        assert abs(lyr.calc_lc_sum() - float(g["lc_sum_cached"])) < 1e-4 * abs(float(g["lc_sum_cached"]))   # Q6
        assert rel(cpu(lyr.calc_lc(opt)).sum(), float(g["lc_sum_cached"])) < 1e-4


def test_q1_stale_sigma_and_second_adam_step(ctx):
    """Second minibatch: sample() must use sigma cached BEFORE the first Adam step (quirk Q1),
    and Adam's bias correction must advance to t=2."""
    g = np.load(os.path.join(GOLD, "vblinear_weight.npz"))
    lyr, opt = _make_layer(g, ctx, "fp32", "weight")
    lyr.compute_prior()
    X = torch.from_numpy(g["X"]).float().cuda()
    lyr.resetAcc()
    for s in range(int(g["S"])):
        lyr.sample(eps=torch.from_numpy(g[f"noise{s}"]).float().cuda(), sample_idx=s)
        lyr.updateOutput(X)
        lyr.accGradParameters(X, torch.from_numpy(g[f"G{s}"]).float().cuda(), 1.0)
    lyr.update(dict(opt, log=False))
    assert lyr.t == 1
    lyr.resetAcc()
    lyr.sample(eps=torch.from_numpy(g["noise0"]).float().cuda(), sample_idx=0)
    assert rel(cpu(lyr.weight), g["W_step2"]) < 1e-5
    lyr.updateOutput(X)
    lyr.accGradParameters(X, torch.from_numpy(g["G0"]).float().cuda(), 1.0)
    lyr.update(dict(opt, log=False))
    assert lyr.t == 2
    assert rel(cpu(lyr.means), g["means2"]) < 1e-5
    assert rel(cpu(lyr.lvars), g["lvars2"]) < 1e-5


def test_compute_grads_api(ctx):
    g = np.load(os.path.join(GOLD, "vblinear_weight.npz"))
    lyr, opt = _make_layer(g, ctx, "fp32", "weight")
    o = O.default_opt(B=float(g["B"]), S=int(g["S"]), mu_init=1, var_init=0.01)
    ref = O.VBLinearOracle(int(g["I"]), int(g["O"]), o, torch.float64)
    ref.means.copy_(torch.from_numpy(g["means0"])); ref.lvars.copy_(torch.from_numpy(g["lvars0"]))
    ref.gradWeight.copy_(torch.from_numpy(g["gradWeight"])); ref.gradSum.copy_(torch.from_numpy(g["gradSum"]))
    ref.compute_prior()
    from vbnn_b200 import _lib as L
    lyr.set(L.BUF_GRAD_WEIGHT, g["gradWeight"]); lyr.set(L.BUF_GRAD_SUM, g["gradSum"])
    lyr.compute_prior()
    mleg, mlcg = lyr.compute_mugrads(opt)
    vleg, vlcg = lyr.compute_vargrads(opt)
    rm, rc = ref.compute_mugrads(o)
    rv, rvc = ref.compute_vargrads(o)
    for a, b in [(mleg, rm), (mlcg, rc), (vleg, rv), (vlcg, rvc)]:
        assert rel(cpu(a), b.numpy()) < 1e-5
    assert rel(cpu(lyr.gradWeight), rm.numpy()) < 1e-6       # divided in place like VBLinear.lua:92


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_properties_c2_layer(ctx, precision):
    """BASELINE C2 layer (1200 x 1200, batch 1024): size-independent properties --
    clamp_to_map == plain linear (invariant 5); eps == 0 => gradSum == 0 and gradWeight == G^T X
    (invariant 6); accGradParameters is linear in gradOutput; LC identity (invariant 3)."""
    import vbnn_b200
    from vbnn_b200 import _lib as L
    I = O_ = 1200
    N = 1024
    opt = vbnn_b200.default_opt(B=58.6, S=1, mu_init=1, msr_init=True, precision=precision, log=False,
                                strict_reference=False)
    lyr = vbnn_b200.VBLinear(I, O_, opt, ctx)
    gen = torch.Generator(device="cpu").manual_seed(1)
    X = torch.randn(N, I, generator=gen).cuda()
    G1 = (torch.randn(N, O_, generator=gen) / N).cuda()
    G2 = (torch.randn(N, O_, generator=gen) / N).cuda()
    tol = 1e-5 if precision == "fp32" else 1e-2
    lyr.clamp_to_map()
    Y = lyr.updateOutput(X)
    ref = X.double() @ lyr.means.double().t() + lyr.bias.double()
    assert rel(cpu(Y), ref.cpu().numpy()) < tol
    # invariant 6
    lyr.resetAcc()
    lyr.sample(eps=torch.zeros(O_, I, device="cuda"), sample_idx=0)
    lyr.updateOutput(X)
    lyr.accGradParameters(X, G1, 1.0)
    assert float(lyr.gradSum.abs().max()) == 0.0
    gw1 = lyr.gradWeight.clone()
    assert rel(cpu(gw1), (G1.double().t() @ X.double()).cpu().numpy()) < tol
    # linearity: acc(G1) + acc(G2) == acc(G1 + G2)
    lyr.accGradParameters(X, G2, 1.0)
    both = lyr.gradWeight.clone()
    lyr.resetAcc()
    lyr.accGradParameters(X, G1 + G2, 1.0)
    assert rel(cpu(lyr.gradWeight), cpu(both)) < (1e-5 if precision == "fp32" else 1e-2)
    # invariant 3
    lyr.compute_prior()
    lc = lyr.calc_lc_sum()
    var = torch.exp(lyr.lvars.double())
    expect = float((0.5 * torch.log(lyr.var_hat / var)).sum() / opt["B"])
    assert abs(lc - expect) < 1e-3 * abs(expect) + 1e-6


def test_error_behaviour(ctx):
    import vbnn_b200
    opt = vbnn_b200.default_opt(S=1)
    lyr = vbnn_b200.VBLinear(8, 4, opt, ctx)
    with pytest.raises(vbnn_b200.VbnnError):
        lyr.updateOutput(torch.zeros(3, 9, device="cuda"))          # Torch would raise a size mismatch
    with pytest.raises(vbnn_b200.VbnnError):
        lyr.sample(eps=torch.zeros(5, 5, device="cuda"))
    # ragged / tiny sizes run (N = 1, I not a multiple of 4)
    lyr2 = vbnn_b200.VBLinear(10, 3, opt, ctx)
    lyr2.sample(sample_idx=0)
    y = lyr2.updateOutput(torch.ones(1, 10, device="cuda"))
    assert y.shape == (1, 3) and bool(torch.isfinite(y).all())
    assert lyr2.snr_prune_count(1e9) == 30


def test_layer_on_legacy_default_stream_with_bound_storage():
    """The layer-level drop-in as the Lua shim uses it: a context on the LEGACY DEFAULT stream
    (vbnn_ctx_create_ex, VBNN_CTX_STREAM_LEGACY_DEFAULT -- where cutorch's nn.ReLU / criterion run) and
    weight / bias / gradWeight / gradBias adopted from caller-owned flat storage (vbnn_layer_bind, the
    getParameters() re-flattening of mlp.lua:37), interleaved with default-stream torch work and no explicit
    synchronisation; against the committed fp64 fixture."""
    import vbnn_b200
    from vbnn_b200 import _lib as L
    g = np.load(os.path.join(GOLD, "vblinear_weight.npz"))
    I, Ol, S = int(g["I"]), int(g["O"]), int(g["S"])
    old_stream = torch.cuda.current_stream()
    lctx = vbnn_b200.Context(0, seed=5, stream_mode=L.STREAM_LEGACY_DEFAULT)
    try:
        opt = vbnn_b200.default_opt(B=float(g["B"]), S=S, mu_init=1, var_init=0.01, log=False)
        lyr = vbnn_b200.VBLinear(I, Ol, opt, lctx)
        lyr.set(L.BUF_MEANS, g["means0"]); lyr.set(L.BUF_LVARS, g["lvars0"])
        flat_p = torch.zeros(Ol * I + Ol, device="cuda")                 # getParameters(): {weight, bias}
        flat_p[Ol * I:] = torch.from_numpy(g["bias0"]).float().cuda()    # the tensor is the parameter: adopted as is
        flat_g = torch.full((Ol * I + Ol,), 7.0, device="cuda")          # garbage until gradParameters:zero()
        for which, t in ((L.BUF_WEIGHT, flat_p[:Ol * I]), (L.BUF_BIAS, flat_p[Ol * I:]),
                         (L.BUF_GRAD_WEIGHT, flat_g[:Ol * I]), (L.BUF_GRAD_BIAS, flat_g[Ol * I:])):
            L.check(L.lib().vbnn_layer_bind(lyr.handle, which, C.c_void_p(t.data_ptr())))
        assert lyr.gradWeight.data_ptr() == flat_g.data_ptr()            # the module's view follows the flat storage
        assert rel(cpu(lyr.bias), g["bias0"]) < 1e-7                     # the caller's values are the layer's now
        lyr.compute_prior()
        with torch.cuda.stream(torch.cuda.default_stream()):
            flat_g.zero_()                                               # gradParameters:zero() (mlp.lua:63), stream 0
            lyr.resetAcc()
            X = torch.from_numpy(g["X"]).float().cuda()
            for s in range(S):
                G = torch.from_numpy(g[f"G{s}"]).float().cuda()
                lyr.sample(eps=torch.from_numpy(g[f"noise{s}"]).float().cuda(), sample_idx=s)
                X2 = (X * 3.0) / 3.0 + 0.0                               # default-stream producer of the input
                Y = lyr.updateOutput(X2)
                Yc = Y * 1.0                                             # default-stream consumer of the output
                dX = lyr.updateGradInput(X2, G * 1.0)
                lyr.accGradParameters(X2, G, 1.0)
                assert rel(cpu(Yc), g[f"Y{s}"]) < 1e-5 and rel(cpu(dX), g[f"dX{s}"]) < 2e-5
            assert rel(cpu(flat_g[:Ol * I]), g["gradWeight"].ravel()) < 1e-5     # accumulated IN the flat storage
            assert rel(cpu(flat_g[Ol * I:]), g["gradBias"]) < 1e-5
            assert rel(cpu(flat_p[:Ol * I].view(Ol, I)), cpu(lyr._view(L.BUF_WEIGHT))) == 0.0
            lyr.update(opt)
            assert rel(cpu(flat_p[Ol * I:]), g["bias1"]) < 1e-5          # SGD on the bound bias (VBLinear.lua:125-128)
            assert rel(cpu(lyr.means), g["means1"]) < 1e-5
        for which in (L.BUF_WEIGHT, L.BUF_BIAS, L.BUF_GRAD_WEIGHT, L.BUF_GRAD_BIAS):
            L.check(L.lib().vbnn_layer_bind(lyr.handle, which, None))    # hand back before the tensors die
        assert lyr.gradWeight.data_ptr() != flat_g.data_ptr()
        assert rel(cpu(lyr.bias), g["bias1"]) < 1e-5
        del lyr
    finally:
        lctx.close()
        torch.cuda.set_stream(old_stream)


@pytest.mark.parametrize("force_pair", [False, True])
@pytest.mark.parametrize("M,N,K,batch", [(100, 10, 784, 1), (300, 520, 1000, 2), (540, 600, 552, 1)])
def test_gemm_writes_stay_in_bounds(ctx, M, N, K, batch, force_pair):
    """compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer.md), so out-of-bounds writes are looked
    for directly: the output of the tcgen05 GEMM sits in the middle of a canary-filled allocation with a row pitch
    wider than N; ragged tiles (TMA zero fill on the loads, predicated stores in the epilogue) must leave every
    canary -- the guard rows above and below, and the pitch columns beside every row -- untouched."""
    import vbnn_b200
    from vbnn_b200 import _lib as L
    old = [vbnn_b200.knob("tc_bn", 256 if force_pair else L.KNOB_DEFAULT), vbnn_b200.knob("tc_cg", 2 if force_pair else L.KNOB_DEFAULT)]
    try:
        g = torch.Generator().manual_seed(M + N + K)
        r8 = lambda v: (v + 7) // 8 * 8
        A = torch.randn(batch, M, K, generator=g).bfloat16(); B = torch.randn(batch, N, K, generator=g).bfloat16()
        lda, ldb = r8(K), r8(K)
        Ad = torch.zeros(batch, M, lda, dtype=torch.bfloat16); Ad[:, :, :K] = A
        Bd = torch.zeros(batch, N, ldb, dtype=torch.bfloat16); Bd[:, :, :K] = B
        Ad, Bd = Ad.cuda(), Bd.cuda()
        ldd, guard = N + 12, 64                                           # ldd % 4 == 0 keeps the staged (coalesced) epilogue
        CANARY = 1234.5
        buf = torch.full((batch, M + 2 * guard, ldd), CANARY, device="cuda")
        D = buf[:, guard:guard + M, :]
        L.check(L.lib().vbnn_gemm_bf16(ctx.handle, C.c_void_p(Ad.data_ptr()), lda, 1, C.c_void_p(Bd.data_ptr()), ldb, 1,
                                       C.c_void_p(D[0].data_ptr()), ldd, M, N, K, batch, M * lda, N * ldb, (M + 2 * guard) * ldd))
        ctx.synchronize()
        out = buf.cpu().numpy()
        assert (out[:, :guard, :] == CANARY).all() and (out[:, guard + M:, :] == CANARY).all(), "guard rows overwritten"
        assert (out[:, guard:guard + M, N:] == CANARY).all(), "pitch columns overwritten"
        ref = torch.matmul(A.double(), B.double().transpose(1, 2)).numpy()
        assert rel(out[:, guard:guard + M, :N], ref) < 1e-5
    finally:
        vbnn_b200.knob("tc_bn", old[0]); vbnn_b200.knob("tc_cg", old[1])
