"""CPU oracle for the VBLinear hot path of louissmit/VBNN.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path (``vbnn_b200`` +
``libvbnn.so``) never routes through this file and has no CPU fallback.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 and Torch7/LuaJIT cannot run in this image, so this
restatement is pinned only by (a) a line-by-line reading of the reference
files cited below and (b) the six analytic invariants of SURVEY.md section 4,
checked in ``tests/test_oracle.py``.

What is restated (reference file:line):
  * ``nn.VBLinear``            -- VBLinear.lua:9-166
  * inherited ``nn.Linear``    -- un-vendored torch/nn (c. 2015): output = X W^T + b,
                                  gradInput = G W, gradWeight += s G^T X, gradBias += s G^T 1
  * ``optim.sgd``/``optim.adam`` -- un-vendored torch/optim (c. 2015), call sites
                                  VBLinear.lua:125,135,140 and mlp.lua:120
  * ``mlp`` net object         -- mlp.lua:7-142
  * one training minibatch     -- main.lua:19-51
  * accuracy                   -- utils.lua:11-27
  * local reparameterisation   -- NOT in the reference;
                                  formulas restated from the north star and verified
                                  against torch fp64 autograd in tests/test_oracle.py

``operand_round`` (default: identity) restates the tensor-core path's operand staging: the
function is applied exactly where libvbnn's VBNN_PREC_BF16 mode rounds a GEMM operand to bf16
(sampled weights, mu / sigma^2 copies, activations, X^2, R, back-propagated gradients); every
accumulation stays in the oracle's dtype, as the GPU accumulates in fp32.  With it the bf16 GPU
path is checked to ~1e-3; without it the same tests state the bf16-vs-fp64 gap separately.

All arithmetic is torch-CPU in the dtype given at construction (float64 for the
checker, float32 + 8 threads for the timed "reference CPU path").  Noise is
always injectable so that the CUDA path and the oracle see identical epsilon.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch


# ----------------------------------------------------------------------------
# config.lua restated as a plain dict factory (config.lua:1-68)
# ----------------------------------------------------------------------------
def default_opt(**over):
    opt = dict(
        threads=8,                      # config.lua:5
        type="vb",                      # config.lua:8
        cuda=True,                      # config.lua:10
        batchSize=1,                    # config.lua:11
        testBatchSize=100,              # config.lua:12
        trainSize=100, testSize=1000,   # config.lua:15-16
        classes=list("0123456789"),     # config.lua:17
        geometry=(28, 28),              # config.lua:18
        input_size=28 * 28,             # config.lua:19
        B=1000000.0,                    # config.lua:30
        hidden=[10],                    # config.lua:31
        S=30,                           # config.lua:32
        testSamples=30,                 # config.lua:33
        quicktest=False,                # config.lua:34 (commented out)
        log=True,                       # config.lua:35
        mu_init=0,                      # config.lua:43
        var_init=0.001,                 # config.lua:44
        msr_init=False,                 # config.lua:45 (commented out)
        state=dict(learningRate=0.001),         # config.lua:51-54
        varState=dict(learningRate=0.05),       # config.lua:55-59
        meanState=dict(learningRate=0.0001),    # config.lua:60-64
        # keys added by the new build ("Config / flags")
        reparam="weight",               # 'weight' (reference) or 'local' (A12)
        strict_reference=True,          # reproduce quirk Q1 (stale sigma in sample)
        vb_output=False,                # convnet.lua:30 uses a VBLinear output layer (Q8)
    )
    opt.update(over)
    return opt


# ----------------------------------------------------------------------------
# optim.sgd / optim.adam (torch/optim, un-vendored).  Call sites:
# VBLinear.lua:125-128 (bias), :135-138 (means), :140-143 (lvars), mlp.lua:120-123.
# Both return (x, applied_step) -- quirk Q5: the reference reads a third return
# value "update"; we define it as x_new - x_old.
# ----------------------------------------------------------------------------
def round_bf16(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bfloat16 and back: the operand_round of VBNN_PREC_BF16."""
    return x.to(torch.bfloat16).to(x.dtype)


def _ident(x):
    return x


def optim_sgd(x: torch.Tensor, dfdx: torch.Tensor, state: dict):
    lr = state.get("learningRate", 1e-3)
    lrd = state.get("learningRateDecay", 0.0)
    n = state.get("evalCounter", 0)
    clr = lr / (1 + n * lrd)
    step = -clr * dfdx
    x.add_(step)
    state["evalCounter"] = n + 1
    return x, step


def optim_adam(x: torch.Tensor, dfdx: torch.Tensor, state: dict):
    lr = state.get("learningRate", 1e-3)
    beta1 = state.get("beta1", 0.9)
    beta2 = state.get("beta2", 0.999)
    eps = state.get("epsilon", 1e-8)
    if "t" not in state:
        state["t"] = 0
        state["m"] = torch.zeros_like(dfdx)
        state["v"] = torch.zeros_like(dfdx)
    state["t"] += 1
    t = state["t"]
    state["m"].mul_(beta1).add_(dfdx, alpha=1 - beta1)
    state["v"].mul_(beta2).addcmul_(dfdx, dfdx, value=1 - beta2)
    denom = state["v"].sqrt().add_(eps)
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    step_size = lr * math.sqrt(bc2) / bc1
    step = -step_size * state["m"] / denom
    x.add_(step)
    return x, step


# ----------------------------------------------------------------------------
# nn.VBLinear (VBLinear.lua:7-166) on top of nn.Linear
# ----------------------------------------------------------------------------
class VBLinearOracle:
    def __init__(self, inputSize: int, outputSize: int, opt: dict,
                 dtype=torch.float64, rng: Optional[np.random.RandomState] = None, operand_round=None):
        # VBLinear.lua:9-47
        self.opt = opt
        self.dtype = dtype
        self.q = operand_round or _ident
        self.I, self.O = inputSize, outputSize
        rng = rng or np.random.RandomState(3)               # config.lua:40 manualSeed(3)
        self.rng = rng
        stdv = 1.0 / math.sqrt(inputSize)                   # nn.Linear:reset()
        self.weight = torch.from_numpy(rng.uniform(-stdv, stdv, (outputSize, inputSize))).to(dtype)
        self.bias = torch.zeros(outputSize, dtype=dtype)    # VBLinear.lua:13
        self.gradWeight = torch.zeros(outputSize, inputSize, dtype=dtype)
        self.gradBias = torch.zeros(outputSize, dtype=dtype)
        self.var_init = opt["var_init"]                     # :12
        if opt.get("msr_init"):
            self.var_init = 2.0 / inputSize                 # :14-16
        self.lvars = torch.full((outputSize, inputSize), math.log(self.var_init), dtype=dtype)  # :18
        self.gradSum = torch.zeros(outputSize, inputSize, dtype=dtype)                          # :20
        self.W = outputSize * inputSize                     # :21
        if opt["mu_init"] == 0:                             # :22-29
            self.means = torch.zeros(outputSize, inputSize, dtype=dtype)
        else:
            std_init = math.sqrt(self.var_init)
            self.means = torch.from_numpy(rng.normal(0.0, std_init, (outputSize, inputSize))).to(dtype)
        self.biasState = dict(opt["state"])                 # :31-33 (shallow copies)
        self.meanState = dict(opt["meanState"])
        self.varState = dict(opt["varState"])
        self.e = torch.zeros(outputSize, inputSize, dtype=dtype)  # :37
        # local-reparameterisation scratch (A12; not in the reference)
        self.lrt = opt.get("reparam", "weight") == "local"
        self.zeta = None
        self._R = None
        self.compute_prior()                                # :46

    # ---- VBLinear.lua:49-64 -------------------------------------------------
    def sample(self, eps: Optional[torch.Tensor] = None):
        if eps is None:
            # randomkit.normal == numpy legacy RandomState.normal (rk_gauss, polar method)
            eps = torch.from_numpy(self.rng.normal(0.0, 1.0, (self.O, self.I)))
        self.e = eps.to(self.dtype).reshape(self.O, self.I)
        if self.opt.get("strict_reference", True):
            stdv = self.stdv                                # :59 -- cached by compute_prior (quirk Q1)
        else:
            stdv = torch.exp(0.5 * self.lvars)
        w = self.means + stdv * self.e                      # :59
        self.weight.copy_(self.q(w))                        # :63

    # ---- VBLinear.lua:77-88 -------------------------------------------------
    def compute_prior(self):
        self.vars = torch.exp(self.lvars)                   # :78
        self.stdv = torch.sqrt(self.vars)                   # :79
        self.mu_hat = 0.0                                   # :81
        self.mu_sqe = (self.means - self.mu_hat).pow(2)     # :82
        self.var_hat = float((1.0 / self.W) * torch.sum(self.vars + self.mu_sqe))  # :86
        return self.mu_hat, self.var_hat

    # ---- VBLinear.lua:90-98 -------------------------------------------------
    def compute_mugrads(self, opt):
        lcg = (self.means - self.mu_hat) / (opt["B"] * self.var_hat)      # :91
        return self.gradWeight.div_(opt["S"]), lcg                       # :92 (in place)

    def compute_vargrads(self, opt):
        lcg = (-self.vars.pow(-1) + 1.0 / self.var_hat) / (2 * opt["B"])   # :96
        if self.lrt:
            # A12: gradSum holds H^T X^2 with H = G*zeta/(2 sqrt V); d/d(log s2) = (.)*s2
            leg = self.gradSum.div_(opt["S"]).mul_(self.vars)
        else:
            leg = self.gradSum.div_(2 * opt["S"]).mul_(self.stdv)          # :97 (in place)
        return leg, lcg * self.vars                                        # :97

    # ---- VBLinear.lua:99-107 ------------------------------------------------
    def calc_lc(self, opt):
        LCfirst = -torch.log(torch.sqrt(self.vars)) + math.log(math.sqrt(self.var_hat))    # :100
        LCsecond = (self.mu_sqe + (self.vars - self.var_hat)) / (2 * self.var_hat)         # :101
        return (LCfirst + LCsecond) * (1.0 / opt["B"])                                     # :102

    def clamp_to_map(self):
        self.weight.copy_(self.q(self.means))               # :106

    # ---- inherited nn.Linear (un-vendored) ----------------------------------
    def updateOutput(self, input: torch.Tensor, zeta: Optional[torch.Tensor] = None):
        q = self.q
        input = q(input)
        self.input = input
        if self.lrt and getattr(self, "_map", False):
            self.output = input @ q(self.means).t() + self.bias
            return self.output
        if self.lrt:
            # A12 forward: M = X mu^T + b, V = X^2 (s2)^T, Y = M + sqrt(V) zeta
            s2 = q(torch.exp(self.lvars))
            M = input @ q(self.means).t() + self.bias
            V = q(input * input) @ s2.t()
            if zeta is None:
                zeta = torch.from_numpy(self.rng.normal(0.0, 1.0, tuple(M.shape)))
            self.zeta = zeta.to(self.dtype)
            sq = torch.sqrt(V)
            self._R = q(self.zeta / (2.0 * sq))
            self.output = M + sq * self.zeta
            return self.output
        self.output = input @ self.weight.t() + self.bias   # nn.Linear:updateOutput
        return self.output

    def updateGradInput(self, input, gradOutput):
        q = self.q
        input, gradOutput = q(input), q(gradOutput)
        if self.lrt and not getattr(self, "_map", False):
            s2 = q(torch.exp(self.lvars))
            H = q(gradOutput * self._R)
            self.gradInput = gradOutput @ q(self.means) + 2.0 * input * (H @ s2)
            return self.gradInput
        self.gradInput = gradOutput @ self.weight           # nn.Linear:updateGradInput
        return self.gradInput

    # ---- VBLinear.lua:112-118 -----------------------------------------------
    def accGradParameters(self, input, gradOutput, scale=1.0):
        q = self.q
        input, gradOutput = q(input), q(gradOutput)
        if self.lrt and not getattr(self, "_map", False):
            H = q(gradOutput * self._R)
            self.gradWeight.add_(scale * (gradOutput.t() @ input))
            self.gradBias.add_(scale * gradOutput.sum(0))
            self.gradSum.add_(H.t() @ q(input * input))
            return
        self.gradWeight.add_(scale * (gradOutput.t() @ input))   # :113 parent
        self.gradBias.add_(scale * gradOutput.sum(0))            # :113 parent
        grad = gradOutput.t() @ input                            # :114 (redundant GEMM, Q2)
        self.gradSum.add_(grad * self.e)                         # :115 (ignores scale)

    def resetAcc(self):
        self.gradSum.zero_()                                # :121

    # ---- VBLinear.lua:124-166 -----------------------------------------------
    def update(self, opt):
        _, bias_step = optim_sgd(self.bias, self.gradBias, self.biasState)       # :125-128
        self.compute_prior()                                                     # :130
        mleg, mlcg = self.compute_mugrads(opt)                                   # :131
        mugrad = mleg + mlcg                                                     # :132
        vleg, vlcg = self.compute_vargrads(opt)                                  # :133
        vgrad = vleg + vlcg                                                      # :134
        x, mu_step = optim_adam(self.means, mugrad, self.meanState)              # :135-138
        mu_normratio = float(torch.norm(mu_step) / torch.norm(x))               # :139
        x, var_step = optim_adam(self.lvars, vgrad, self.varState)               # :140-143
        var_normratio = float(torch.norm(var_step) / torch.norm(x))             # :144
        vars_ = torch.exp(self.lvars)                                            # :145
        nl, nm = float(self.lvars.norm()), float(self.means.norm())
        # the 14 diagnostics of VBLinear.lua:150-163, in file order
        self.stats = {
            "vlc grad": float(vlcg.norm()) / nl,
            "vle grad": float(vleg.norm()) / nl,
            "mlc grad": float(mlcg.norm()) / nm if nm > 0 else float("inf"),
            "mle grad": float(mleg.norm()) / nm if nm > 0 else float("inf"),
            "min variance": float(vars_.min()),
            "max variance": float(vars_.max()),
            "mean variance": float(vars_.mean()),
            "var hat": self.var_hat,
            "mean means": float(self.means.mean()),
            "std means": float(self.means.std()),
            "min. means": float(self.means.min()),
            "max. means": float(self.means.max()),
            "mu normratio": mu_normratio,
            "var normratio": var_normratio,
        }
        return self.stats

    # nn.Module protocol
    def parameters(self):
        return [self.weight, self.bias], [self.gradWeight, self.gradBias]

    def forward(self, input, zeta=None):
        return self.updateOutput(input, zeta)

    def backward(self, input, gradOutput, scale=1.0):
        gi = self.updateGradInput(input, gradOutput)
        self.accGradParameters(input, gradOutput, scale)
        return gi


class LinearOracle:
    """Plain nn.Linear (mlp.lua:29 output layer; quirk Q8)."""

    def __init__(self, inputSize, outputSize, dtype=torch.float64, rng=None, operand_round=None):
        rng = rng or np.random.RandomState(3)
        self.q = operand_round or _ident
        stdv = 1.0 / math.sqrt(inputSize)
        self.weight = torch.from_numpy(rng.uniform(-stdv, stdv, (outputSize, inputSize))).to(dtype)
        self.bias = torch.from_numpy(rng.uniform(-stdv, stdv, (outputSize,))).to(dtype)
        self.gradWeight = torch.zeros_like(self.weight)
        self.gradBias = torch.zeros_like(self.bias)

    def updateOutput(self, input):
        self.output = self.q(input) @ self.q(self.weight).t() + self.bias
        return self.output

    def updateGradInput(self, input, gradOutput):
        self.gradInput = self.q(gradOutput) @ self.q(self.weight)
        return self.gradInput

    def accGradParameters(self, input, gradOutput, scale=1.0):
        input, gradOutput = self.q(input), self.q(gradOutput)
        self.gradWeight.add_(scale * (gradOutput.t() @ input))
        self.gradBias.add_(scale * gradOutput.sum(0))


# ----------------------------------------------------------------------------
# criterion + accuracy
# ----------------------------------------------------------------------------
def log_softmax(x):                     # nn.LogSoftMax (mlp.lua:30)
    m = x.max(dim=1, keepdim=True).values
    z = x - m
    return z - torch.log(torch.exp(z).sum(dim=1, keepdim=True))


def class_nll_forward(logp, targets1):  # nn.ClassNLLCriterion, sizeAverage (mlp.lua:32)
    n = logp.shape[0]
    idx = (targets1.long() - 1)         # targets are 1-based (data.lua:16)
    return float(-logp[torch.arange(n), idx].sum() / n)


def class_nll_backward(logp, targets1):
    n = logp.shape[0]
    g = torch.zeros_like(logp)
    g[torch.arange(n), targets1.long() - 1] = -1.0 / n
    return g


def log_softmax_backward(logp, grad_out):
    return grad_out - torch.exp(logp) * grad_out.sum(dim=1, keepdim=True)


def get_accuracy(outputs, targets1):    # utils.lua:11-27 (percent)
    idx = outputs.argmax(dim=1) + 1
    return float((idx == targets1.long()).sum()) / outputs.shape[0] * 100.0


# ----------------------------------------------------------------------------
# mlp.lua net object
# ----------------------------------------------------------------------------
class MLPOracle:
    """mlp.lua:7-142.  Reshape -> [VBLinear -> ReLU] x H -> Linear -> LogSoftMax."""

    def __init__(self, opt: dict, dtype=torch.float64, seed: int = 3, operand_round=None):
        self.buildModel(opt, dtype, seed, operand_round)

    def buildModel(self, opt, dtype=torch.float64, seed=3, operand_round=None):      # mlp.lua:7-60
        self.opt = opt
        self.dtype = dtype
        self.q = operand_round or _ident
        rng = np.random.RandomState(seed)
        self.rng = rng
        sizes = [opt["input_size"]] + list(opt["hidden"])
        self.vb: List[VBLinearOracle] = []
        for i in range(1, len(sizes)):                            # :13-28
            self.vb.append(VBLinearOracle(sizes[i - 1], sizes[i], opt, dtype, rng, operand_round))
        C = len(opt["classes"])
        if opt.get("vb_output"):                                  # convnet.lua:30 (Q8)
            self.out = VBLinearOracle(sizes[-1], C, opt, dtype, rng, operand_round)
            self.vb_all = self.vb + [self.out]
        else:
            self.out = LinearOracle(sizes[-1], C, dtype, rng, operand_round)     # :29
            self.vb_all = list(self.vb)
        # mlp.lua:47-55: re-init every Linear's *weight* ~ N(0, sqrt(2/fan_in)), bias = 0.
        # For VB layers this touches .weight (overwritten by the next sample()), not .means.
        for lyr in self.vb + [self.out]:
            fan_in = lyr.weight.shape[1]
            lyr.bias.zero_()
            lyr.weight.copy_(torch.from_numpy(
                rng.normal(0.0, math.sqrt(2.0 / fan_in), tuple(lyr.weight.shape))).to(dtype))
        self.state = dict(opt["state"])                           # :57
        return self

    def resetGradients(self):                                     # mlp.lua:62-67
        for lyr in self.vb + [self.out]:
            lyr.gradWeight.zero_()
            lyr.gradBias.zero_()
        for lyr in self.vb_all:
            lyr.resetAcc()

    def sample(self, eps_list: Optional[Sequence[torch.Tensor]] = None):  # mlp.lua:69-74
        for k, lyr in enumerate(self.vb_all):
            if lyr.lrt:
                continue
            lyr.sample(None if eps_list is None else eps_list[k])

    def run(self, inputs, targets, zeta_list=None, backward=True):        # mlp.lua:76-84
        q = self.q
        x = q(inputs.reshape(inputs.shape[0], -1).to(self.dtype))         # nn.Reshape (:12)
        acts = [x]
        for k, lyr in enumerate(self.vb):
            z = None if zeta_list is None else zeta_list[k]
            y = lyr.updateOutput(acts[-1], z) if lyr.lrt else lyr.updateOutput(acts[-1])
            acts.append(q(torch.clamp(y, min=0)))                         # nn.ReLU (:19,27)
        if isinstance(self.out, VBLinearOracle) and self.out.lrt:
            z = None if zeta_list is None else zeta_list[len(self.vb)]
            logits = self.out.updateOutput(acts[-1], z)
        else:
            logits = self.out.updateOutput(acts[-1])
        logp = log_softmax(logits)                                        # :30
        self.outputs = logp
        if backward:
            df_do = class_nll_backward(logp, targets)                     # :78
            g = q(log_softmax_backward(logp, df_do))                      # :79 model:backward
            g_in = self.out.updateGradInput(acts[-1], g)
            self.out.accGradParameters(acts[-1], g, 1.0)
            for k in range(len(self.vb) - 1, -1, -1):
                g = q(g_in * (acts[k + 1] > 0).to(self.dtype))            # ReLU backward
                g_in = self.vb[k].updateGradInput(acts[k], g)             # computed even for k=0
                self.vb[k].accGradParameters(acts[k], g, 1.0)
        error = class_nll_forward(logp, targets)                          # :80
        accuracy = get_accuracy(logp, targets)                            # :82
        return error, accuracy

    def test(self, inputs, targets, eps_lists=None, zeta_lists=None):     # mlp.lua:86-107
        if self.opt.get("quicktest"):
            for lyr in self.vb_all:
                lyr.clamp_to_map()
                lyr._map = True
            r = self.run(inputs, targets, backward=False)
            for lyr in self.vb_all:
                lyr._map = False
            return r
        err = acc = 0.0
        T = self.opt["testSamples"]
        for t in range(T):
            self.sample(None if eps_lists is None else eps_lists[t])
            e, a = self.run(inputs, targets, None if zeta_lists is None else zeta_lists[t],
                            backward=False)
            err += e
            acc += a
        return err / T, acc / T

    def calc_lc(self, opt=None):                                          # mlp.lua:109-115
        opt = opt or self.opt
        return sum(float(l.calc_lc(opt).sum()) for l in self.vb_all)

    def update(self, opt=None):                                           # mlp.lua:117-142
        opt = opt or self.opt
        if not isinstance(self.out, VBLinearOracle):
            # mlp.lua:120-123; quirk Q4: the reference hard-codes a 110-element slice, the
            # *intent* (SGD over the whole output layer, un-normalised by S -- Q3) is restated.
            n = self.state.get("evalCounter", 0)
            lr = self.state.get("learningRate", 1e-3)
            self.out.weight.add_(-lr * self.out.gradWeight)
            self.out.bias.add_(-lr * self.out.gradBias)
            self.state["evalCounter"] = n + 1
        for lyr in self.vb_all:                                           # :138-140
            lyr.update(opt)


def train_minibatch(net: MLPOracle, inputs, targets, opt=None, eps=None, zeta=None):
    """One pass of the closure at main.lua:19-51.  ``eps[s][k]`` / ``zeta[s][k]`` inject the
    noise of MC sample s for VB layer k.  Returns (mean error, mean accuracy) as main.lua:38-39."""
    opt = opt or net.opt
    net.resetGradients()                                                  # main.lua:28
    serr = sacc = 0.0
    for s in range(opt["S"]):                                             # main.lua:32-37
        net.sample(None if eps is None else eps[s])
        e, a = net.run(inputs, targets, None if zeta is None else zeta[s])
        serr += e
        sacc += a
    net.update(opt)                                                       # main.lua:40
    return serr / opt["S"], sacc / opt["S"]


def train_epoch(net: MLPOracle, dataset: dict, opt: dict, order: Sequence[int], eps_fn=None, zeta_fn=None):
    """main:train (main.lua:13-53): B = trainSize / batchSize minibatches visited in the (shuffled,
    utils.lua:90-94) order of start indices `order`; minibatch assembly as data.lua:9-20 (rows
    [t, t + batchSize) of the dataset).  eps_fn(i) / zeta_fn(i) return the injected noise of the i-th
    visited minibatch.  Returns (accuracy / B, error / B) as main.lua:52."""
    B = opt["trainSize"] / opt["batchSize"]                               # main.lua:17
    accuracy = error = 0.0
    for i, t in enumerate(order):                                         # main.lua:18-19
        hi = min(t + opt["batchSize"], opt["trainSize"])
        inputs, targets = dataset["inputs"][t:hi], dataset["targets"][t:hi]   # main.lua:21
        err, acc = train_minibatch(net, inputs, targets, opt,
                                   eps=None if eps_fn is None else eps_fn(i),
                                   zeta=None if zeta_fn is None else zeta_fn(i))
        accuracy += acc                                                   # main.lua:38
        error += err                                                      # main.lua:39
    return accuracy / B, error / B


def snr_prune_mask(means: torch.Tensor, lvars: torch.Tensor, thresh: float = 0.005):
    """mainviz.lua:20-22: pruned = torch.lt(torch.abs(torch.cdiv(means, torch.sqrt(vars))), 0.005), with
    vars = exp(lvars).  Returns (mask, count)."""
    pruned = torch.lt(torch.abs(means / torch.sqrt(torch.exp(lvars))), thresh)
    return pruned, int(pruned.sum())


# ----------------------------------------------------------------------------
# Philox4x32-10 + Box-Muller, the counter layout of libvbnn.so (csrc/philox.cuh).
# Lets tests regenerate on the CPU exactly the epsilon the fused GPU kernels draw.
# ----------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: (n,4) uint32, key: (2,) uint32 -> (n,4) uint32."""
    c = ctr.astype(np.uint32).copy()
    k0 = np.uint32(key[0])
    k1 = np.uint32(key[1])
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c[:, 0].astype(np.uint64)
            p1 = _PHILOX_M1 * c[:, 2].astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            n0 = hi1 ^ c[:, 1] ^ k0
            n2 = hi0 ^ c[:, 3] ^ k1
            c = np.stack([n0, lo1, n2, lo0], axis=1)
            k0 = np.uint32(k0 + _PHILOX_W0)
            k1 = np.uint32(k1 + _PHILOX_W1)
    return c


def philox_normal_matrix(seed: int, step: int, stream: int, sample: int,
                         rows: int, cols: int, row0: int = 0) -> np.ndarray:
    """N(0,1) matrix [rows x cols] exactly as csrc/philox.cuh lays counters out:
    counter = (uint32(row*ceil(cols/4) + col/4), stream, sample, step), key = (seed_lo, seed_hi);
    the 4 outputs of one counter are the normals of columns 4q..4q+3 (Box-Muller on
    pairs (x0,x1) -> (n0,n1), (x2,x3) -> (n2,n3))."""
    q = (cols + 3) // 4
    r = np.arange(row0, row0 + rows, dtype=np.uint64)[:, None]
    c = np.arange(q, dtype=np.uint64)[None, :]
    idx = (r * np.uint64(q) + c).reshape(-1)
    ctr = np.zeros((idx.size, 4), dtype=np.uint32)
    ctr[:, 0] = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ctr[:, 1] = np.uint32(stream)
    ctr[:, 2] = np.uint32(sample)
    ctr[:, 3] = np.uint32(step)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    x = philox4x32_10(ctr, key)
    u = ((x >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    out = np.empty((idx.size, 4), dtype=np.float32)
    for a in (0, 2):
        rad = np.sqrt(np.float32(-2.0) * np.log(u[:, a]))
        ang = np.float32(2.0 * math.pi) * u[:, a + 1]
        out[:, a] = rad * np.cos(ang)
        out[:, a + 1] = rad * np.sin(ang)
    return out.reshape(rows, q * 4)[:, :cols]
