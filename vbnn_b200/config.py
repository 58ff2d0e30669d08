"""config.lua restated (reference config.lua:1-68): the same keys with the same shipped values,
plus the keys the B200 build adds (reparam, precision, seed, ngpu)."""
from __future__ import annotations

from . import _lib as L


def default_opt(**over):
    opt = dict(
        threads=8, network_to_load="", network_name="exp", type="vb", dataset="mnist",
        cuda=True, batchSize=1, testBatchSize=100,                      # config.lua:5-12
        trainSize=100, testSize=1000, classes=list("0123456789"),       # config.lua:15-17
        geometry=(28, 28), input_size=28 * 28,                          # config.lua:18-19
        plot=True, B=1000000.0, hidden=[10], S=30, testSamples=30,      # config.lua:28-33
        quicktest=False, log=True,                                      # config.lua:34-35
        mu_init=0, var_init=0.001, msr_init=False,                      # config.lua:43-45
        state=dict(learningRate=0.001),                                 # config.lua:51-54
        varState=dict(learningRate=0.05),                               # config.lua:55-59
        meanState=dict(learningRate=0.0001),                            # config.lua:60-64
        # new keys
        reparam="weight",        # 'weight' = reference sampling, 'local' = local reparameterisation
        precision="fp32",        # 'fp32' (exact-parity CUDA-core GEMM) or 'bf16' (tcgen05)
        strict_reference=True,   # reproduce quirks Q1/Q6
        vb_output=False,         # convnet.lua:30: VBLinear output layer instead of nn.Linear
        seed=3,                  # config.lua:40 torch.manualSeed(3)
        ngpu=1,
    )
    opt.update(over)
    return opt


def opts_struct(opt) -> L.VbnnOpts:
    o = L.VbnnOpts()
    L.lib().vbnn_opts_default(o)
    o.var_init = float(opt["var_init"])
    o.msr_init = 1 if opt.get("msr_init") else 0
    o.mu_init = float(opt["mu_init"])
    o.B = float(opt["B"])
    o.S = int(opt["S"])
    o.lr_bias = float(opt["state"]["learningRate"])
    o.lr_mu = float(opt["meanState"]["learningRate"])
    o.lr_var = float(opt["varState"]["learningRate"])
    o.adam_beta1 = float(opt["meanState"].get("beta1", 0.9))
    o.adam_beta2 = float(opt["meanState"].get("beta2", 0.999))
    o.adam_eps = float(opt["meanState"].get("epsilon", 1e-8))
    o.reparam = L.REPARAM_LOCAL if opt.get("reparam", "weight") == "local" else L.REPARAM_WEIGHT
    o.precision = L.PREC_BF16 if opt.get("precision", "fp32") == "bf16" else L.PREC_FP32
    o.strict_reference = 1 if opt.get("strict_reference", True) else 0
    return o
