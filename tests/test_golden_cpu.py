"""The committed golden fixtures are what the oracle produces (fp64) and what its fp32 mode --
the timed 'reference CPU path' -- reproduces to single precision."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402
from oracle import vbnn_oracle as O  # noqa: E402


@pytest.mark.parametrize("name,fn,args", [
    ("vblinear_weight.npz", make_golden.layer_case, ("weight",)),
    ("vblinear_local.npz", make_golden.layer_case, ("local",)),
    ("mlp_weight.npz", make_golden.mlp_case, ()),
])
def test_fixture_is_reproducible(name, fn, args):
    g = np.load(os.path.join(HERE, "golden", name))
    fresh = fn(*args)
    for k in g.files:
        a, b = g[k], np.asarray(fresh[k])
        if a.dtype.kind in "US":
            assert list(a) == list(b)
        else:
            assert np.allclose(a, b, rtol=1e-12, atol=1e-14), k


def test_fp32_oracle_matches_fp64_golden():
    g = np.load(os.path.join(HERE, "golden", "vblinear_weight.npz"))
    opt = O.default_opt(B=float(g["B"]), S=int(g["S"]), mu_init=1, var_init=0.01)
    lyr = O.VBLinearOracle(int(g["I"]), int(g["O"]), opt, torch.float32)
    lyr.means.copy_(torch.from_numpy(g["means0"])); lyr.lvars.copy_(torch.from_numpy(g["lvars0"]))
    lyr.bias.copy_(torch.from_numpy(g["bias0"])); lyr.compute_prior()
    lyr.resetAcc(); lyr.gradWeight.zero_(); lyr.gradBias.zero_()
    X = torch.from_numpy(g["X"]).float()
    for s in range(int(g["S"])):
        lyr.sample(torch.from_numpy(g[f"noise{s}"]))
        Y = lyr.updateOutput(X)
        assert np.allclose(Y.numpy(), g[f"Y{s}"], rtol=1e-4, atol=1e-5)
        lyr.accGradParameters(X, torch.from_numpy(g[f"G{s}"]).float())
    assert np.allclose(lyr.gradSum.numpy(), g["gradSum"], rtol=1e-4, atol=1e-6)
    lyr.update(opt)
    assert np.allclose(lyr.means.numpy(), g["means1"], rtol=1e-4, atol=1e-6)
    assert np.allclose(lyr.lvars.numpy(), g["lvars1"], rtol=1e-4, atol=1e-5)
