"""Checkpoint / interop files of the reference, written from libvbnn.so's device state.

  * `safe_save(object, folder, name)`  -- utils.lua:73-80: previous file renamed `<name>.old`, then
    torch.save.  main.lua:181 saves the whole `net` object graph each epoch; here the net's state
    (MLP.state_dict(): means, lvars, bias, Adam m/v/t per layer, output layer, Philox step, opt) is
    saved as a Torch7 table of FloatTensors under the same file name `model`.
  * `save_parameters(net, model_dir)`  -- the files mainviz.lua:11-15 loads and that nothing in the
    reference writes: `<model_dir>/parameters/means`, `<model_dir>/parameters/vars` (sigma^2, not log
    sigma^2: mainviz.lua:16,20 takes sqrt(vars)) and `<model_dir>/opt`, as flat FloatTensors over all
    VB layers (plus `means_<k>` / `vars_<k>` per layer, [O x I]).
  * `load_net(net, folder)`            -- resume, main.lua:146-148.
All files are Torch7 binary serialisation (vbnn_b200/t7.py), readable by `torch.load` in Lua."""
from __future__ import annotations

import os

import numpy as np

from . import _lib as L
from . import t7


def safe_save(obj, folder, name):                                       # utils.lua:73-80
    os.makedirs(folder, exist_ok=True)
    filename = os.path.join(folder, name)
    if os.path.isfile(filename):
        os.replace(filename, filename + ".old")
    t7.save(filename, obj)
    return filename


def _plain_opt(opt):
    out = {}
    for k, v in opt.items():
        if isinstance(v, tuple):
            v = list(v)
        if isinstance(v, (dict, list, str, int, float, bool)) or v is None:
            out[k] = v
    return out


def save_net(net, folder, name="model"):
    """main.lua:181 `u.safe_save(net, opt.network_name, 'model')`."""
    sd = net.state_dict()
    obj = {"opt": _plain_opt(net.opt)}
    for k, v in sd.items():
        obj[k] = v.numpy().astype(np.float32) if hasattr(v, "numpy") else v
    return safe_save(obj, folder, name)


def load_net(net, folder, name="model"):
    """main.lua:146-148: torch.load(opt.network_to_load) and continue training."""
    import torch
    obj = t7.load(os.path.join(folder, name))
    sd = {}
    for k, v in obj.items():
        if k == "opt":
            continue
        sd[k] = torch.from_numpy(v) if isinstance(v, np.ndarray) else v
    net.load_state_dict(sd)
    return obj.get("opt")


def save_parameters(net, model_dir):
    """The inputs of mainviz.lua:11-15."""
    pdir = os.path.join(model_dir, "parameters")
    os.makedirs(pdir, exist_ok=True)
    means, vars_ = [], []
    for k, m in enumerate(net.model):
        if m.kind != L.KIND_VB:
            continue
        mu = m.get(L.BUF_MEANS).numpy().astype(np.float32)
        s2 = np.exp(m.get(L.BUF_LVARS).numpy().astype(np.float64)).astype(np.float32)
        t7.save(os.path.join(pdir, f"means_{k}"), mu)
        t7.save(os.path.join(pdir, f"vars_{k}"), s2)
        means.append(mu.ravel()); vars_.append(s2.ravel())
    t7.save(os.path.join(pdir, "means"), np.concatenate(means))
    t7.save(os.path.join(pdir, "vars"), np.concatenate(vars_))
    t7.save(os.path.join(model_dir, "opt"), _plain_opt(net.opt))
    return pdir


def snr_pruned(means, vars_, thresh=0.005):
    """mainviz.lua:20-24 on the loaded files: mask = |mu| / sqrt(vars) < thresh; returns
    (mask as float32, pruned count, mean of vars, mean of mask .* vars)."""
    pruned = (np.abs(means / np.sqrt(vars_)) < thresh).astype(np.float32)
    return pruned, float(pruned.sum()), float(vars_.mean()), float((pruned * vars_).mean())
