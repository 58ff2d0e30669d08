"""world_size-2 gloo test of the data-parallel host logic: rank r takes
rows [r*N/G, (r+1)*N/G) of each minibatch with the loss gradient scaled by 1/N_global; the
sum-allreduce of {gradWeight, gradSum, gradBias} equals the single-rank full-batch accumulators,
so the replicated update is identical on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(opt_over=None):
    from oracle import vbnn_oracle as O
    opt = O.default_opt(input_size=10, hidden=[8], classes=list("abc"), S=2, B=20.0, batchSize=8,
                        mu_init=1, var_init=0.01)
    opt.update(opt_over or {})
    net = O.MLPOracle(opt, torch.float64, seed=3)
    return O, opt, net


def _grads(O, net, opt, X, T, eps, scale_rows):
    """accumulators after S x (sample, run) on the row shard X with dLoss scaled to the global batch"""
    net.resetGradients()
    for s in range(opt["S"]):
        net.sample(eps[s])
        x = X
        acts = [x]
        for lyr in net.vb:
            acts.append(torch.clamp(lyr.updateOutput(acts[-1]), min=0))
        logp = O.log_softmax(net.out.updateOutput(acts[-1]))
        g = O.log_softmax_backward(logp, O.class_nll_backward(logp, T)) * scale_rows
        g_in = net.out.updateGradInput(acts[-1], g)
        net.out.accGradParameters(acts[-1], g, 1.0)
        for k in range(len(net.vb) - 1, -1, -1):
            g = g_in * (acts[k + 1] > 0).to(torch.float64)
            g_in = net.vb[k].updateGradInput(acts[k], g)
            net.vb[k].accGradParameters(acts[k], g, 1.0)
    flat = []
    for lyr in net.vb:
        flat += [lyr.gradWeight.reshape(-1), lyr.gradSum.reshape(-1), lyr.gradBias.reshape(-1)]
    flat += [net.out.gradWeight.reshape(-1), net.out.gradBias.reshape(-1)]
    return torch.cat(flat)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O, opt, net = _make()
    rng = np.random.RandomState(0)
    N = 8
    X = torch.from_numpy(rng.randn(N, 10))
    T = torch.from_numpy(rng.randint(1, 4, N).astype(np.float64))
    eps = [[torch.from_numpy(rng.randn(l.O, l.I)) for l in net.vb] for _ in range(opt["S"])]
    lo, hi = rank * N // world, (rank + 1) * N // world
    # local criterion averages over N/world rows; rescale to the global mean (1/N_global)
    local = _grads(O, net, opt, X[lo:hi], T[lo:hi], eps, scale_rows=(hi - lo) / N)
    dist.all_reduce(local, op=dist.ReduceOp.SUM)
    full = _grads(O, net, opt, X, T, eps, scale_rows=1.0)
    ret[rank] = float((local - full).abs().max() / full.abs().max())
    dist.destroy_process_group()


def test_row_sharded_allreduce_equals_full_batch():
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] < 1e-12, ret[r]


def test_philox_rows_are_shard_invariant():
    from oracle import vbnn_oracle as O
    full = O.philox_normal_matrix(5, 3, 1 << 16, 0, rows=16, cols=12)
    parts = [O.philox_normal_matrix(5, 3, 1 << 16, 0, rows=8, cols=12, row0=r * 8) for r in range(2)]
    assert np.array_equal(np.concatenate(parts), full)


# ---- peer mode (csrc/peer.cu): reduce-scatter by row owner, owner-only update, all-gather ----
def _shard(O, G, q):
    """vbnn_peer_shard through the C ABI (pure host code: loads without a GPU)."""
    import ctypes as C
    from vbnn_b200 import _lib as L
    r0, rows = C.c_int(), C.c_int()
    rpo = L.lib().vbnn_peer_shard(O, G, q, C.byref(r0), C.byref(rows))
    return rpo, r0.value, rows.value


def test_peer_shard_layout():
    for O, G in [(4096, 8), (1000, 8), (10, 8), (1200, 2), (100, 4), (33, 2), (1, 8)]:
        covered = []
        rpos = set()
        for q in range(G):
            rpo, r0, rows = _shard(O, G, q)
            rpos.add(rpo)
            assert rpo % 32 == 0 and rpo * G >= O and 0 <= rows <= rpo
            covered += list(range(r0, r0 + rows))
        assert covered == list(range(O)), (O, G)      # a partition, in rank order
        assert len(rpos) == 1


def _peer_worker(rank, world, port, ret, wire_bf16=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import copy
    O, opt, net = _make({"hidden": [40, 36]})
    rng = np.random.RandomState(0)
    N = 8
    X = torch.from_numpy(rng.randn(N, 10))
    T = torch.from_numpy(rng.randint(1, 4, N).astype(np.float64))
    eps = [[torch.from_numpy(rng.randn(l.O, l.I)) for l in net.vb] for _ in range(opt["S"])]
    ref = copy.deepcopy(net)
    lo, hi = rank * N // world, (rank + 1) * N // world
    _grads(O, net, opt, X[lo:hi], T[lo:hi], eps, scale_rows=(hi - lo) / N)
    # "receive slots": every rank's accumulators, summed by the row owner in rank order
    for lyr in net.vb:
        slots = {}
        for name in ("gradWeight", "gradSum", "gradBias"):
            mine = getattr(lyr, name).clone()
            if wire_bf16 and name != "gradBias":
                mine = O.round_bf16(mine)                           # knob peer_wire_bf16: each rank's tile is rounded once
            allr = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            slots[name] = allr
        rpo, r0, rows = _shard(lyr.O, world, rank)
        lyr.gradBias.copy_(sum(slots["gradBias"]))                 # bias: every rank sums all slots
        lyr.gradWeight.zero_(); lyr.gradSum.zero_()                # rows of other owners: never touched
        lyr.gradWeight[r0:r0 + rows] = sum(s[r0:r0 + rows] for s in slots["gradWeight"])
        lyr.gradSum[r0:r0 + rows] = sum(s[r0:r0 + rows] for s in slots["gradSum"])
        before = (lyr.means.clone(), lyr.lvars.clone())
        lyr.update(opt)                                            # sigma_hat^2 from the replicated parameters
        # keep only the owned rows, then all-gather the shards
        for t, old in zip((lyr.means, lyr.lvars), before):
            own = t[r0:r0 + rows].clone()
            t.copy_(old)
            t[r0:r0 + rows] = own
            for q in range(world):
                _, q0, qrows = _shard(lyr.O, world, q)
                buf = t[q0:q0 + qrows].clone()
                dist.broadcast(buf, q)
                t[q0:q0 + qrows] = buf
    # single-process reference: full batch, ordinary update
    _grads(O, ref, opt, X, T, eps, scale_rows=1.0)
    for lyr in ref.vb:
        lyr.update(opt)
    err = 0.0
    for a, b in zip(net.vb, ref.vb):
        if wire_bf16:     # relative Frobenius error: a rounded partial flips the sign of a near-zero gradient now and then,
            err = max(err, float((a.means - b.means).norm() / b.means.norm()),          # and Adam's first step is +-lr
                      float((a.lvars - b.lvars).norm() / b.lvars.norm()), float((a.bias - b.bias).abs().max()))
        else:
            err = max(err, float((a.means - b.means).abs().max()), float((a.lvars - b.lvars).abs().max()),
                      float((a.bias - b.bias).abs().max()))
    ret[rank] = err
    dist.destroy_process_group()


def test_peer_reduce_scatter_sharded_update_allgather_equals_single():
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_peer_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] < 1e-12, ret[r]


def test_peer_exchange_with_bf16_tiles_on_the_wire_is_bounded():
    """knob peer_wire_bf16 (opt-in, strong-scaling regime): every rank's gradient tile is rounded to bf16 once before the
    owner sums the slots in fp32.  Same exchange algebra; the parameters after the update stay within the bound bench.py's
    dp_parity states for that mode (relative Frobenius error <= 1e-3), and all ranks still agree exactly."""
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_peer_worker, args=(world, port, ret, True), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert 0.0 < ret[r] < 1e-3, ret[r]
    assert ret[0] == ret[1]
