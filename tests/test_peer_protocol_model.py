"""Model check of peer mode's ordering protocol (vbnn_b200/csrc/peer.cu) on the CPU.

The CUDA implementation orders four kinds of cross-GPU accesses with two flag arrays and two per-layer
step counters per rank (mseq on the main stream, sseq on the side stream):

  * dW epilogue of rank r, step t, layer j   WRITES  slot[q][j][r] on every owner q      (NVLink stores)
  * owner update of rank q, step t, layer j  READS   slot[q][j][*], WRITES operands[*][j][q]
  * forward / backward-data of rank r        READS   operands[r][j][*]

  grad_ready[q][j][r] = t   raised on q by r's main stream after its dW_j(t)
  param_ready[r][j][q] = t  raised on r by q's side stream after its pushes of layer j, step t
  main stream, step t, before forward_j : waits param_ready[r][j][q] >= t-1 for all q (PER LAYER, right before
                                         layer j's operands are first read)
  main stream, backward of step t      : the backward-data chain dX_{L-1} .. dX_1 first, then the parameter-gradient
                                         GEMMs dW_0 .. dW_{L-1} in FORWARD order, so the layer the next forward needs
                                         first finishes its exchange first and the tail belongs to the last layer
  side stream, layer j, step t : starts after the own dW_j(t) (event), waits grad_ready[q][j][r] >= t for all r

This test replays that protocol as 2G sequential "streams" under random interleavings and asserts the data
hazards it is meant to exclude: no slot is overwritten before the owner consumed it, no update reads a
slot of the wrong step, no operand shard is replaced while a reader still needs the old version, every
forward sees exactly the previous step's operands, and the system never deadlocks.  It checks the DESIGN,
not the kernels (tests/test_gpu_peer.py and tools/dp_check.py do that on real GPUs)."""
import random

import pytest


class Model:
    def __init__(self, G, L, T):
        self.G, self.L, self.T = G, L, T
        z3 = lambda: [[[0] * G for _ in range(L)] for _ in range(G)]
        self.grad_ready = z3()        # [q][j][r]
        self.param_ready = z3()       # [r][j][q]
        self.slot_version = z3()      # [q][j][r]: step whose gradient tile sits in q's slot from r
        self.slot_read = [[0] * L for _ in range(G)]          # [q][j]: last step whose update consumed the slots
        self.operand_version = z3()   # [r][j][q]: version of owner q's shard of layer j held by rank r
        self.reading_until = [[0] * L for _ in range(G)]      # [r][j]: last step in which r finished reading layer j's operands
        self.dw_done = [[0] * L for _ in range(G)]            # [r][j]: last step whose dW_j was enqueued+finished on r's main

    # ---- the two programs; each yields ("wait", predicate) or ("do", action) -------------------------
    def main_stream(self, r):
        G, L = self.G, self.L
        for t in range(1, self.T + 1):
            for j in range(L):                                     # forward, per-layer wait
                yield ("wait", lambda j=j, t=t: all(self.param_ready[r][j][q] >= t - 1 for q in range(G)))
                def fwd(j=j, t=t):
                    assert all(v == t - 1 for v in self.operand_version[r][j]), ("forward saw stale/new operands", r, j, t)
                yield ("do", fwd)
            for j in range(L - 1, 0, -1):                          # backward-data chain (layer 0's is skipped)
                def dx(j=j, t=t):
                    assert all(v == t - 1 for v in self.operand_version[r][j]), ("backward-data saw wrong operands", r, j, t)
                    self.reading_until[r][j] = t
                yield ("do", dx)
            for j in range(L):                                     # parameter gradients in forward order
                def dw(j=j, t=t):
                    if j == 0:
                        self.reading_until[r][0] = t               # layer 0's operands were last read by its forward
                    for q in range(G):                             # reduce-scatter fused into the epilogue
                        assert self.slot_read[q][j] >= t - 1, ("slot overwritten before the owner consumed it", r, q, j, t)
                        self.slot_version[q][j][r] = t
                    self.dw_done[r][j] = t
                yield ("do", dw)
                def signal(j=j, t=t):
                    for q in range(G):
                        self.grad_ready[q][j][r] = t
                yield ("do", signal)

    def side_stream(self, q):
        G, L = self.G, self.L
        for t in range(1, self.T + 1):
            for j in range(L):
                yield ("wait", lambda j=j, t=t: self.dw_done[q][j] >= t)                       # cudaStreamWaitEvent(ev_dw)
                yield ("wait", lambda j=j, t=t: all(self.grad_ready[q][j][r] >= t for r in range(G)))
                def update(j=j, t=t):
                    assert all(v == t for v in self.slot_version[q][j]), ("update read a slot of the wrong step", q, j, t)
                    self.slot_read[q][j] = t
                yield ("do", update)
                for r in range(G):                                 # all-gather: one push per rank
                    def push(r=r, j=j, t=t):
                        assert self.reading_until[r][j] >= t, ("operands replaced while still being read", q, r, j, t)
                        assert self.operand_version[r][j][q] == t - 1
                        self.operand_version[r][j][q] = t
                    yield ("do", push)
                def signal(j=j, t=t):
                    for r in range(G):
                        self.param_ready[r][j][q] = t
                yield ("do", signal)


def run(G, L, T, seed):
    rng = random.Random(seed)
    m = Model(G, L, T)
    streams = [m.main_stream(r) for r in range(G)] + [m.side_stream(q) for q in range(G)]
    pending = [next(s) for s in streams]          # the instruction each stream is blocked on / about to run
    alive = [True] * len(streams)
    while any(alive):
        ready = [i for i in range(len(streams)) if alive[i] and (pending[i][0] == "do" or pending[i][1]())]
        assert ready, ("deadlock", G, L, T, seed)
        i = rng.choice(ready)
        if pending[i][0] == "do":
            pending[i][1]()
        try:
            pending[i] = next(streams[i])
        except StopIteration:
            alive[i] = False
    for r in range(G):                            # everybody ends with everybody's final operands
        assert all(v == T for j in range(L) for v in m.operand_version[r][j])


@pytest.mark.parametrize("G,L", [(2, 1), (2, 3), (3, 2), (4, 3), (8, 2)])
def test_peer_protocol_has_no_hazard_or_deadlock(G, L):
    for seed in range(60):
        run(G, L, T=4, seed=seed)


def test_model_detects_a_missing_wait():
    """Sanity of the checker itself: without the per-layer waits a fast rank's next forward must trip
    one of the assertions under some interleaving."""
    class Broken(Model):
        def main_stream(self, r):
            for ins in super().main_stream(r):
                if ins[0] == "wait":
                    continue
                yield ins
    tripped = False
    for seed in range(200):
        rng = random.Random(seed)
        m = Broken(2, 2, 3)
        streams = [m.main_stream(r) for r in range(2)] + [m.side_stream(q) for q in range(2)]
        pending = [next(s) for s in streams]
        alive = [True] * 4
        try:
            while any(alive):
                ready = [i for i in range(4) if alive[i] and (pending[i][0] == "do" or pending[i][1]())]
                if not ready:
                    break
                i = rng.choice(ready)
                if pending[i][0] == "do":
                    pending[i][1]()
                try:
                    pending[i] = next(streams[i])
                except StopIteration:
                    alive[i] = False
        except AssertionError:
            tripped = True
            break
    assert tripped
