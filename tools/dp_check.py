"""Data-parallel parity on real GPUs (run under torchrun, world size G):
every rank holds rows [r*N/G, (r+1)*N/G) of the same global minibatch; after vbnn_mlp_step with the
NCCL allreduce of the gradient arena, parameters must equal those of ONE rank stepping the full
minibatch (identical Philox eps on every rank; LRT zeta indexed by the global row).

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbnn_b200


def build(ctx, sizes, N, S, reparam, precision):
    opt = vbnn_b200.default_opt(input_size=sizes[0], hidden=sizes[1:-1], classes=[str(i) for i in range(sizes[-1])],
                                S=S, B=50.0, batchSize=N, mu_init=1, msr_init=True, reparam=reparam,
                                precision=precision, strict_reference=False, log=False, seed=5)
    net = vbnn_b200.MLP(opt, ctx, max_batch=N)
    net.init_params(seed=4, he_means=True)
    return net


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{lr}"))
    ctx_dp = vbnn_b200.Context(lr, seed=5)

    def bcast(buf):
        t = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{lr}")
        if rank == 0:
            t.copy_(torch.tensor(list(buf), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())
    ctx_dp.init_comm(rank, world, bcast)

    def gather_bytes(blob):
        t = torch.tensor(list(blob), dtype=torch.uint8, device=f"cuda:{lr}")
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [bytes(o.cpu().tolist()) for o in out]
    use_peer = os.environ.get("VBNN_DP", "peer") == "peer"
    ok = True
    for reparam, precision, sizes, N, S in [("weight", "fp32", [64, 96, 80, 10], 64, 2),
                                           ("local", "fp32", [64, 96, 80, 10], 64, 2),
                                           ("local", "bf16", [512, 512, 512, 100], 1024, 1),
                                           ("weight", "bf16", [256, 384, 256, 10], 512, 3)]:
        g = torch.Generator().manual_seed(11)
        X = torch.randn(N, sizes[0], generator=g)
        T = torch.randint(1, sizes[-1] + 1, (N,), generator=g).float()
        n_loc = N // world
        net = build(ctx_dp, sizes, n_loc, S, reparam, precision)
        if use_peer:
            net.enable_peer(gather_bytes)
        ctx_dp.set_step(7)
        dist.barrier()
        for it in range(3):
            net.train_step(X[rank * n_loc:(rank + 1) * n_loc].cuda(), T[rank * n_loc:(rank + 1) * n_loc].cuda())
        if use_peer:
            torch.cuda.synchronize(); dist.barrier()      # every rank's side stream has finished its last shard update
            net.sync_replicas()
            torch.cuda.synchronize(); dist.barrier()
        dp = [m.means.clone() for m in net.model[:-1]] + [net.model[-1].weight.clone()]
        dp_lv = [m.lvars.clone() for m in net.model[:-1]]
        from vbnn_b200 import _lib as VL
        adam_ids = (VL.BUF_ADAM_M_MU, VL.BUF_ADAM_V_MU, VL.BUF_ADAM_M_VAR, VL.BUF_ADAM_V_VAR)
        dp_adam = [[m.get(b) for b in adam_ids] for m in net.model[:-1]]     # needs sync_replicas in peer mode
        # all ranks must hold identical parameters
        for t in dp:
            ref = t.clone(); dist.broadcast(ref, 0)
            ok &= bool(torch.equal(ref, t))
        if rank == 0:
            # single-GPU reference: a context without a communicator (nranks = 1) on the same device
            ctx1 = vbnn_b200.Context(lr, seed=5)
            one = build(ctx1, sizes, N, S, reparam, precision)
            ctx1.set_step(7)
            for it in range(3):
                one.train_step(X.cuda(), T.cuda())
            sg = [m.means for m in one.model[:-1]] + [one.model[-1].weight]
            tol = 2e-4 if precision == "fp32" else 5e-3
            for a, b in zip(dp, sg):
                e = float((a - b).norm() / b.norm())
                ok &= e < tol
                print(f"{reparam}/{precision} ({'peer' if net.peer_active else 'nccl'}): rel err DP vs single = {e:.2e}")
            for a, m in zip(dp_lv, one.model[:-1]):
                e = float((a - m.lvars).norm() / m.lvars.norm())
                ok &= e < tol
            for bufs, m in zip(dp_adam, one.model[:-1]):                       # optimiser state of every shard
                for a, b in zip(bufs, adam_ids):
                    r = m.get(b)
                    e = float((a - r).norm() / max(float(r.norm()), 1e-30))
                    ok &= e < 10 * tol
            assert all(m.t == 3 for m in net.model)
            torch.cuda.set_stream(ctx_dp.stream)
        torch.cuda.synchronize(); dist.barrier()      # nobody still writes into a peer's buffers
        del net
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{lr}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP CHECK", "OK" if int(flag[0]) == 1 else "FAILED")
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
