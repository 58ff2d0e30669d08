"""Per-layer error of the bf16 fused step vs the fp64 oracle AND vs a torch emulation of the same
bf16 rounding points (to separate inherent bf16 error from bugs)."""
import math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import vbnn_b200
from oracle import vbnn_oracle as O
from test_gpu_mlp import build_pair, oracle_step, rel, cpu

def b(x): return x.to(torch.bfloat16).to(torch.float64)

def emul_weight(ref, X, T, eps, S):
    L = len(ref.vb)
    gW = [torch.zeros_like(l.means) for l in ref.vb] + [torch.zeros_like(ref.out.weight)]
    gS = [torch.zeros_like(l.means) for l in ref.vb]
    for s in range(S):
        Ws = [b(l.means + l.stdv * eps[s][k]) for k, l in enumerate(ref.vb)] + [b(ref.out.weight)]
        acts = [b(X)]
        for j in range(L):
            acts.append(b(torch.clamp(acts[-1] @ Ws[j].t() + ref.vb[j].bias, min=0)))
        logits = acts[-1] @ Ws[-1].t() + ref.out.bias
        p = torch.softmax(logits, 1); p[torch.arange(X.shape[0]), T.long() - 1] -= 1
        G = b(p / X.shape[0])
        for j in range(L, -1, -1):
            D = G.t() @ acts[j]
            gW[j] += D
            if j < L: gS[j] += D * eps[s][j]
            if j > 0: G = b((G @ Ws[j]) * (acts[j] > 0))
    return gW, gS

ctx = vbnn_b200.default_context(0, seed=5)
for sizes, N, S in [([40, 48, 36, 6], 24, 3), ([64, 128, 128, 16], 128, 2), ([256, 512, 512, 10], 256, 1)]:
    for precision in ("fp32", "bf16"):
        ctx.set_step(3)
        net, ref, gopt, oopt = build_pair(ctx, sizes, N, S, 30.0, precision, "weight")
        rng = np.random.RandomState(5)
        Xn = rng.randn(N, sizes[0]); Tn = rng.randint(1, sizes[-1] + 1, N).astype(np.float64)
        step = ctx.get_step()
        noise = [[torch.from_numpy(cpu(net.model[k].draw_noise(step, s)).astype(np.float64)) for k in range(2)] for s in range(S)]
        X, T = torch.from_numpy(Xn), torch.from_numpy(Tn)
        egW, egS = emul_weight(ref, X, T, noise, S)
        err, acc = net.train_step(X.float().cuda(), T.float().cuda())
        rerr, racc, accs = oracle_step(ref, oopt, X, T, eps=noise)
        print(f"{sizes} N={N} S={S} {precision}: err {err:.5f} oracle {rerr:.5f}")
        for k in range(3):
            gl = net.model[k]
            if k < 2:
                gw, gs, gb = accs[k]
                print(f"  layer {k}: gW vs oracle {rel(cpu(gl.gradWeight), gw.numpy()):.4f}  emul-bf16 vs oracle {rel(egW[k].numpy(), gw.numpy()):.4f}  gpu vs emul {rel(cpu(gl.gradWeight), egW[k].numpy()):.4f}"
                      f" | gS vs oracle {rel(cpu(gl.gradSum), gs.numpy()):.4f} emul {rel(egS[k].numpy(), gs.numpy()):.4f} | gb {rel(cpu(gl.gradBias), gb.numpy()):.4f}")
            else:
                print(f"  out layer: gW gpu vs emul {rel(cpu(gl.gradWeight), egW[k].numpy()):.4f}")
        del net
