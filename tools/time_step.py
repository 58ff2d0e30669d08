"""Time vbnn_mlp_step for a workload with and without CUDA-graph replay (no per-kernel profiling)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench, vbnn_b200
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
w = dict(bench.WORKLOADS[wl])
ctx = vbnn_b200.Context(0, seed=5)
net, opt = bench.build_net(w, ctx, w["N"])
g = torch.Generator().manual_seed(3)
X = torch.randn(w["N"], w["sizes"][0], generator=g).cuda()
T = torch.randint(1, w["sizes"][-1] + 1, (w["N"],), generator=g).float().cuda()
for _ in range(10):
    net.train_step(X, T, sync=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = net.launch_count()
e0.record()
for _ in range(steps):
    net.train_step(X, T, sync=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"{wl}: {ms:.4f} ms/step  {w['N'] / ms * 1e3:,.0f} samples/s  launches/step {(net.launch_count() - l0) / steps:.1f}  "
      f"graph={'off' if os.environ.get('VBNN_NO_GRAPH') == '1' else 'on'}  err {float(net._res.cpu()[0]):.4f}")
