"""Descriptor probe for the tcgen05 GEMM: runs small structured cases in subprocesses (a trapped
kernel kills its CUDA context, not the probe) and prints an error map that shows WHICH rows /
columns / k-slices are wrong, so a descriptor mistake can be diagnosed from one GPU call."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASE = r'''
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, %(root)r)
import vbnn_b200
from vbnn_b200 import _lib as L
ak, bk, M, N, K, bn = %(ak)d, %(bk)d, %(M)d, %(N)d, %(K)d, %(bn)d
ctx = vbnn_b200.default_context(0, 5)
g = torch.Generator().manual_seed(1)
A = torch.randint(-3, 4, (M, K), generator=g).float()
B = torch.randint(-3, 4, (N, K), generator=g).float()
ref = (A.double() @ B.double().t()).numpy()
r8 = lambda v: (v + 7) // 8 * 8
if ak: lda = r8(K); Ad = torch.zeros(M, lda, dtype=torch.bfloat16); Ad[:, :K] = A
else:  lda = r8(M); Ad = torch.zeros(K, lda, dtype=torch.bfloat16); Ad[:, :M] = A.t()
if bk: ldb = r8(K); Bd = torch.zeros(N, ldb, dtype=torch.bfloat16); Bd[:, :K] = B
else:  ldb = r8(N); Bd = torch.zeros(K, ldb, dtype=torch.bfloat16); Bd[:, :N] = B.t()
Ad, Bd = Ad.cuda(), Bd.cuda()
D = torch.full((M, N), float('nan'), device='cuda')
rc = L.lib().vbnn_gemm_bf16(ctx.handle, C.c_void_p(Ad.data_ptr()), lda, ak, C.c_void_p(Bd.data_ptr()), ldb, bk,
                            C.c_void_p(D.data_ptr()), N, M, N, K, 1, 0, 0, 0)
if rc: print('rc', rc, L.lib().vbnn_last_error()); sys.exit(1)
torch.cuda.synchronize()
d = D.cpu().numpy()
err = np.abs(d - ref)
bad = ~(err < 1e-3)
print('maxerr %%.4g  bad %%d/%%d  nan %%d' %% (np.nanmax(err) if np.isfinite(err).any() else float('nan'), bad.sum(), bad.size, np.isnan(d).sum()))
if bad.any():
    rb, cb = 16, 16
    print('bad fraction per (16-row, 16-col) block:')
    for r in range(0, min(M, 128), rb):
        print(' '.join('%%3d' %% int(100 * bad[r:r+rb, c:c+cb].mean()) for c in range(0, min(N, 256), cb)))
    # does the result match a GEMM over a subset of k-slices?
    for ks in range(0, K, 16):
        sub = (A[:, ks:ks+16].double() @ B[:, ks:ks+16].double().t()).numpy()
        if np.abs(d - sub).max() < 1e-3: print('  == only k-slice', ks)
    print('d[0,:8]  ', d[0, :8]); print('ref[0,:8]', ref[0, :8])
    print('d[:8,0]  ', d[:8, 0]); print('ref[:8,0]', ref[:8, 0])
'''


def main():
    cases = []
    for ak, bk in [(1, 1), (1, 0), (0, 1), (0, 0)]:
        for (M, N, K) in [(128, 128, 16), (128, 256, 64), (256, 256, 256), (300, 520, 200)]:
            for bn, cg in ((128, 1), (256, 1), (256, 2)):
                if N < bn and bn == 256:
                    continue
                cases.append(dict(ak=ak, bk=bk, M=M, N=N, K=K, bn=bn, cg=cg))
    for c in cases:
        env = dict(os.environ, VBNN_TC_BN=str(c["bn"]), VBNN_TC_CG=str(c["cg"]))
        src = CASE % dict(c, root=ROOT)
        try:
            p = subprocess.run([sys.executable, "-c", src], capture_output=True, text=True, timeout=120, env=env)
            out = (p.stdout + p.stderr[-600:]).strip()
        except subprocess.TimeoutExpired:
            out = "TIMEOUT"
        print(f"--- ak={c['ak']} bk={c['bk']} M={c['M']} N={c['N']} K={c['K']} bn={c['bn']} cg={c['cg']}\n{out}", flush=True)


if __name__ == "__main__":
    main()
