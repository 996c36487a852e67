"""Isolated GEMM micro-benchmark (for ncu): python tools/bench_gemm.py M N K epi [iters]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theia_b200 import _lib as L

M, N, K = (int(v) for v in sys.argv[1:4])
wgrad = sys.argv[4].startswith("wgrad")
epi = 32 if wgrad else int(sys.argv[4])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
lib = L.lib()
if len(sys.argv) > 6:
    lib.theia_debug_set(7, int(sys.argv[6]))  # 1 = no loads, 2 = no epilogue
if len(sys.argv) > 7:
    lib.theia_debug_set(8, int(sys.argv[7]))  # 1 = single-CTA kernels only
dev = "cuda"
a = torch.randn(M, K, device=dev).to(torch.bfloat16)
b = (0.05 * torch.randn(N, K, device=dev)).to(torch.bfloat16)
bias = torch.randn(N, device=dev)
aux = torch.randn(M, N, device=dev).to(torch.bfloat16)
out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
out2 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
cs = torch.zeros(N, device=dev)
d = L.GemmDesc()
d.M, d.N, d.K, d.A, d.lda, d.B, d.ldb = M, N, K, a.data_ptr(), K, b.data_ptr(), K
d.out, d.ldo, d.out2, d.aux, d.bias, d.colsum = out.data_ptr(), N, out2.data_ptr(), aux.data_ptr(), bias.data_ptr(), cs.data_ptr()
d.epi, d.splits, d.batch_z = epi, 1, 1
if wgrad:  # dW[M=Nout, N=Kin] += dY[K=tokens, Nout]^T X[K=tokens, Kin]
    a = torch.randn(K, M, device=dev).to(torch.bfloat16)
    b = torch.randn(K, N, device=dev).to(torch.bfloat16)
    out = torch.zeros(M, N, dtype=torch.float32, device=dev)
    d.A, d.lda, d.B, d.ldb, d.out, d.ldo = a.data_ptr(), M, b.data_ptr(), N, out.data_ptr(), N
    d.a_mode, d.b_mode, d.bias, d.aux = 1, 1, 0, 0
    d.splits = int(sys.argv[4][5:] or 1)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    L.check(lib.theia_gemm(C.byref(d), s))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    L.check(lib.theia_gemm(C.byref(d), s))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"M={M} N={N} K={K} epi={epi}: {ms:.4f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")

# ---- wgrad shape (both operands MN-major, fp32 atomics, split-K): python tools/bench_gemm.py Nout Kin Mtok wgrad<splits> [iters] [dbg7] [dbg8]
