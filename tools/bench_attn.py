"""Attention micro-benchmark: python tools/bench_attn.py B H [N]   (N > 208: forward only, the teacher kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theia_b200 import _lib as L
B, H = int(sys.argv[1]), int(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else 197
D = H * 64
lib = L.lib()
qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.empty(B * N, D, dtype=torch.bfloat16, device="cuda")
do = torch.randn(B * N, D, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
lse = torch.empty(B, H, N, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
fl = 4.0 * N * N * 64 * B * H
ms = timeit(lambda: L.check(lib.theia_attention_tc_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, s)))
print(f"theia_attention_tc_fwd N={N}: {ms:.4f} ms  {fl/ms/1e9:.1f} TFLOP/s (algorithmic)")
if N <= 208:
    ms = timeit(lambda: L.check(lib.theia_attention_tc_bwd(qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, N, H, s)))
    print(f"theia_attention_tc_bwd N={N}: {ms:.4f} ms  {2.5*fl/ms/1e9:.1f} TFLOP/s (algorithmic, 2.5x fwd)")
