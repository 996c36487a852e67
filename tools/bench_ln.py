"""LayerNorm fwd/bwd micro-benchmark: python tools/bench_ln.py M D"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theia_b200 import _lib as L
M, D = int(sys.argv[1]), int(sys.argv[2])
lib = L.lib()
dev = "cuda"
x = torch.randn(M, D, device=dev).to(torch.bfloat16); dy = torch.randn_like(x); dadd = torch.randn_like(x)
y = torch.empty_like(x); dx = torch.empty_like(x)
g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev); dxs = torch.zeros(D, device=dev)
s = torch.cuda.current_stream().cuda_stream
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
f = lambda: L.check(lib.theia_layernorm_fwd(x.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, D, 1e-12, s))
bw = lambda: L.check(lib.theia_layernorm_bwd(dy.data_ptr(), x.data_ptr(), g.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dadd.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), dxs.data_ptr(), M, D, s))
tf, tb = timeit(f), timeit(bw)
print(f"LN fwd {tf*1e3:.1f} us  {M*D*4/tf/1e6:.0f} GB/s   LN bwd {tb*1e3:.1f} us  {M*D*8/tb/1e6:.0f} GB/s")

# ---- LayerNorm[C,H,W] of the heads (stats from the conv epilogue): python tools/bench_ln.py M D ln3d
if len(sys.argv) > 3 and sys.argv[3] == "ln3d":
    B, C, HW = 256, 768, 256
    n = HW * C
    x3 = torch.randn(B, n, device=dev).to(torch.bfloat16); dy3 = torch.randn_like(x3); y3 = torch.empty_like(x3); dx3 = torch.empty_like(x3)
    stats = torch.stack([x3.float().sum(1), (x3.float() ** 2).sum(1)], 1).contiguous()
    g3 = torch.ones(n, device=dev); b3 = torch.zeros(n, device=dev); red = torch.zeros(B, 2, device=dev)
    dg3 = torch.zeros(n, device=dev); db3 = torch.zeros(n, device=dev)
    fa = lambda: L.check(lib.theia_ln3d_apply(x3.data_ptr(), stats.data_ptr(), g3.data_ptr(), b3.data_ptr(), y3.data_ptr(), B, n, 1e-5, C, 0, 0, s))
    fb = lambda: L.check(lib.theia_ln3d_bwd(dy3.data_ptr(), x3.data_ptr(), stats.data_ptr(), g3.data_ptr(), red.data_ptr(), dx3.data_ptr(), dg3.data_ptr(), db3.data_ptr(), B, n, 1e-5, 1, C, 0, 0, s))
    ta, tb3 = timeit(fa), timeit(fb)
    print(f"ln3d apply {ta*1e3:.1f} us ({B*n*4/ta/1e6:.0f} GB/s)   ln3d bwd (reduce+apply) {tb3*1e3:.1f} us ({B*n*10/tb3/1e6:.0f} GB/s)")
