"""Row-LayerNorm micro-benchmark (HBM-bound): python tools/bench_ln.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theia_b200 import _lib as L
lib = L.lib()
s = torch.cuda.current_stream().cuda_stream
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for M, D in ((50432, 192), (50432, 384), (50432, 768), (32896, 1024)):
    x = torch.randn(M, D, device="cuda").to(torch.bfloat16); dy = torch.randn_like(x); da = torch.randn_like(x)
    y = torch.empty_like(x); dx = torch.empty_like(x)
    g = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
    mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
    dg = torch.zeros(D, device="cuda"); db = torch.zeros(D, device="cuda"); ds = torch.zeros(D, device="cuda")
    def fwd():
        flush.zero_()
        L.check(lib.theia_layernorm_fwd(x.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, D, 1e-6, s))
    def bwd():
        flush.zero_()
        L.check(lib.theia_layernorm_bwd(dy.data_ptr(), x.data_ptr(), g.data_ptr(), mean.data_ptr(), rstd.data_ptr(), da.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), ds.data_ptr(), M, D, s))
    t0 = timeit(lambda: flush.zero_())
    tf, tb = timeit(fwd) - t0, timeit(bwd) - t0
    by = M * D * 2
    print(f"M={M} D={D}: fwd {tf*1e3:.1f} us ({2*by/tf/1e9:.2f} TB/s)   bwd {tb*1e3:.1f} us ({4*by/tb/1e9:.2f} TB/s)")
# LayerNorm[C,16,16] of the heads (NHWC): apply + two-pass backward
for B, C in ((256, 768), (256, 192)):
    n = 256 * C
    x = torch.relu(torch.randn(B, n, device="cuda")).to(torch.bfloat16); dy = torch.randn(B, n, device="cuda").to(torch.bfloat16)
    y = torch.empty_like(x); dx = torch.empty_like(x)
    g = torch.ones(n, device="cuda"); b = torch.zeros(n, device="cuda")
    stats = torch.stack([x.float().sum(1), (x.float() ** 2).sum(1)], 1).contiguous()
    red = torch.zeros(B, 2, device="cuda"); dg = torch.zeros(n, device="cuda"); db = torch.zeros(n, device="cuda")
    def app():
        flush.zero_()
        L.check(lib.theia_ln3d_apply(x.data_ptr(), stats.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(), B, n, 1e-5, C, 0, 0, s))
    def bwd3():
        flush.zero_()
        L.check(lib.theia_ln3d_bwd(dy.data_ptr(), x.data_ptr(), stats.data_ptr(), g.data_ptr(), red.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), B, n, 1e-5, 1, C, 0, 0, s))
    t0 = timeit(lambda: flush.zero_())
    ta, tb = timeit(app) - t0, timeit(bwd3) - t0
    by = B * n * 2
    print(f"ln3d B={B} C={C}: apply {ta*1e3:.1f} us ({2*by/ta/1e9:.2f} TB/s)   bwd (reduce + apply) {tb*1e3:.1f} us ({5*by/tb/1e9:.2f} TB/s)")
