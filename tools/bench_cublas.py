"""cuBLAS (torch.matmul, bf16) on the step's GEMM shapes, as a library yardstick for tools/bench_gemm.py."""
import sys, torch
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
dev = "cuda"
M = 50432
for (m, n, k, tag) in [(M, 3072, 768, "fc1 fwd"), (M, 768, 3072, "fc2 fwd"), (M, 2304, 768, "qkv fwd"), (M, 768, 768, "proj fwd"),
                       (768, 3072, M, "fc1 wgrad"), (3072, 768, M, "fc2 wgrad"), (M, 3072, 8192, "big K")]:
    a = torch.randn(m, k, device=dev, dtype=torch.bfloat16)
    b = torch.randn(n, k, device=dev, dtype=torch.bfloat16)
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    ms = t(lambda: torch.matmul(a, b.t(), out=out))
    print(f"cuBLAS {tag:10s} M={m} N={n} K={k}: {ms*1e3:.1f} us  {2.0*m*n*k/ms/1e9:.0f} TFLOP/s")
    if tag == "fc1 fwd":
        bias = torch.randn(n, device=dev, dtype=torch.bfloat16)
        ms = t(lambda: torch.nn.functional.gelu(torch.nn.functional.linear(a, b, bias)))
        print(f"torch linear+gelu (2 kernels) {ms*1e3:.1f} us  {2.0*m*n*k/ms/1e9:.0f} TFLOP/s")
