"""Repeat the long-sequence attention forward (head dim 64 and 80) on several geometries: every run must match
torch fp32 and be bit-identical to the first one (race / ordering check).  python tools/stress_attn_long.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theia_b200 import _lib as L
lib = L.lib(); s = torch.cuda.current_stream().cuda_stream
bad = 0; runs = 0
for hd, fn in ((64, lib.theia_attention_tc_fwd), (80, lib.theia_attention_fwd_hd80)):
    for (B, H, N) in ((128, 16, 257), (37, 16, 257), (9, 16, 258), (50, 12, 272), (64, 16, 259), (21, 7, 230)):
        D = H * hd
        g = torch.Generator(device="cuda").manual_seed(B * 1000 + N + hd)
        qkv = (1.5 * torch.randn(B * N, 3 * D, device="cuda", generator=g)).to(torch.bfloat16)
        x = qkv.float().view(B, N, 3, H, hd)
        q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
        ref = (torch.softmax((q @ k.transpose(2, 3)) * hd ** -0.5, -1) @ v).transpose(1, 2).reshape(B * N, D)
        first = None
        for it in range(12):
            out = torch.full((B * N, D), float("nan"), dtype=torch.bfloat16, device="cuda")
            L.check(fn(qkv.data_ptr(), out.data_ptr(), 0, B, N, H, s))
            torch.cuda.synchronize()
            runs += 1
            e = ((out.float() - ref).norm() / ref.norm()).item()
            if first is None: first = out.clone()
            same = torch.equal(out, first)
            if not (e < 8e-3) or not same:
                bad += 1; print("MISMATCH", hd, B, H, N, it, e, same)
print("stress runs", runs, "bad", bad)
