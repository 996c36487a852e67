"""debug: per-row error pattern of the attention backward"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from theia_b200 import _lib as L
lib = L.lib()
Bn, H = int(sys.argv[1]), int(sys.argv[2])
N, D = 197, H * 64
g = torch.Generator().manual_seed(1)
qkv = (1.5 * torch.randn(Bn * N, 3 * D, generator=g)).to(torch.bfloat16).cuda()
do = torch.randn(Bn * N, D, generator=g).to(torch.bfloat16).cuda()
out = torch.empty(Bn * N, D, dtype=torch.bfloat16, device="cuda")
lse = torch.empty(Bn, H, N, device="cuda")
s = torch.cuda.current_stream().cuda_stream
L.check(lib.theia_attention_tc_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), Bn, N, H, s))
x = qkv.float().view(Bn, N, 3, H, 64).requires_grad_(True)
q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
sc = (q @ k.transpose(2, 3)) * 0.125
ref = (torch.softmax(sc, -1) @ v).transpose(1, 2).reshape(Bn * N, D)
ref.backward(do.float())
gr = x.grad.view(Bn, N, 3, H, 64)
nfail = 0
for rep in range(int(sys.argv[3]) if len(sys.argv) > 3 else 20):
    dqkv = torch.full_like(qkv, float("nan"))
    L.check(lib.theia_attention_tc_bwd(qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), Bn, N, H, s))
    torch.cuda.synchronize()
    got = dqkv.float().view(Bn, N, 3, H, 64)
    for i, nm in enumerate("qkv"):
        tot = ((got[:, :, i] - gr[:, :, i]).norm() / gr[:, :, i].norm()).item()
        if not (tot < 0.015):
            nfail += 1
            for b in range(Bn):
                for h in range(H):
                    e = (got[b, :, i, h] - gr[b, :, i, h]).norm(dim=-1) / (gr[b, :, i, h].norm(dim=-1) + 1e-9)
                    bad = (~(e < 0.05)).nonzero().flatten().tolist()
                    if bad:
                        print("rep", rep, nm, "b", b, "h", h, "total %.4f" % tot, "bad rows:", bad[:10], "...", bad[-10:], len(bad))
print("failures:", nfail)
