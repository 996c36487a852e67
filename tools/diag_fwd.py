"""Stage-by-stage forward comparison: CUDA path vs the oracle (fp32 and bf16-storage emulation)."""
import ctypes as C, dataclasses, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle import theia_oracle as O
from theia_b200 import RobotVisionFM, _lib as L

backbone = sys.argv[1] if len(sys.argv) > 1 else "facebook/deit-tiny-patch16-224"
tset = sys.argv[2] if len(sys.argv) > 2 else "dinov2"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = O.make_config(backbone, tset)
P = O.init_params(cfg, seed=0)
m = RobotVisionFM(backbone=backbone, target_feature_sizes=dict(cfg.teachers)).cuda()
m.load_state_dict(P)
Pd = {k: v.cuda() for k, v in P.items()}
images, targets = O.synthetic_batch(cfg, B, seed=0, device="cuda")
with torch.no_grad():
    pred = m(images, do_resize=False)
    t32, te = {}, {}
    p32 = O.forward(Pd, images, cfg, taps=t32, do_resize=False)
    pe = O.forward(Pd, images, dataclasses.replace(cfg, emulate_bf16=True), taps=te, do_resize=False)

def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()

def fetch(name, i, shape):
    ptr, n, f32 = C.c_void_p(), C.c_longlong(), C.c_int()
    L.check(L.lib().theia_model_debug_ptr(m._handle, name.encode(), i, C.byref(ptr), C.byref(n), C.byref(f32)), name)
    assert n.value == torch.Size(shape).numel(), (name, n.value, shape)
    buf = (C.c_uint16 * n.value).from_address(0)  # placeholder, not used
    t = torch.empty(shape, dtype=torch.bfloat16, device="cuda")
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12") if False else None
    # device-to-device copy through torch: wrap raw pointer via from_blob-like trick using cuda IPC-free path
    src = torch.cuda.ByteStorage  # noqa
    return _from_ptr(ptr.value, shape)

def _from_ptr(p, shape):
    n = torch.Size(shape).numel()
    # __cuda_array_interface__ wrapper
    class W:
        pass
    w = W()
    w.__cuda_array_interface__ = {"shape": (n,), "typestr": "<u2", "data": (p, False), "version": 2}
    raw = torch.as_tensor(w, device="cuda")
    return raw.view(torch.bfloat16).view(shape).clone()

D = cfg.hidden
print(f"{'stage':16s} vs_fp32    vs_bf16emu   (fp32 vs emu)")
rows = [("x", 0)]
for l in (0, 1, 5, 11):
    rows += [("ln1", l), ("qkv", l), ("attn", l), ("xmid", l), ("ln2", l), ("a", l), ("x", l + 1)]
for name, l in rows:
    width = {"qkv": 3 * D, "h": 4 * D, "a": 4 * D}.get(name, D)
    mine = fetch(name, l, (B, 197, width))
    print(f"{name+'['+str(l)+']':16s} {rel(mine, t32[(name, l)]):.5f}    {rel(mine, te[(name, l)]):.5f}     {rel(te[(name,l)], t32[(name,l)]):.5f}")
mine = fetch("tokens", 0, (B, 197, D))
print(f"{'tokens':16s} {rel(mine, t32[('tokens', 0)]):.5f}    {rel(mine, te[('tokens', 0)]):.5f}     {rel(te[('tokens',0)], t32[('tokens',0)]):.5f}")
for i, t in enumerate(cfg.teachers):
    for name in ("padout", "hln0", "c1", "hln1", "c2", "hln2"):
        mine = fetch(name, i, (B, 16, 16, D))
        print(f"{name+'['+t[:12]+']':28s} {rel(mine, t32[(name, t)]):.5f}    {rel(mine, te[(name, t)]):.5f}     {rel(te[(name,t)], t32[(name,t)]):.5f}")
    print(f"{'pred['+t[:12]+']':28s} {rel(pred[t], p32[t]):.5f}    {rel(pred[t], pe[t]):.5f}     {rel(pe[t], p32[t]):.5f}")
