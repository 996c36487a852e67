"""Per-parameter gradient error of the CUDA path vs the oracle (GPU fp32 eager).  Diagnostic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import theia_oracle as O
from theia_b200 import RobotVisionFM

backbone = sys.argv[1] if len(sys.argv) > 1 else "facebook/deit-tiny-patch16-224"
tset = sys.argv[2] if len(sys.argv) > 2 else "dinov2"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = O.make_config(backbone, tset)
P = O.init_params(cfg, seed=0)
m = RobotVisionFM(backbone=backbone, target_feature_sizes=dict(cfg.teachers)).cuda()
m.load_state_dict(P)
Pd = {k: v.cuda() for k, v in P.items()}
images, targets = O.synthetic_batch(cfg, B, seed=0, device="cuda")
pred_o, lo, go = O.distill_step(Pd, images, targets, cfg, do_resize=False)
pred = m(images, do_resize=False)
losses = m.get_loss(pred, targets)
(0.9 * losses["cos_loss"] + 0.1 * losses["l1_loss"]).backward()
g = {k: p.grad for k, p in m.named_parameters()}
gmax = max(v.norm().item() for v in go.values())
rows = []
for k, v in go.items():
    e = (g[k].double() - v.double()).norm().item()
    cosang = torch.nn.functional.cosine_similarity(g[k].flatten().double(), v.flatten().double(), dim=0).item()
    rows.append((k, e / (v.double().norm().item() + 1e-30), v.norm().item() / gmax, g[k].norm().item() / (v.norm().item() + 1e-30), cosang))
print(f"{'param':90s} relerr   |ref|/max  |got|/|ref|  cos")
for r in rows:
    print(f"{r[0]:90s} {r[1]:.4f}  {r[2]:.2e}  {r[3]:.4f}  {r[4]:.5f}")
