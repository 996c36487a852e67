"""Teacher inference throughput: python tools/bench_teacher.py [dinov2|clip|vith] [B] [--hf]
DINOv2-L / CLIP ViT-L/14 architecture, seeded random weights, pixel_values resident on the GPU.
Prints one JSON line; --hf also times the HF model (fp32 and autocast bf16) the reference would run."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests._teacher_util import _build, _randomize
from theia_b200 import _lib as L
from theia_b200 import teachers as T

# one process per GPU under torchrun: independent replicas, each on its own images (no collective on the data path)
RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(LOCAL)
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
kind = sys.argv[1] if len(sys.argv) > 1 else "dinov2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ARCH = {"dinov2": ("dinov2", (1024, 16, 24, 14)), "clip": ("clip", (1024, 16, 24, 14)), "vith": ("vit", (1280, 16, 32, 14))}[kind]
hf = _randomize(_build(*ARCH), seed=1)
teacher = T.TeacherViT.from_hf(hf, device=torch.device("cuda", LOCAL), residual_fp32="--bf16-residual" not in sys.argv)
pv = torch.randn(B, 3, 224, 224, device="cuda")


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


l0 = L.lib().theia_launch_count()
if WORLD > 1:
    dist.barrier()
ms = timeit(lambda: teacher(pv, out_dtype=torch.bfloat16))
if WORLD > 1:  # slowest rank
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
launches = (L.lib().theia_launch_count() - l0) // 7
N, D, Ly = 257, ARCH[1][0], ARCH[1][2]
flops = B * Ly * (2.0 * N * D * D * 12 + 4.0 * N * N * D) + B * 2.0 * N * 592 * D
out = {"teacher": {"dinov2": "dinov2-L/14", "clip": "clip-L/14", "vith": "vit-H/14"}[kind], "n_gpus": WORLD, "batch_per_gpu": B, "ms": ms, "img_per_s": WORLD * B / ms * 1e3,
       "tflops_per_gpu": flops / ms / 1e9, "launches_per_forward": launches, "scaling": "weak (independent replicas)",
       "residual_stream": "fp32" if teacher.residual_fp32 else "bf16"}
if "--hf" in sys.argv:
    hf = hf.to("cuda")
    with torch.no_grad():
        out["hf_fp32_img_per_s"] = B / timeit(lambda: hf(pixel_values=pv), n=2, warm=1) * 1e3
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out["hf_autocast_bf16_img_per_s"] = B / timeit(lambda: hf(pixel_values=pv), n=3, warm=1) * 1e3
if RANK == 0:
    print(json.dumps(out))
if WORLD > 1:
    dist.destroy_process_group()
