"""Launched by torchrun (2+ GPUs): RobotVisionFM under the reference's DDP wrapper (train_rvfm.py:258).
Checks: gradients of an N-rank run with per-rank batch B equal those of a single-GPU run on the
concatenated N*B batch; parameters stay bit-identical across ranks after optimizer steps."""
import os
import sys

import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import theia_oracle as O  # noqa: E402
from theia_b200 import RobotVisionFM  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    backbone, tset, B = "facebook/deit-tiny-patch16-224", "cdiv", 4
    cfg = O.make_config(backbone, tset)
    P = O.init_params(cfg, seed=0)

    def make():
        m = RobotVisionFM(backbone=backbone, target_feature_sizes=dict(cfg.teachers), max_batch=B * world)
        m.load_state_dict(P)
        return m.to(dev)

    shards = [O.synthetic_batch(cfg, B, seed=10 + r, device=dev) for r in range(world)]
    images, targets = shards[rank]
    m = make()
    ddp = DDP(m, device_ids=[local], find_unused_parameters=False)
    pred = ddp(images, do_resize=False)
    losses = m.get_loss(pred, targets)
    (0.9 * losses["cos_loss"] + 0.1 * losses["l1_loss"]).backward()
    g_ddp = torch.cat([p.grad.flatten() for p in m.parameters()])
    # single-process reference on the concatenated batch (every rank computes it: cheap at this size)
    ref = make()
    big_i = torch.cat([s[0] for s in shards])
    big_t = {t: torch.cat([s[1][t] for s in shards]) for t in cfg.teachers}
    pr = ref(big_i, do_resize=False)
    lr = ref.get_loss(pr, big_t)
    (0.9 * lr["cos_loss"] + 0.1 * lr["l1_loss"]).backward()
    g_ref = torch.cat([p.grad.flatten() for p in ref.parameters()])
    rel = ((g_ddp - g_ref).norm() / g_ref.norm()).item()
    assert rel < 2e-2, f"rank {rank}: DDP grads differ from the {world}x batch run: rel {rel}"
    # the module's own single flat all-reduce (no DDP wrapper) gives the same gradients
    m2 = make()
    m2.sync_gradients(True)
    p2 = m2(images, do_resize=False)
    l2 = m2.get_loss(p2, targets)
    (0.9 * l2["cos_loss"] + 0.1 * l2["l1_loss"]).backward()
    g_flat = torch.cat([p.grad.flatten() for p in m2.parameters()])
    rel2 = ((g_flat - g_ref).norm() / g_ref.norm()).item()
    assert rel2 < 2e-2, f"rank {rank}: flat all-reduce grads differ: rel {rel2}"
    # optimizer steps keep ranks bit-identical
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0.01, fused=True)
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        pred = ddp(images, do_resize=False)
        losses = m.get_loss(pred, targets)
        (0.9 * losses["cos_loss"] + 0.1 * losses["l1_loss"]).backward()
        opt.step()
    flats = [torch.empty_like(m._flat) for _ in range(world)]
    dist.all_gather(flats, m._flat)
    for r in range(1, world):
        assert torch.equal(flats[0], flats[r]), f"parameters diverged between rank 0 and {r}"
    # the wrapper-free path: sync_gradients() (ONE all-reduce of the flat gradient buffer inside backward) stepped
    # by the fused FlatAdamW -- parameters AND the bf16 operand copies must stay bit-identical across ranks
    from theia_b200.optim import FlatAdamW
    m3 = make()
    m3.sync_gradients(True)
    opt3 = FlatAdamW(m3, lr=1e-3, weight_decay=0.01)
    for _ in range(3):
        p3 = m3(images, do_resize=False)
        l3 = m3.get_loss(p3, targets)
        opt3.zero_grad(set_to_none=True)
        (0.9 * l3["cos_loss"] + 0.1 * l3["l1_loss"]).backward()
        opt3.step()
    flats = [torch.empty_like(m3._flat) for _ in range(world)]
    dist.all_gather(flats, m3._flat)
    for r in range(1, world):
        assert torch.equal(flats[0], flats[r]), f"FlatAdamW + sync_gradients: parameters diverged (rank 0 vs {r})"
    with torch.no_grad():
        feat = m3.forward_feature(shards[0][0], do_resize=False).contiguous()  # same images on every rank
    feats = [torch.empty_like(feat) for _ in range(world)]
    dist.all_gather(feats, feat)
    for r in range(1, world):
        assert torch.equal(feats[0], feats[r]), f"forward differs between rank 0 and {r} after FlatAdamW steps"
    dist.barrier()
    if rank == 0:
        print(f"DDP_CHECK_OK world={world} grad_rel={rel:.3e} flat_allreduce_rel={rel2:.3e} "
              f"flat_adamw_sync_gradients=bit-identical-across-ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
