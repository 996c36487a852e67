"""Needs >= 2 GPUs (gpurun --gpus 2): the reference's DDP wrapper around the CUDA path."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_two_ranks_match_single_gpu_double_batch():
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(here, "ddp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DDP_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_devices_in_one_process():
    """One process driving two GPUs: kernels launch on the MODEL's device whatever the current one is, and the
    per-device state of the library (dynamic shared-memory limits, resize tables, SM count) is set up on each."""
    from oracle import theia_oracle as O
    from theia_b200 import RobotVisionFM
    cfg = O.make_config("facebook/deit-tiny-patch16-224", "dinov2")
    P = O.init_params(cfg, seed=0)
    images, targets = O.synthetic_batch(cfg, 3, seed=0)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        m = RobotVisionFM(backbone="facebook/deit-tiny-patch16-224", translator="lconv",
                          target_feature_sizes=dict(cfg.teachers), translator_kwargs={"hidden_size_factor": 1.0})
        m.load_state_dict(P)
        m = m.to(dev)
        torch.cuda.set_device(0)  # the current device stays cuda:0 throughout
        pred = m(images.to(dev))  # default do_resize=True: the resize tables too
        loss = m.get_loss(pred, {k: v.to(dev) for k, v in targets.items()})
        (0.9 * loss["cos_loss"] + 0.1 * loss["l1_loss"]).backward()
        torch.cuda.synchronize(dev)
        g = m.get_parameter("backbone.model.encoder.layer.11.output.dense.weight").grad
        outs.append((float(loss["cos_loss"].detach()), g.float().cpu(), {k: v.detach().float().cpu() for k, v in pred.items()}))
    assert abs(outs[0][0] - outs[1][0]) < 1e-5
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    # not bit-identical: fp32 atomics (per-image LayerNorm[C,H,W] sums in the conv epilogues, split-K, bias sums) land in
    # a different order on each run, which moves a few bf16 roundings; a device-state bug would give garbage or an error
    for k in outs[0][2]:
        assert rel(outs[0][2][k], outs[1][2][k]) < 2e-3, (k, rel(outs[0][2][k], outs[1][2][k]))
    assert rel(outs[0][1], outs[1][1]) < 2e-2, rel(outs[0][1], outs[1][1])
