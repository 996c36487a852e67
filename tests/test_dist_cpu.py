"""CPU, world_size 2, gloo: the host-side logic of the N>1 path (no GPU compute):
  * the reference's DDP wrapper accepts theia_b200.RobotVisionFM (flat-buffer parameter views), and its
    construction-time broadcast makes rank 1's flat buffer equal rank 0's;
  * allreduce_flat_grads averages a flat gradient buffer across ranks;
  * shard_batch partitions a global batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from theia_b200 import RobotVisionFM
        from theia_b200.dist import allreduce_flat_grads, shard_batch
        from torch.nn.parallel import DistributedDataParallel as DDP
        torch.manual_seed(100 + rank)  # different init per rank: DDP must broadcast rank 0's
        m = RobotVisionFM(backbone="facebook/deit-tiny-patch16-224",
                          target_feature_sizes={"facebook/dinov2-large": (1024, 16, 16)})
        before = m._flat.clone()
        ddp = DDP(m)  # CPU module + gloo: construction broadcasts parameters from rank 0
        flats = [torch.zeros_like(m._flat) for _ in range(world)]
        dist.all_gather(flats, m._flat)
        assert torch.equal(flats[0], flats[1]), "flat parameter buffers differ after DDP broadcast"
        if rank == 1:
            assert not torch.equal(before, m._flat), "rank 1 kept its own init"
        # parameters are still views of the flat buffer after DDP touched them
        name, shape, off = m._param_meta[7]
        assert m._param_list[7].data_ptr() == m._flat.data_ptr() + 4 * off
        assert len(list(ddp.parameters())) == len(m._param_list)
        g = torch.full((1000,), float(rank + 1))
        allreduce_flat_grads(g)
        assert torch.allclose(g, torch.full((1000,), 1.5))
        assert shard_batch(512, rank, world) == (256 * rank, 256 * (rank + 1))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.timeout(300)
def test_ddp_wrapper_and_flat_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
