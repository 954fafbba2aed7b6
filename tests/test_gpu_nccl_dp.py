"""Multi-GPU data parallelism over NCCL: needs >= 2 visible GPUs (run with `gpurun --gpus 2`), skipped otherwise.

  * the overlapped, bucketed all-reduce of the gradient arena equals the SUM of the ranks' local gradients (fp32, <= 1e-6 rel);
  * N ranks x B images == 1 rank x N*B images for a BatchNorm-free model (ViT): mean-loss gradient = average of the shard
    gradients, so parameters after a step agree;
  * replicas stay bit-identical; the single-graph step (NCCL inside the CUDA graph) equals the eager step.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _vit(seed=0):
    from deeplearning_b200.classification.vision_transformer.vit_model import VisionTransformer

    torch.manual_seed(seed)
    return VisionTransformer(img_size=224, patch_size=16, embed_dim=768, depth=2, num_heads=12, num_classes=16)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from deeplearning_b200.engine.trainer import TrainStep

        B = 8
        g = torch.Generator().manual_seed(11)
        x_all = torch.randn(world * B, 3, 224, 224, generator=g)
        y_all = torch.randint(0, 16, (world * B,), generator=g)
        x, y = x_all[rank * B:(rank + 1) * B].to(dev), y_all[rank * B:(rank + 1) * B].to(dev)
        res = {}
        # ---- (1) local gradients, no communication (world_size=1 arena on this rank)
        m = _vit().to(dev).train()
        tr = TrainStep(m, lr=0.0, momentum=0.0, weight_decay=0.0, world_size=1, broadcast=False)
        tr.step_eager(x, y)
        local = tr.arena.flat_g.clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        expect_sum = torch.stack(gathered).double().sum(0)
        # ---- (2) overlapped bucketed all-reduce (small buckets: many collectives in flight during the backward pass)
        m2 = _vit().to(dev).train()
        tr2 = TrainStep(m2, lr=0.05, momentum=0.9, weight_decay=5e-5, bucket_mb=4.0)
        before = tr2.arena.flat_p.clone()
        tr2.step_eager(x, y)
        got = tr2.arena.flat_g.double()
        res["allreduce_rel"] = float((got - expect_sum).norm() / expect_sum.norm())
        res["buckets"] = tr2.arena.buckets_launched
        res["p_after_eager"] = tr2.arena.flat_p.clone().cpu()
        # ---- (3) one trailing all-reduce gives the same reduced gradient
        m3 = _vit().to(dev).train()
        tr3 = TrainStep(m3, lr=0.05, momentum=0.9, weight_decay=5e-5, overlap=False)
        tr3.step_eager(x, y)
        res["overlap_vs_single_rel"] = float((tr3.arena.flat_g.double() - got).norm() / got.norm())
        # ---- (4) captured step (NCCL nodes inside the graph) == eager step
        m4 = _vit().to(dev).train()
        tr4 = TrainStep(m4, lr=0.05, momentum=0.9, weight_decay=5e-5, bucket_mb=4.0)
        tr4.capture(x, y)
        assert torch.equal(tr4.arena.flat_p, before)          # capture() is side-effect free
        tr4.step(x, y)
        torch.cuda.synchronize()
        res["graph_vs_eager_rel"] = float((tr4.arena.flat_p.double().cpu() - res["p_after_eager"].double()).norm()
                                          / (res["p_after_eager"].double() - before.double().cpu()).norm())
        res["update_norm"] = float((res["p_after_eager"].double() - before.double().cpu()).norm())
        out[rank] = res
        # a CUDA graph that contains NCCL kernels must be gone before the communicator is destroyed
        del tr4, tr3, tr2, tr
        import gc

        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_step_equals_single_process_double_batch():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    print({k: v for k, v in r0.items() if not torch.is_tensor(v)})
    assert r0["allreduce_rel"] <= 1e-6 and r1["allreduce_rel"] <= 1e-6, (r0["allreduce_rel"], r1["allreduce_rel"])
    assert r0["buckets"] >= 3                                   # really bucketed
    assert r0["overlap_vs_single_rel"] <= 1e-6
    assert torch.equal(r0["p_after_eager"], r1["p_after_eager"])   # replicas stay bit-identical
    assert r0["graph_vs_eager_rel"] <= 1e-3, r0["graph_vs_eager_rel"]
    # single process, double batch (BN-free model): parameters after the same step
    from deeplearning_b200.engine.trainer import TrainStep

    B = 8
    g = torch.Generator().manual_seed(11)
    x_all = torch.randn(world * B, 3, 224, 224, generator=g).cuda()
    y_all = torch.randint(0, 16, (world * B,), generator=g).cuda()
    m = _vit().cuda().train()
    tr = TrainStep(m, lr=0.05, momentum=0.9, weight_decay=5e-5, world_size=1, broadcast=False)
    before = tr.arena.flat_p.clone().cpu()
    tr.step_eager(x_all, y_all)
    single = tr.arena.flat_p.cpu()
    upd_s, upd_d = (single - before).double(), (r0["p_after_eager"] - before).double()
    rel = float((upd_s - upd_d).norm() / upd_s.norm())
    print(f"2 ranks x {B} vs 1 rank x {2 * B}: relative difference of the parameter update {rel:.3g}")
    assert rel <= 2e-3, rel
