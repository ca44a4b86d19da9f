"""GPU (>= 2 devices; skipped on a single-GPU box): the fused exchange step.  Each rank's final token-stack kernel stores its joints
straight into every rank's gathered tensor (peer stores into symmetric memory + arrival counter, runtime.PeerExchange); the result must
be bit-identical to an ncclAllGather of the ranks' own joints, step after step (double buffering, step counter), eager and from a
captured CUDA graph."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from keypointfusion_b200 import ops
    from keypointfusion_b200.dataloader.loader import loader
    from keypointfusion_b200.model.model import KPFusion
    from keypointfusion_b200.runtime import GraphedFusionPath, PeerExchange
    from keypointfusion_b200.utils import synth
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        B, J = 3, 21
        net = KPFusion(joint_num=J)
        synth.fill_state_dict(net, 0)
        net = net.to(dev).eval()
        px = PeerExchange(B, J, dev)
        L = loader(img_size=128)
        ok = True
        for step in range(3):        # eager steps
            inp = synth.make_inputs(B, 128, J, 128, seed=100 * rank + step, bf16_round=True)
            c = {k: v.to(dev) for k, v in inp.items()}
            with torch.no_grad():
                pcl, _ = ops.getpcl(c["img"], c["center"], c["cube"], c["M"], c["cam"], seed=1)
                px.begin_step()
                res, _, _ = net.forward_path(c["img_offset"].bfloat16(), c["img_feat"].bfloat16(), None, c["img_feat_rgb"].bfloat16(), c["img"], pcl, L,
                                             c["center"], c["M"], c["cube"], c["cam"], 0.8, exchange=px)
                px.flush()
            ref = torch.empty(world * B, J, 3, device=dev)
            dist.all_gather_into_tensor(ref, res[-1].contiguous())
            torch.cuda.synchronize()
            ok = ok and torch.equal(px.gathered(), ref)
        # the same inside a captured graph, replayed with fresh inputs
        inp = synth.make_inputs(B, 128, J, 128, seed=100 * rank + 50, bf16_round=True)
        ex = {k: inp[k].to(dev) for k in GraphedFusionPath.KEYS}
        for k in ("img_feat", "img_feat_rgb", "img_offset"):
            ex[k] = ex[k].bfloat16()
        gp = GraphedFusionPath(net, L, ex, exchange=px)
        for step in range(3):
            inp = synth.make_inputs(B, 128, J, 128, seed=100 * rank + 60 + step, bf16_round=True)
            d = {k: inp[k].to(dev) for k in GraphedFusionPath.KEYS}
            for k in ("img_feat", "img_feat_rgb", "img_offset"):
                d[k] = d[k].bfloat16()
            joints = gp(d)["joints"]          # the replay completes the PREVIOUS step's gather at its start and leaves this one in flight
            ref = torch.empty(world * B, J, 3, device=dev)
            dist.all_gather_into_tensor(ref, joints.contiguous())
            if step == 1:
                px.flush()                    # a consumer that wants the current step's gather completes it explicitly
            torch.cuda.synchronize()
            if step == 1:
                ok = ok and torch.equal(px.gathered(), ref)
        px.flush()
        torch.cuda.synchronize()
        ok = ok and torch.equal(px.gathered(), ref)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs with peer access")
def test_fused_exchange_matches_nccl_all_gather():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert res == [(0, True), (1, True)], res
