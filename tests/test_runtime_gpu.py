"""GPU: the CUDA-graph runtime and the host-fed pipelined runner reproduce the eager path bit for bit."""
import pytest
import torch

from keypointfusion_b200.utils import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def net(path_params):
    from keypointfusion_b200.model.model import KPFusion
    n = KPFusion(joint_num=21)
    n.load_state_dict(path_params, strict=False)
    return n.to(DEV).eval()


def make(B, seed):
    inp = synth.make_inputs(B, 128, 21, 128, seed=seed)
    for k in ("img_feat", "img_feat_rgb", "img_offset"):
        inp[k] = inp[k].bfloat16()
    inp.pop("img_rgb")
    return inp


def eager(net, d, seed=0):
    from keypointfusion_b200 import ops
    from keypointfusion_b200.dataloader.loader import loader
    with torch.no_grad():
        pcl, _ = ops.getpcl(d["img"], d["center"], d["cube"], d["M"], d["cam"], 1024, seed=seed)
        res, _, _ = net.forward_path(d["img_offset"], d["img_feat"], None, d["img_feat_rgb"], d["img"], pcl, loader(img_size=128),
                                     d["center"], d["M"], d["cube"], d["cam"], 0.8)
    return res[-1]


def test_graph_and_pipelined_runner_match_eager(net):
    from keypointfusion_b200.dataloader.loader import loader
    from keypointfusion_b200.runtime import GraphedFusionPath, PipelinedRunner
    B = 6
    hosts = [make(B, 700 + i) for i in range(3)]
    devs = [{k: v.to(DEV) for k, v in h.items()} for h in hosts]
    ref = [eager(net, d).clone() for d in devs]
    ldr = loader(img_size=128)
    # static-buffer graph: inputs copied in, one replay
    g = GraphedFusionPath(net, ldr, devs[0], sample_num=1024, kernel=0.8, seed=0)
    for d, r in zip(devs, ref):
        assert torch.equal(g(d)["joints"], r)
    # graph bound to the caller's tensors: replay reads whatever they hold now
    gb = GraphedFusionPath(net, ldr, devs[1], sample_num=1024, kernel=0.8, seed=0, bind=True)
    assert torch.equal(gb()["joints"], ref[1])
    for k in devs[1]:
        devs[1][k].copy_(devs[2][k])
    assert torch.equal(gb()["joints"], ref[2])
    # pipelined runner: pinned arenas (one upload per step) and plain pinned tensors (per-tensor uploads)
    runner = PipelinedRunner(net, ldr, devs[0], sample_num=1024, kernel=0.8, seed=0)
    arenas = []
    for h in hosts:
        d = runner.new_host_inputs()
        for k, v in h.items():
            d[k].copy_(v)
        arenas.append(d)
    plain = [{k: v.pin_memory() for k, v in h.items()} for h in hosts]
    order = [0, 1, 2, 2, 0, 1]
    for src in (arenas, plain):
        slots = []
        outs = []
        for i in order:
            slots.append(runner.submit(src[i]))
            if len(slots) > 1:
                outs.append(runner.fetch(slots.pop(0)).clone())
        outs.append(runner.fetch(slots.pop(0)).clone())
        for i, o in zip(order, outs):
            assert torch.equal(o, ref[i].cpu()), i
