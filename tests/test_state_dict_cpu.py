"""CPU: the drop-in modules expose exactly the reference's state_dict keys/shapes (SURVEY.md 8b contract; the key
lists were dumped from the reference's own modules by tests/golden/make_golden.py)."""
import pytest
import torch

from keypointfusion_b200.model.fusion_layer import ACFusion, FSP, RGBDFusion
from keypointfusion_b200.model.model import Block_KPFusion, KPFusion
from keypointfusion_b200.model.transfusion_head import MultiheadAttention, updatedDecoder
from keypointfusion_b200.utils import synth


def _same(mod, ref):
    sd = {k: list(v.shape) for k, v in mod.state_dict().items()}
    assert sd == ref


def test_block_keys(golden_meta):
    """The key list was dumped from the reference running under transformers 5.x, which no longer saves BertEmbeddings.position_ids;
    the reference's pin (4.25.1, requirements.txt:40) registers it as a persistent buffer, so released checkpoints carry it.  The
    drop-in exposes the 4.25.1 list (superset) and loads either (tests/test_packers_cpu.py)."""
    ref = dict(golden_meta["Block_KPFusion_keys"])
    for tr in ("init_TR", "final_TR"):
        ref[f"{tr}.bert.embeddings.position_ids"] = [1, 512]
    _same(Block_KPFusion(joint_num=21), ref)


def test_decoder_and_fusion_keys(golden_meta):
    _same(updatedDecoder(joint_num=21, num_decoder_layers=4), golden_meta["updatedDecoder_keys"])
    _same(RGBDFusion(64, 64), golden_meta["RGBDFusion_keys"])
    _same(ACFusion(64, 64), golden_meta["ACFusion_keys"])
    _same(FSP(64, 64), golden_meta["FSP_keys"])


def test_kpfusion_loads_reference_style_checkpoint(golden_meta, path_params):
    net = KPFusion(joint_num=21)
    missing, unexpected = net.load_state_dict(path_params, strict=True)
    assert not missing and not unexpected
    # DataParallel-style 'module.' prefix filtered like train.py:102-107
    ck = {"module." + k: v for k, v in path_params.items()}
    own = net.state_dict()
    filt = {k[len("module."):]: v for k, v in ck.items() if k[len("module."):] in own}
    assert len(filt) == len([k for k in own if not k.endswith("position_ids")])   # (position_ids: only 4.25.1-era checkpoints carry it)


def test_init_matches_reference_rules():
    torch.manual_seed(0)
    blk = Block_KPFusion()
    # nn.Linear re-initialised by Block_KPFusion.apply(_init_weights) (model.py:269, :282-283) ...
    assert blk.init_TR.cls_head.weight.std() < 0.002 and blk.crossTR.decoder[0].linear1.weight.std() < 0.002
    # ... but in_proj_weight is a bare Parameter and keeps xavier_uniform (transfusion_head.py:668-672)
    assert blk.crossTR.decoder[0].multihead_attn.in_proj_weight.std() > 0.05
    assert float(blk.weight_dis) == 0.0


def test_mha_has_no_cpu_path():
    """MultiheadAttention.forward runs on the kernels of csrc/attn_general.cu for any (L, S); there is no library / CPU route
    (its numerics are checked on the GPU: tests/test_heads_gpu.py, against the reference's goldens and torch's own
    multi_head_attention_forward)."""
    import pytest
    m = MultiheadAttention(64, 4).eval()
    with pytest.raises(RuntimeError, match="CUDA"), torch.no_grad():
        m(torch.randn(5, 2, 64), torch.randn(9, 2, 64), torch.randn(9, 2, 64))
