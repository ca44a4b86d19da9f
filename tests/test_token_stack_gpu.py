"""GPU: tensor-core token stacks (csrc/token_stack.cu) vs the fp32 oracle.  bf16 operands -> north_star's 1e-2 relative
bar, taken as the RMS relative error ||a-b||/||b|| after FOUR stacked transformer layers (individual entries pass through
zero, so an entry-wise relative bound is meaningless); the worst entry is additionally held to 3e-2 of max|ref|."""
import numpy as np
import pytest
import torch

from keypointfusion_b200.utils import synth
from oracle import kpf_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_err(a, b):
    """relative error of a tensor: ||a-b|| / ||b|| (RMS); the worst single entry is bounded separately."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    worst = float((a - b).abs().max() / b.abs().max())
    assert worst < 3e-2, worst
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("which,D", [("init_TR", 128), ("final_TR", 131)])
@pytest.mark.parametrize("B", [1, 6, 7, 64])
def test_token_encoder(path_params, which, D, B):
    from keypointfusion_b200 import ops
    prefix = f"block1.{which}."
    wmat, wvec, D_, L, F = ops.pack_token_encoder(path_params, prefix, 21)
    assert (D_, L, F) == (D, 4, 16)
    rs = np.random.RandomState(B + D)
    x = torch.from_numpy(rs.standard_normal((B, 21, D)).astype(np.float32))
    if D == 131:
        x[:, :, :3] *= 0.3
    tok, pred = ops.token_encoder(x.to(DEV), wmat.to(DEV), wvec.to(DEV), L, F)
    rtok, rpred = O.kp_interaction_tr(path_params, prefix, x)
    assert rel_err(tok, rtok) < 1e-2, rel_err(tok, rtok)
    assert rel_err(pred, rpred) < 1e-2, rel_err(pred, rpred)


@pytest.mark.parametrize("B", [2, 13])
def test_token_cross(golden, golden_meta, B):
    from keypointfusion_b200 import ops
    sd = synth.fill_state_dict({k: torch.zeros(s) for k, s in golden_meta["updatedDecoder_keys"].items()}, golden_meta["seed"])
    wmat, wvec, F = ops.pack_token_cross(sd, "decoder.3.", 21)
    if B == 2:
        a, k = torch.from_numpy(golden["a13_anchor"]), torch.from_numpy(golden["a13_key"])
    else:
        rs = np.random.RandomState(B)
        a = torch.from_numpy(rs.standard_normal((B, 21, 128)).astype(np.float32))
        k = torch.from_numpy(rs.standard_normal((B, 21, 128)).astype(np.float32))
    jc = torch.zeros(B, 21, 131, device=DEV)
    out = ops.token_cross(a.to(DEV), k.to(DEV), wmat.to(DEV), wvec.to(DEV), F, out_jc=jc, out_jc_c0=3)
    ref = O.updated_decoder(sd, "", a, k)
    assert rel_err(out, ref) < 1e-2, rel_err(out, ref)
    assert torch.equal(jc[:, :, 3:], out.permute(0, 2, 1)) and not jc[:, :, :3].any()
    if B == 2:
        assert rel_err(out, torch.from_numpy(golden["a13_out"])) < 1e-2
