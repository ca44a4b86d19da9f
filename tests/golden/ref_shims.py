"""Import shims that let the UNMODIFIED reference (/root/reference) run on CPU in the
build container.  Test infrastructure only: used by make_golden.py (and nothing that ships).

Every shim is glue (missing third-party modules, removed transformers helpers, hard-coded
.cuda()) except `pointnet2_ops.QueryAndGroup`, which is an arithmetic restatement of
pointnet2_ops 3.0.0 (requirements.txt:15; not vendored, CUDA-only) -- parity for DESA is
therefore "unpinned" (SURVEY.md 8c).
"""
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("KPF_REFERENCE", "/root/reference")


class _Permissive(types.ModuleType):
    """Stub module: any missing attribute resolves to a do-nothing placeholder (import-time glue only)."""

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return lambda *a, **k: None


def _stub(name, **attrs):
    m = _Permissive(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class QueryAndGroup(nn.Module):
    """pointnet2_ops.pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=True).

    ball_query: for each centre, the first `nsample` point indices (ascending) with d^2 < r^2;
    remaining slots are filled with the first hit; idx is zero-initialised.
    group: cat([xyz[idx] - centre, feat[idx]], dim=1) -> [B, 3+C, J, nsample].
    """

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        B, N, _ = xyz.shape
        J = new_xyz.shape[1]
        d2 = ((new_xyz.unsqueeze(2) - xyz.unsqueeze(1)) ** 2).sum(-1)  # B J N
        inside = d2 < self.radius ** 2
        ar = torch.arange(N, device=xyz.device).view(1, 1, N).expand(B, J, N)
        key = torch.where(inside, ar, torch.full_like(ar, N))
        order = key.sort(dim=-1)[0][:, :, :self.nsample]  # first nsample hits ascending, N = miss
        cnt = inside.sum(-1, keepdim=True)
        first = order[:, :, :1]
        first = torch.where(cnt > 0, first, torch.zeros_like(first))
        slot = torch.arange(self.nsample, device=xyz.device).view(1, 1, -1)
        idx = torch.where(slot < cnt, order, first.expand_as(order))
        gi = idx.reshape(B, 1, J * self.nsample)
        grouped_xyz = torch.gather(xyz.transpose(1, 2), 2, gi.expand(B, 3, -1)).view(B, 3, J, self.nsample)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        C = features.shape[1]
        grouped_feat = torch.gather(features, 2, gi.expand(B, C, -1)).view(B, C, J, self.nsample)
        return torch.cat([grouped_xyz, grouped_feat], dim=1)


_installed = False


def install():
    global _installed
    if _installed:
        return
    _installed = True
    # -- removed transformers helper (transfusion_head.py:13)
    import transformers.pytorch_utils as tpu
    if not hasattr(tpu, "torch_int_div"):
        tpu.torch_int_div = lambda a, b: torch.div(a, b, rounding_mode="floor")
    # -- init_weights() on transformers 5.x (model.py:43,:114)
    from transformers.models.bert.modeling_bert import BertPreTrainedModel

    def _init_weights_compat(self):
        if getattr(self, "_kpf_in_init", False):
            return
        self._kpf_in_init = True
        try:
            self.post_init()
        finally:
            self._kpf_in_init = False
    BertPreTrainedModel.init_weights = _init_weights_compat
    # -- missing third-party modules
    _stub("pointnet2_ops", pointnet2_utils=types.SimpleNamespace(QueryAndGroup=QueryAndGroup))
    _stub("pointnet2_ops.pointnet2_utils", QueryAndGroup=QueryAndGroup)
    timm = _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", trunc_normal_=nn.init.trunc_normal_, DropPath=lambda *a, **k: nn.Identity())
    _stub("timm.models.registry", register_model=lambda f: f)
    for name in ["pycocotools", "pycocotools.coco", "matplotlib", "matplotlib.pyplot", "trimesh", "pytorch3d",
                 "pytorch3d.transforms", "tensorboardX", "chumpy", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "sklearn.decomposition", "dataloader.webuser", "dataloader.webuser.smpl_handpca_wrapper_HAND_only"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name)
    sys.modules["pycocotools.coco"].__dict__.setdefault("COCO", object)
    mpl = sys.modules["matplotlib"]
    if isinstance(mpl, _Permissive):
        mpl.__path__ = []            # make the stub a package so `import matplotlib.colors` resolves to the stubs below
        mpl.cm = _stub("matplotlib.cm")
        mpl.colors = _stub("matplotlib.colors")
    sys.modules["tensorboardX"].__dict__.setdefault("SummaryWriter", object)
    sys.modules["mpl_toolkits.mplot3d"].__dict__.setdefault("Axes3D", object)
    # -- hard-coded .cuda() on a CPU box
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    if REF not in sys.path:
        sys.path.insert(0, REF)
    os.chdir(REF)  # BertConfig.from_pretrained("./config/") is cwd-relative (model.py:222)


def ref_loader(img_size=128):
    """Geometry helper object the reference passes into forward() as `loader`."""
    install()
    from dataloader.loader import loader
    obj = loader.__new__(loader)
    try:
        loader.__init__(obj, '', 'test', img_size, 'joint_mean', 'x')
    except Exception:
        pass
    obj.img_size = img_size
    obj.flip = 1
    return obj
