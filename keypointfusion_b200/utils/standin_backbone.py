"""Stock-PyTorch stand-in for the reference's ConvNeXt-tiny UNet backbones (convNeXT/resnetUnet.py:60-152, convNeXT/convnext.py).

The backbones are OUT OF SCOPE of this repo (north_star: "the ConvNeXt / ResNet-UNet backbones ... stay as they are in PyTorch") and the
reference package cannot be imported on the GPU box, so the full-model configurations of bench.py (BASELINE.json configs 3 and 5) and
the `KPFusion.forward` tests need *a* backbone with the reference's contract:

    forward(img [B,Cin,S,S]) -> (img_offset [B,5J,S/4,S/4], img_feat [B,128,S/4,S/4])

This module has that contract and the reference's size class: a ConvNeXt-T encoder (depths 3-3-9-3, dims 96-192-384-768, 4x4 stride-4
stem with Cin input channels like convNeXTUnet's replaced stem :105-109) and a three-level bilinear-upsample UNet decoder with one
residual unit per skip / up / fusion position, 128-channel output and the three 1x1 heads (3J | J | J).  Plain torch.nn, random
init: it is a workload stand-in, not a re-implementation of the backbone (no weights can be loaded into it).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _LayerNorm2d(nn.LayerNorm):
    def forward(self, x):   # NCHW
        return F.layer_norm(x.permute(0, 2, 3, 1), self.normalized_shape, self.weight, self.bias, self.eps).permute(0, 3, 1, 2)


class _Block(nn.Module):   # ConvNeXt block: dw 7x7 -> LN -> 1x1 (4x) -> GELU -> 1x1, layer scale, residual
    def __init__(self, dim):
        super().__init__()
        self.dw = nn.Conv2d(dim, dim, 7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.pw1, self.pw2 = nn.Linear(dim, 4 * dim), nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(1e-6 * torch.ones(dim))

    def forward(self, x):
        y = self.dw(x).permute(0, 2, 3, 1)
        y = self.pw2(F.gelu(self.pw1(self.norm(y)))) * self.gamma
        return x + y.permute(0, 3, 1, 2)


class _Residual(nn.Module):   # the UNet's residual unit (model/hourglass.py:87-119 shape class): 1x1 -> 3x3 -> 1x1 bottleneck + skip
    def __init__(self, cin, cout):
        super().__init__()
        mid = cout // 2
        self.body = nn.Sequential(nn.BatchNorm2d(cin), nn.ReLU(inplace=True), nn.Conv2d(cin, mid, 1), nn.BatchNorm2d(mid), nn.ReLU(inplace=True),
                                  nn.Conv2d(mid, mid, 3, padding=1), nn.BatchNorm2d(mid), nn.ReLU(inplace=True), nn.Conv2d(mid, cout, 1))
        self.skip = nn.Identity() if cin == cout else nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        return self.body(x) + self.skip(x)


class StandInBackbone(nn.Module):
    def __init__(self, in_ch=1, joint_num=21, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), deconv_dim=128):
        super().__init__()
        self.down = nn.ModuleList([nn.Sequential(nn.Conv2d(in_ch, dims[0], 4, stride=4), _LayerNorm2d(dims[0], eps=1e-6))])
        for i in range(3):
            self.down.append(nn.Sequential(_LayerNorm2d(dims[i], eps=1e-6), nn.Conv2d(dims[i], dims[i + 1], 2, stride=2)))
        self.stages = nn.ModuleList([nn.Sequential(*[_Block(dims[i]) for _ in range(depths[i])]) for i in range(4)])
        up = lambda c: nn.Sequential(_Residual(c, c), nn.Upsample(scale_factor=2, mode="bilinear"))
        self.skip4, self.up4, self.fuse4 = _Residual(dims[2], dims[2]), up(dims[3]), _Residual(dims[2] + dims[3], dims[2])
        self.skip3, self.up3, self.fuse3 = _Residual(dims[1], dims[1]), up(dims[2]), _Residual(dims[2] + dims[1], dims[1])
        self.skip2, self.up2, self.fuse2 = _Residual(dims[0], dims[0]), up(dims[1]), _Residual(dims[1] + dims[0], deconv_dim)
        self.result_emb = _Residual(deconv_dim, deconv_dim)
        self.finals = nn.ModuleList([nn.Conv2d(deconv_dim, o, 1) for o in (3 * joint_num, joint_num, joint_num)])

    def forward(self, img):
        c = []
        x = img
        for d, s in zip(self.down, self.stages):
            x = s(d(x))
            c.append(x)
        c1, c2, c3, c4 = c
        x = self.fuse4(torch.cat((self.up4(c4), self.skip4(c3)), 1))
        x = self.fuse3(torch.cat((self.up3(x), self.skip3(c2)), 1))
        feat = self.fuse2(torch.cat((self.up2(x), self.skip2(c1)), 1))
        pcl_feature = self.result_emb(feat)
        return torch.cat([f(pcl_feature) for f in self.finals], 1), pcl_feature
