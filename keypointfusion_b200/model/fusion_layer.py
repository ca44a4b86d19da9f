"""Drop-in for the reference's model/fusion_layer.py: same classes, constructor arguments, state_dict keys and return
structure; forward runs on the B200 kernels (K7).  Inference only (no autograd through the kernels)."""
import torch
import torch.nn as nn

from .. import custom_ops  # noqa: F401  (registers torch.ops.kpf.*)
from .. import ops


def _inference_only(module, *tensors):
    """These modules sit inside TRAINED backbones in the reference (model/resnet.py:438-498); the kernels have no backward, so
    running them under autograd would silently drop the gradients of the gates and of everything upstream."""
    if module.training or (torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors)):
        raise RuntimeError(f"{type(module).__name__}: inference only (no autograd through the B200 kernels); call .eval() and run under "
                           "torch.no_grad()")


class FilterLayer(nn.Module):
    # model/fusion_layer.py:6-22
    def __init__(self, in_planes, out_planes, reduction=16):
        super(FilterLayer, self).__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(in_planes, out_planes // reduction), nn.ReLU(inplace=True),
                                nn.Linear(out_planes // reduction, out_planes), nn.Sigmoid())
        self.out_planes = out_planes

    def forward(self, x):
        _inference_only(self, x)
        b, c = x.shape[:2]
        y = ops.channel_mean(x)
        y = torch.sigmoid(torch.addmm(self.fc[2].bias, torch.relu(torch.addmm(self.fc[0].bias, y, self.fc[0].weight.t())),
                                      self.fc[2].weight.t()))
        return y.view(b, self.out_planes, 1, 1)


class FSP(nn.Module):
    # model/fusion_layer.py:28-37
    def __init__(self, in_planes, out_planes, reduction=16):
        super(FSP, self).__init__()
        self.filter = FilterLayer(2 * in_planes, out_planes, reduction)

    def forward(self, guidePath, mainPath):
        _inference_only(self, guidePath, mainPath)
        fc = self.filter.fc
        return torch.ops.kpf.fsp(guidePath, mainPath, fc[0].weight, fc[0].bias, fc[2].weight, fc[2].bias)


class RGBDFusion(nn.Module):
    # model/fusion_layer.py:40-83
    def __init__(self, in_planes, out_planes, reduction=16, bn_momentum=0.0003):
        self.init__ = super(RGBDFusion, self).__init__()
        self.in_planes = in_planes
        self.bn_momentum = bn_momentum
        self.fsp_rgb = FSP(in_planes, out_planes, reduction)      # constructed but unused by forward, like the reference
        self.fsp_depth = FSP(in_planes, out_planes, reduction)
        self.gate_rgb = nn.Conv2d(in_planes * 2, 1, kernel_size=1, bias=True)
        self.gate_depth = nn.Conv2d(in_planes * 2, 1, kernel_size=1, bias=True)
        self.relu1 = nn.ReLU()
        self.relu2 = nn.ReLU()
        self.softmax = nn.Softmax(dim=1)

    def forward(self, x, train_writer=None, global_step=0, layer_stage=0):
        rgb, depth = x
        _inference_only(self, rgb, depth)
        gw = torch.cat([self.gate_rgb.weight.reshape(1, -1), self.gate_depth.weight.reshape(1, -1)], 0)
        gb = torch.cat([self.gate_rgb.bias, self.gate_depth.bias])
        if train_writer is None:
            rgb_out, depth_out, merge = torch.ops.kpf.rgbd_fusion(rgb, depth, gw, gb)
            return [rgb_out, depth_out], merge
        rgb_out, depth_out, merge, amean = ops.rgbd_fusion(rgb, depth, gw, gb, want_attn_mean=True)
        if train_writer is not None:  # model/fusion_layer.py:68-72 (a full reduction + host sync: optional slow path)
            train_writer.add_scalar('RGB_weight_fusion_stage{}'.format(layer_stage), amean[0].detach(), global_step)
            train_writer.add_scalar('Depth_weight_fusion_stage{}'.format(layer_stage), amean[1].detach(), global_step)
        return [rgb_out, depth_out], merge


class ACFusion(nn.Module):
    # model/fusion_layer.py:87-116
    def __init__(self, in_planes, out_planes, reduction=16, bn_momentum=0.0003):
        self.init__ = super(ACFusion, self).__init__()
        self.in_planes = in_planes
        self.bn_momentum = bn_momentum
        self.cam_rgb = nn.Conv2d(in_planes, in_planes, kernel_size=1, bias=True)
        self.cam_depth = nn.Conv2d(in_planes, in_planes, kernel_size=1, bias=True)
        self.sigmoid = nn.Sigmoid()
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.relu1 = nn.ReLU()
        self.relu2 = nn.ReLU()

    def forward(self, x, train_writer=None, global_step=0, layer_stage=0):
        rgb, depth = x
        _inference_only(self, rgb, depth)
        rgb_out, depth_out, merge = torch.ops.kpf.ac_fusion(rgb, depth, self.cam_rgb.weight, self.cam_rgb.bias, self.cam_depth.weight,
                                                            self.cam_depth.bias)
        return [rgb_out, depth_out], merge
