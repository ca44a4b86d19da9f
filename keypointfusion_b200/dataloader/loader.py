"""Drop-in for the tensor geometry helpers of the reference's `loader` object (dataloader/loader.py:752-1005), i.e.
the object KPFusion.forward receives as its `loader` argument (model/model.py:395, train.py:324)."""
import torch

from .. import custom_ops  # noqa: F401  (registers torch.ops.kpf.*)
from .. import ops


class loader(object):
    """Only the tensor helpers the hot path calls; dataset I/O stays with the reference's dataloader."""

    def __init__(self, root_dir='', phase='test', img_size=128, center_type='joint_mean', dataset_name='x', flip=1):
        self.img_size = img_size
        self.flip = flip      # every dataset class sets 1 (loader.py:1033, :1222, :1510)
        self.phase = phase
        self.dataset_name = dataset_name

    # dataloader/loader.py:775-789
    def uvd_nl2xyznl_tensor(self, uvd, center, m, cube, cam_paras):
        return torch.ops.kpf.uvd2xyz(uvd, center, m, cube, cam_paras, float(self.img_size), float(self.flip))

    # dataloader/loader.py:821-834
    def xyz_nl2uvdnl_tensor(self, joint_xyz, center, M, cube_size, cam_paras):
        return torch.ops.kpf.xyz2uvd(joint_xyz, center, M, cube_size, cam_paras, float(self.img_size), float(self.flip))

    # dataloader/loader.py:936-967
    def img2pcl_index(self, pcl, img, center, M, cube, cam_para, select_num=9):
        return torch.ops.kpf.img2pcl_index(pcl, img, center, M, cube, cam_para, float(self.img_size), select_num, float(self.flip), True, None)

    # dataloader/loader.py:791-819
    def img2anchor_dis(self, joint_uvd, img, center, M, cube, cam_para, gamma=10):
        return torch.ops.kpf.img2anchor_dis(joint_uvd, img, center, M, cube, cam_para, float(self.img_size), float(gamma), float(self.flip))

    # dataloader/loader.py:993-1005
    def img2pcl(self, img):
        B, _, W, H = img.shape
        t = 2.0 * (torch.arange(W, device=img.device, dtype=torch.float32) + 0.5) / W - 1.0
        u = t.view(1, 1, 1, W).expand(B, 1, W, W)
        v = t.view(1, 1, W, 1).expand(B, 1, W, W)
        return torch.cat((u, v, img), dim=1).view(B, 3, H * W).permute(0, 2, 1)


GeometryHelper = loader
