"""Drop-in for the reference's util/img2pcl.py (Pcl_utils).  The reference class is an unfinished torch port that
crashes (util/img2pcl.py:24, :32-33, :53); this keeps its batched signature (util/img2pcl.py:11) and gives it the
semantics of the numpy code that actually runs (dataloader/loader.py:843-893, :1173-1186) on the B200 kernel K1."""
import torch

from .. import ops


class Pcl_utils(object):
    def __init__(self, seed=0, clamp=False):
        self.flip = 1            # util/img2pcl.py:8
        self.sample_num = 1024   # util/img2pcl.py:9
        self.seed = seed         # the reference draws from np.random (loader.py:1182-1184); here a counter-based permutation
        self.clamp = clamp       # HO3D / demo clamp to [-1,1] (loader.py:1399, demo_RGBD.py:332)
        self.calls = 0

    def getpcl(self, imgD, com3D, cube, M, cam_para, select=None):
        """imgD [B,1,S,S] (or [B,S,S]), com3D [B,3], cube [B,3], M [B,3,3], cam_para [B,4] -> [B,sample_num,3].
        `select` [B,sample_num] optional explicit ranks into the ordered valid-pixel list."""
        if imgD.dim() == 3:
            imgD = imgD.unsqueeze(1)
        pcl, self.last_count = ops.getpcl(imgD, com3D, cube, M, cam_para, self.sample_num, ranks=select,
                                          seed=self.seed + self.calls, clamp=self.clamp, flip=self.flip)
        self.calls += 1
        return pcl

    def depthTopcl(self, dpt, T, paras, background_val=torch.tensor(0.)):
        """dpt [B,S,S] depth crop in mm (0 = invalid), T [B,3,3] crop transform, paras [B,4] (fx,fy,fu,fv) -> (xyz [B,S*S,3] camera-space
        mm, count [B]): every valid pixel back-projected like the numpy depthToPCL the reference actually runs (loader.py:874-893; its
        torch port util/img2pcl.py:42-64 crashes), ordered row-major; rows past the per-sample count are zero.  Runs kpf_backproject_all
        with centre 0 and cube 2 (so the normalisation is the identity); a pixel within 1e-5 of exactly 1.0 mm counts as background."""
        B = dpt.shape[0]
        zeros = torch.zeros(B, 3, device=dpt.device)
        two = torch.full((B, 3), 2.0, device=dpt.device)
        # dpt(mm) = img * cube_z/2 + com_z with cube=2, com=0 -> identity; background_val 0 == invalid
        img = torch.where(dpt == 0, torch.ones_like(dpt), dpt)
        xyz, pix, count = ops.backproject_all(img.reshape(B, 1, dpt.shape[-2], dpt.shape[-1]), zeros, two, T, paras, flip=self.flip)
        return xyz, count
