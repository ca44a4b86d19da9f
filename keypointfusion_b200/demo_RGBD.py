"""Drop-in for the per-frame numpy front end of the reference's demo_RGBD.py (Model_RGBD, demo_RGBD.py:65-173), batched on the
GPU: bbox -> centre, RGB / depth crops with the reference's exact integer geometry (comToBounds, cv2 INTER_NEAREST, centred
paste), depth normalisation, back-projection + 1024-point resample + clamp.  SURVEY.md 8f-3 / BASELINE.json config 5."""
import torch

from . import ops


class Model_RGBD(object):
    """Front end + (optionally) the network.  `net`: a keypointfusion_b200.model.model.KPFusion with backbones, or None when
    only the preprocessing is wanted.  `cam_para` = (fx, fy, fu, fv) as in demo_RGBD.py:585."""

    def __init__(self, net=None, cam_para=(617.0, 617.0, 312.0, 241.0), cube=(250, 250, 250), img_size=128, sample_num=1024, seed=0):
        self.net, self.cam_para, self.cube, self.img_size, self.sample_num, self.seed = net, tuple(cam_para), list(cube), img_size, sample_num, seed
        self.flip = 1
        from .dataloader.loader import loader
        self.depthloader = loader(img_size=img_size)   # the geometry helper the reference passes to net() (demo_RGBD.py:111)

    # demo_RGBD.py:253-276 (batched: depth [B,Hf,Wf] uint16 on the GPU, bbx [B,4] xywh)
    def get_center_from_bbx(self, depth, bbx, upper=1500, lower=171):
        return ops.center_from_bbox(depth, bbx, upper, lower)

    # demo_RGBD.py:464-517 + ToTensor()/255 (:87)
    def Crop_Image_deep_pp_RGB(self, rgb, com, size=None, dsize=None, paras=None):
        return ops.crop_rgb(rgb, com, size or self.cube, paras or self.cam_para, (dsize or (self.img_size,))[0])

    # demo_RGBD.py:305-343
    def process_depth(self, cube_size, depth, center):
        img, M, com3d, cube = ops.crop_depth(depth, center, cube_size, self.cam_para, self.img_size)
        cam = torch.tensor(self.cam_para, device=img.device, dtype=torch.float32).expand(img.shape[0], 4).contiguous()
        pcl, _ = ops.getpcl(img, com3d, cube, M, cam, self.sample_num, seed=self.seed, clamp=True, flip=self.flip)  # :319-332
        return img, pcl, com3d, M, cube

    def prepare_batch(self, rgb, depth, bbox):
        """rgb [B,Hf,Wf,3] uint8 (BGR), depth [B,Hf,Wf] uint16, bbox [B,4] -> the nine tensors KPFusion.forward consumes."""
        center_uvd = self.get_center_from_bbx(depth, bbox)
        img_rgb = self.Crop_Image_deep_pp_RGB(rgb, center_uvd)
        img, pcl, com3d, M, cube = self.process_depth(self.cube, depth, center_uvd)
        cam = torch.tensor(self.cam_para, device=img.device, dtype=torch.float32).expand(img.shape[0], 4).contiguous()
        return dict(img_rgb=img_rgb, img=img, pcl=pcl, center=com3d, M=M, cube=cube, cam_para=cam, center_uvd=center_uvd)

    def estimate_pose_RGBD(self, rgb, depth, bbox):
        """Batched demo_RGBD.py:65-131: returns (result list of KPFusion.forward, batch dict)."""
        if self.net is None:
            raise RuntimeError("Model_RGBD was built without a network")
        b = self.prepare_batch(rgb, depth, bbox)
        with torch.no_grad():
            result, spatial_weight, _ = self.net(b["img_rgb"], b["img"], b["pcl"], self.depthloader, b["center"], b["M"], b["cube"], b["cam_para"], 0.8)
        return result, b
