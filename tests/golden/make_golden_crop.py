"""Golden fixtures for the crop front end: runs the reference's own demo_RGBD.py methods (unbound, on a stub
object; cv2 is available in the build container) on the repo's only real RGB-D frame (visualization/box*.png,
bbox demo_RGBD.py:578, intrinsics :585) and on synthetic 640x480 frames.  Writes golden_crop.npz (frames are stored
down-cropped around the hand to keep the fixture small).

    python tests/golden/make_golden_crop.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

ref_shims.install()
import cv2  # noqa: E402
import demo_RGBD as D  # noqa: E402


def stub(cam, cube):
    s = types.SimpleNamespace(cam_para=cam, img_size=128, cube=cube, flip=1, sample_num=1024)
    for name in ("get_center_from_bbx", "Crop_Image_deep_pp", "Crop_Image_deep_pp_RGB", "comToBounds", "getCrop", "normalize_img",
                 "jointImgTo3D"):
        setattr(s, name, types.MethodType(getattr(D.Model_RGBD, name), s))
    return s


def run(s, rgb, depth, bbox):
    center = s.get_center_from_bbx(depth, bbox)
    crop_rgb, _ = s.Crop_Image_deep_pp_RGB(rgb, center, s.cube, (128, 128), s.cam_para)
    crop_rgb = crop_rgb.astype(np.float32).transpose(2, 0, 1) / np.float32(255.)        # ToTensor() on float32 HWC, then /255
    crop_d, M = s.Crop_Image_deep_pp(depth, center, s.cube, (128, 128), s.cam_para)
    imgD = s.normalize_img(crop_d.max(), crop_d, center, s.cube)
    com3d = s.jointImgTo3D(center)
    return center, crop_rgb, imgD.astype(np.float32), M, np.asarray(com3d, np.float64)


def main():
    out = {}
    # ---- the repo's demo frame (1920x1080); keep a 512x512 window around the hand, shift bbox / principal point accordingly
    rgb = cv2.imread(os.path.join(ref_shims.REF, "visualization/box.png"))
    depth = cv2.imread(os.path.join(ref_shims.REF, "visualization/box_d.png"), cv2.IMREAD_ANYDEPTH)
    x0, y0 = 640, 256
    rgb, depth = np.ascontiguousarray(rgb[y0:y0 + 512, x0:x0 + 512]), np.ascontiguousarray(depth[y0:y0 + 512, x0:x0 + 512])
    bbox = [885 - 178.0 / 2 - x0, 515.5 - 127.0 / 2 - y0, 178.0, 127.0]
    cam = (906.96, 906.79, 956.75 - x0, 547.23 - y0)
    cube = [250, 250, 250]
    c, cr, cd, M, c3 = run(stub(cam, cube), rgb, depth, bbox)
    out.update(box_rgb=rgb, box_depth=depth, box_bbox=np.array(bbox), box_cam=np.array(cam), box_center=c, box_crop_rgb=cr, box_crop_d=cd,
               box_M=M, box_com3d=c3)
    # ---- synthetic 640x480 frames (BASELINE config 5 shape): hand blob at different places incl. image borders
    rs = np.random.RandomState(3)
    frames, bbs, res = [], [], []
    cam2 = (617.0, 617.0, 312.0, 241.0)
    for i, (cx, cy, dist) in enumerate([(320, 240, 600), (40, 60, 450), (610, 450, 800)]):
        d = np.zeros((480, 640), np.uint16)
        yy, xx = np.mgrid[0:480, 0:640]
        r = 60000 // dist
        m = (xx - cx) ** 2 + (yy - cy) ** 2 < r * r
        d[m] = (dist + rs.randint(-60, 60, size=m.sum())).astype(np.uint16)
        d[rs.rand(480, 640) < 0.004] = rs.randint(200, 3000)           # clutter, partly outside the cube
        bgr = np.stack([(xx // 8 * 13 + yy // 8 * 7 + 17 * i) % 256, (xx // 16 * 31 + yy // 4) % 256, (xx // 4 * 5 + yy // 16 * 11) % 256],
                       -1).astype(np.uint8)  # blocky pattern: every nearest-resize index error shows, and it compresses
        bb = [cx - r - 8.5, cy - r - 3.0, 2 * r + 17.0, 2 * r + 6.0]
        bb[0], bb[1] = max(bb[0], 0.0), max(bb[1], 0.0)
        frames.append((bgr, d))
        bbs.append(bb)
        res.append(run(stub(cam2, cube), bgr, d, bb))
    out.update(syn_rgb=np.stack([f[0] for f in frames]), syn_depth=np.stack([f[1] for f in frames]), syn_bbox=np.array(bbs),
               syn_cam=np.array(cam2), syn_center=np.stack([r[0] for r in res]), syn_crop_rgb=np.stack([r[1] for r in res]),
               syn_crop_d=np.stack([r[2] for r in res]), syn_M=np.stack([r[3] for r in res]), syn_com3d=np.stack([r[4] for r in res]))
    np.savez_compressed(os.path.join(HERE, "golden_crop.npz"), **out)
    print("wrote golden_crop.npz", os.path.getsize(os.path.join(HERE, "golden_crop.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
