"""Drop-in for the live subset of the reference's util/generateFeature.py (GFM) on the B200 kernels.
rigid_align is eval-time numpy (generateFeature.py:681-703) and stays numpy."""
import numpy as np
import torch

from .. import ops


class GFM:
    def __init__(self):
        self.softmax = torch.nn.Softmax(dim=-1)

    # util/generateFeature.py:584-600
    def joint2heatmap(self, joint, std, heatmap_size, sigma=1.5):
        return ops.joint2heatmap(joint, std, heatmap_size, sigma)

    # util/generateFeature.py:59-84
    def joint2offset(self, joint, img, kernel_size, feature_size):
        return ops.joint2offset(joint, img, kernel_size, feature_size, eps=1e-8)

    # util/generateFeature.py:166-195
    def offset2joint_weight(self, offset, depth, kernel_size):
        return ops.offset2joint_weight(offset, depth, kernel_size)

    # util/generateFeature.py:465-488
    def pcl_joint2offset(self, joint, pcl, kernel_size):
        return ops.pcl_joint2offset(joint, pcl, kernel_size)

    # util/generateFeature.py:490-517 (== model/model.py:528-555)
    def pcl_offset2joint_weight(self, pcl_result, pcl, kernel_size):
        from ..model.model import pcl_offset2joint_weight
        return pcl_offset2joint_weight(pcl_result, pcl, kernel_size)

    # util/generateFeature.py:398-431 (dispatcher; only the offset family is live: config.py:72)
    def joint2feature(self, joint, img, feature_paras, feature_size, feature_types):
        feats = []
        for i, ft in enumerate(feature_types):
            if ft in ('offset', 'weight_offset', 'weight_offset_nosoftmax'):
                feats.append(self.joint2offset(joint, img, feature_paras[i], feature_size))
            else:
                raise NotImplementedError(f"feature type {ft!r} is an unreferenced ablation variant in the reference")
        return torch.cat(feats, dim=1)

    # util/generateFeature.py:434-462
    def feature2joint(self, img, pixel_pd, feature_types, feature_paras):
        joint = None
        for i, ft in enumerate(feature_types):
            if ft == 'weight_offset':
                joint = self.offset2joint_weight(pixel_pd, img, feature_paras[i])
            else:
                raise NotImplementedError(f"feature type {ft!r} is an unreferenced ablation variant in the reference")
        return joint

    # util/generateFeature.py:681-703 (numpy, eval only)
    def rigid_transform_3D(self, A, B):
        n, dim = A.shape
        ca, cb = np.mean(A, axis=0), np.mean(B, axis=0)
        H = np.dot(np.transpose(A - ca), B - cb) / n
        U, s, V = np.linalg.svd(H)
        R = np.dot(np.transpose(V), np.transpose(U))
        if np.linalg.det(R) < 0:
            s[-1] = -s[-1]
            V[2] = -V[2]
            R = np.dot(np.transpose(V), np.transpose(U))
        c = 1 / np.var(A, axis=0).sum() * np.sum(s)
        t = -np.dot(c * R, np.transpose(ca)) + np.transpose(cb)
        return c, R, t

    def rigid_align(self, A, B):
        c, R, t = self.rigid_transform_3D(A, B)
        return np.transpose(np.dot(c * R, np.transpose(A))) + t
