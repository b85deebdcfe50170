"""Host mirror of the per-frame ray set-up: what `ZJUMOCAPODPDataset.__getitem__`
(im2mesh/data/zju_mocap_odp.py:250-315) does with numpy and cv2 per frame, as two C-ABI calls on the GPU.

    fr = FrameRays(device)
    verts, bounds = fr.pose_smpl(minimal_shape, posedirs, pose_feature, skinning_weights, bone_transforms, trans)
    out = fr.gen_rays(K, R, T, bounds, H, W)            # dict: pix, ray_dirs, near_far, image_mask, bound_mask, cam_loc, n_rays

`verts` is ArahFrame.smpl_verts, `out['ray_dirs'] / out['near_far']` are the renderer's `ray_dirs` / `body_bounds_intersections`.
The 3x3 camera algebra (K_inv, cam_loc) stays on the host in numpy, exactly as the reference computes it (:216,231).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _host3(a, n):
    a = np.ascontiguousarray(np.asarray(a, np.float32).reshape(-1))
    assert a.size == n
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


class FrameRays:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.ArahError('the ray set-up only exists as CUDA kernels; got device %s' % device)
        self._ws_pose = torch.empty(64, dtype=torch.uint8, device=self.device)

    @property
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def pose_smpl(self, minimal_shape, posedirs, pose_feature, skinning_weights, bone_transforms, trans, box_margin=0.05):
        dev = self.device
        f32 = lambda a: torch.as_tensor(a, dtype=torch.float32).to(dev).contiguous()
        ms, pd, w, B = f32(minimal_shape).view(-1, 3), f32(posedirs).view(-1, 207), f32(skinning_weights), f32(bone_transforms).view(-1, 16)
        pf = torch.as_tensor(np.asarray(pose_feature, np.float64).reshape(-1)).to(dev).contiguous()
        n = ms.shape[0]
        assert pd.shape[0] == 3 * n and w.shape == (n, 24) and B.shape[0] == 24 and pf.numel() == 207
        verts, bounds = torch.empty(n, 3, device=dev), torch.empty(2, 3, device=dev)
        tr, trp = _host3(trans, 3)
        check(_lib.lib().arah_pose_smpl(_ptr(ms), _ptr(pd), _ptr(pf), _ptr(w), _ptr(B), trp, n, float(box_margin), _ptr(verts), _ptr(bounds),
                                        _ptr(self._ws_pose), self._stream))
        self._keep = (ms, pd, pf, w, B, tr)
        return verts, bounds

    def gen_rays(self, K, R, T, bounds, H, W, mask=None):
        dev = self.device
        K = np.asarray(K, np.float32).reshape(3, 3); R = np.asarray(R, np.float32).reshape(3, 3); T = np.asarray(T, np.float32).reshape(3)
        K_inv = np.linalg.inv(K)                                   # zju_mocap_odp.py:231
        cam_loc = np.dot(-R.T, T)                                  # :171-173
        Kh, Kp = _host3(K, 9); Ki, Kip = _host3(K_inv, 9); Rh, Rp = _host3(R, 9); Th, Tp = _host3(T, 3); ch, cp = _host3(cam_loc, 3)
        bounds = torch.as_tensor(bounds, dtype=torch.float32).to(dev).contiguous()
        n = H * W
        m_in = None if mask is None else torch.as_tensor(mask).to(dev).reshape(-1).to(torch.uint8).contiguous()
        bound_mask = torch.empty(n, dtype=torch.uint8, device=dev)
        pix = torch.empty(n, dtype=torch.int32, device=dev)
        dirs, nf = torch.empty(n, 3, device=dev), torch.empty(n, 2, device=dev)
        image_mask = torch.empty(n, dtype=torch.uint8, device=dev)
        count = torch.zeros(1, dtype=torch.int32, device=dev)
        ws_bytes = int(_lib.lib().arah_frame_rays_workspace(H, W))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        check(_lib.lib().arah_frame_rays(Kp, Kip, Rp, Tp, cp, _ptr(bounds), H, W, _ptr(m_in), _ptr(bound_mask), _ptr(pix), _ptr(dirs), _ptr(nf),
                                         _ptr(image_mask), _ptr(count), _ptr(ws), ws_bytes, self._stream))
        P = int(count.item())                                      # the one synchronisation: the ray count sizes the views
        return {'pix': pix[:P], 'ray_dirs': dirs[:P], 'near_far': nf[:P], 'image_mask': image_mask.view(H, W).bool(),
                'bound_mask': (m_in if m_in is not None else bound_mask).view(H, W), 'cam_loc': torch.from_numpy(cam_loc.astype(np.float32)).to(dev),
                'n_rays': P}
