"""Host mirror of the image-space tail of the reference's validation / test step (after the renderer).

    fi = FrameImages(device)
    pred_pixels, pred_normals = fi.assemble(out['rgb_values'], out['points_cam'], rays['pix'], H, W)   # lightning_model.py:176-205
    mse, psnr = fi.psnr(out['rgb_values'], gt_rays)                                                     # :218-221, utils/eval.py:6-9
    ssim = fi.ssim(pred_pixels, gt_pixels, image_mask)                                                  # :222, utils/eval.py:11-19
    maps = fi.normal_maps(verts_cano, faces, verts_posed, cam_rot, cam_trans, K)                        # models/__init__.py:226-309
    # maps: 'output_normal', 'normal_cano_front', 'normal_cano_back' — [1, H, W, 3] in [0, 1], the keys the reference adds

Everything runs as CUDA kernels behind the C ABI (csrc/arah_image.cu); there is no CPU path.  The 3x3 camera algebra
(`look_at_view_transform`, the OpenCV -> pytorch3d conversion) stays on the host in numpy, like the reference's camera set-up.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check

F = np.float32


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


# ----------------------------------------------------------------------------------------------------------------- cameras
def _normalize(v, eps=1e-5):
    v = np.asarray(v, np.float32)
    return v / np.maximum(np.linalg.norm(v).astype(np.float32), F(eps))


def look_at_view_transform(dist=1.0, elev=0.0, azim=0.0):
    """pytorch3d.renderer.look_at_view_transform(dist, elev, azim) in degrees, at = origin, up = +Y -> R [3,3], T [3]
    (row-vector convention: X_view = X_world R + T).  Used at models/__init__.py:265 (front) and :291 (back)."""
    e, a = F(np.pi / 180.0) * F(elev), F(np.pi / 180.0) * F(azim)
    cam = np.array([F(dist) * np.cos(e) * np.sin(a), F(dist) * np.sin(e), F(dist) * np.cos(e) * np.cos(a)], np.float32)
    z = _normalize(-cam)
    x = _normalize(np.cross(np.array([0, 1, 0], np.float32), z))
    y = _normalize(np.cross(z, x))
    if np.allclose(x, 0.0, atol=5e-3):
        x = _normalize(np.cross(y, z))
    R = np.stack([x, y, z], axis=1).astype(np.float32)
    return R, -(cam @ R).astype(np.float32)


def fov_perspective_camera(R, T, fov=60.0):
    """FoVPerspectiveCameras(R=R, T=T) with its defaults (fov 60 degrees, aspect 1)."""
    s = F(1.0) / np.tan(F(np.pi / 180.0) * F(fov) / F(2), dtype=np.float32)
    return _camera(R, T, s, s, 0.0, 0.0)


def opencv_camera(cam_rot, cam_trans, K, H, W):
    """pytorch3d.utils.cameras_from_opencv_projection(cam_rot, cam_trans, K, image_size=(H, W)) (models/__init__.py:247-252):
    focal / principal point in NDC units of the shorter image side, X and Y axes flipped."""
    K = np.asarray(K, np.float32).reshape(3, 3)
    scale = F(min(H, W)) / F(2)
    fx, fy = K[0, 0] / scale, K[1, 1] / scale
    px, py = -(K[0, 2] - F(W) / F(2)) / scale, -(K[1, 2] - F(H) / F(2)) / scale
    R = np.asarray(cam_rot, np.float32).reshape(3, 3).T.copy()
    T = np.asarray(cam_trans, np.float32).reshape(3).copy()
    R[:, :2] *= -1
    T[:2] *= -1
    return _camera(R, T, fx, fy, px, py)


def _camera(R, T, fx, fy, px, py):
    c = _lib.ArahRasterCamera()
    c.R[:] = [float(v) for v in np.asarray(R, np.float32).reshape(9)]
    c.T[:] = [float(v) for v in np.asarray(T, np.float32).reshape(3)]
    c.fx, c.fy, c.px, c.py = float(fx), float(fy), float(px), float(py)
    return c


# ----------------------------------------------------------------------------------------------------------------- kernels
class FrameImages:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.ArahError('the image tail only exists as CUDA kernels; got device %s' % device)
        self._ws = {}

    @property
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, key, nbytes):
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = self._ws[key] = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
        return ws

    def _f32(self, t, cols=3):
        return torch.as_tensor(t, dtype=torch.float32).to(self.device).reshape(-1, cols).contiguous()

    def assemble(self, rgb_values, points_cam, pix, H, W, normals=True):
        """-> pred_pixels [H,W,3], pred_normals [H,W,3] (None if normals=False)."""
        rgb, pts = self._f32(rgb_values), (self._f32(points_cam) if normals else None)
        pix = torch.as_tensor(pix).to(self.device).reshape(-1).to(torch.int32).contiguous()
        P = int(pix.numel())
        if rgb.shape[0] < P or (normals and pts.shape[0] < P):
            raise _lib.ArahError('fewer rows than mask pixels')
        pred_pixels = torch.empty(H, W, 3, device=self.device)
        pred_normals = torch.empty(H, W, 3, device=self.device) if normals else None
        nb = int(_lib.lib().arah_frame_images_workspace(H, W))
        ws = self._workspace('img', nb)
        check(_lib.lib().arah_frame_images(_ptr(rgb), _ptr(pts), _ptr(pix), P, H, W, _ptr(pred_pixels), _ptr(pred_normals), _ptr(ws), ws.numel(),
                                           self._stream))
        self._keep = (rgb, pts, pix)
        return pred_pixels, pred_normals

    def psnr_device(self, pred, gt):
        """-> float64 device tensor [2] = (mse, psnr); no synchronisation."""
        a = torch.as_tensor(pred, dtype=torch.float32).to(self.device).reshape(-1).contiguous()
        b = torch.as_tensor(gt, dtype=torch.float32).to(self.device).reshape(-1).contiguous()
        if a.numel() != b.numel():
            raise _lib.ArahError('psnr: size mismatch')
        out = torch.empty(2, dtype=torch.float64, device=self.device)
        ws = self._workspace('psnr', int(_lib.lib().arah_psnr_workspace()))
        check(_lib.lib().arah_psnr(_ptr(a), _ptr(b), a.numel(), _ptr(out), _ptr(ws), ws.numel(), self._stream))
        self._keep_psnr = (a, b)
        return out

    def psnr(self, pred, gt):
        m, p = self.psnr_device(pred, gt).tolist()                     # the one synchronisation (the reference returns a Python float)
        return m, p

    def ssim_device(self, pred_pixels, gt_pixels, mask):
        """-> float64 device tensor [5] = (ssim, x, y, w, h of the mask's bounding rectangle); no synchronisation."""
        a = torch.as_tensor(pred_pixels, dtype=torch.float32).to(self.device).contiguous()
        b = torch.as_tensor(gt_pixels, dtype=torch.float32).to(self.device).contiguous()
        if a.dim() != 3 or a.shape[-1] != 3 or a.shape != b.shape:
            raise _lib.ArahError('ssim: two [H, W, 3] images expected')
        H, W = int(a.shape[0]), int(a.shape[1])
        m = torch.as_tensor(mask).to(self.device).reshape(-1).to(torch.uint8).contiguous()
        if m.numel() != H * W:
            raise _lib.ArahError('ssim: mask size')
        out = torch.empty(5, dtype=torch.float64, device=self.device)
        ws = self._workspace('ssim', int(_lib.lib().arah_ssim_workspace()))
        check(_lib.lib().arah_ssim(_ptr(a), _ptr(b), _ptr(m), H, W, _ptr(out), _ptr(ws), ws.numel(), self._stream))
        self._keep_ssim = (a, b, m)
        return out

    def ssim(self, pred_pixels, gt_pixels, mask):
        """`ssim_metric(pred_pixels, gt_pixels, bbox_mask)` (im2mesh/utils/eval.py:11-19) -> Python float."""
        v = float(self.ssim_device(pred_pixels, gt_pixels, mask)[0].item())
        if v != v:
            raise ValueError('win_size exceeds image extent')            # what skimage raises for a crop smaller than 7 x 7
        return v

    def validation_tail(self, model_outputs, batch):
        """The body of `LightningModel.validation_step` after the model call (lightning_model.py:176-229) minus SSIM / LPIPS:
        reads `batch['inputs.img_height' / 'inputs.img_width' / 'inputs.image_mask' / 'inputs']` and `model_outputs['rgb_values' /
        'points_cam']` (or a ready 'output_normal') -> {'psnr', 'ssim' (floats), 'rgb_pred', 'normal_pred', 'rgb_gt'} with the images
        channel-first [3, H, W] as the reference returns them (:225-227)."""
        H, W = int(batch['inputs.img_height'].item()), int(batch['inputs.img_width'].item())
        mask = torch.as_tensor(batch['inputs.image_mask']).to(self.device).reshape(-1)
        pix = mask.nonzero().squeeze(1).to(torch.int32)                          # np.where order == masked_scatter_ order
        n = int(pix.numel())
        rgb = torch.as_tensor(model_outputs['rgb_values']).reshape(-1, 3)[:n]
        gt = torch.as_tensor(batch['inputs']).reshape(-1, 3)[:n]
        want_normals = 'output_normal' not in model_outputs
        pred_pixels, pred_normals = self.assemble(rgb, model_outputs['points_cam'].reshape(-1, 3)[:n] if want_normals else None, pix, H, W,
                                                  normals=want_normals)
        if not want_normals:
            pred_normals = model_outputs['output_normal'].squeeze(0)
        gt_pixels, _ = self.assemble(gt, None, pix, H, W, normals=False)
        # psnr_metric runs on the FULL ray lists, not on the [:n] slices (:218-221)
        psnr = self.psnr(model_outputs['rgb_values'].reshape(-1, 3), torch.as_tensor(batch['inputs']).reshape(-1, 3))[1]
        ssim = self.ssim(pred_pixels, gt_pixels, mask)                            # :222 (LPIPS, a VGG network, stays with the caller)
        return {'psnr': psnr, 'ssim': ssim, 'rgb_pred': pred_pixels.permute(2, 0, 1), 'normal_pred': pred_normals.permute(2, 0, 1), 'rgb_gt': gt_pixels.permute(2, 0, 1)}

    def rasterize(self, verts, faces, camera, H=512, W=512, zbuf=False):
        """-> pix_to_face [H,W] int32 (and zbuf [H,W] if asked)."""
        v = self._f32(verts)
        f = torch.as_tensor(faces).to(self.device).reshape(-1, 3).to(torch.int32).contiguous()
        p2f = torch.empty(H, W, dtype=torch.int32, device=self.device)
        zb = torch.empty(H, W, device=self.device) if zbuf else None
        nb = int(_lib.lib().arah_rasterize_mesh_workspace(max(v.shape[0], 1), H, W))
        ws = self._workspace('raster', nb)
        check(_lib.lib().arah_rasterize_mesh(_ptr(v), v.shape[0], _ptr(f), f.shape[0], C.byref(camera), H, W, _ptr(p2f), _ptr(zb), _ptr(ws), ws.numel(),
                                             self._stream))
        self._keep_r = (v, f)
        return (p2f, zb) if zbuf else p2f

    def normal_image(self, verts, faces, pix_to_face, sign=1.0, rot=None, background=0.0):
        v = self._f32(verts)
        f = torch.as_tensor(faces).to(self.device).reshape(-1, 3).to(torch.int32).contiguous()
        H, W = pix_to_face.shape
        pix_to_face = pix_to_face.to(self.device, torch.int32).contiguous()
        img = torch.empty(H, W, 3, device=self.device)
        rp = None
        if rot is not None:
            r = np.ascontiguousarray(np.asarray(torch.as_tensor(rot).detach().cpu().numpy(), np.float32).reshape(9))
            rp = r.ctypes.data_as(C.POINTER(C.c_float))
        check(_lib.lib().arah_face_normal_image(_ptr(v), v.shape[0], _ptr(f), f.shape[0], _ptr(pix_to_face), H, W, float(sign), rp,
                                                float(background), _ptr(img), self._stream))
        self._keep_n = (v, f, pix_to_face)
        return img

    def normal_maps(self, verts_cano, faces, verts_posed, cam_rot, cam_trans, K, H=512, W=512):
        """The three images `MetaAvatarRender.forward(gen_cano_mesh=True)` adds (models/__init__.py:226-309): the posed mesh seen
        by the frame's camera (normals negated and rotated into the camera frame, background -1) and the canonical mesh from
        the front / back (FoV camera at distance 2, background 0), all mapped with (n + 1) / 2 and clipped."""
        t = lambda a: np.asarray(torch.as_tensor(a).detach().cpu().numpy(), np.float32)
        cam_rot, cam_trans, K = t(cam_rot).reshape(3, 3), t(cam_trans).reshape(3), t(K).reshape(3, 3)
        out = {}
        p2f = self.rasterize(verts_posed, faces, opencv_camera(cam_rot, cam_trans, K, H, W), H, W)
        out['output_normal'] = self.normal_image(verts_posed, faces, p2f, -1.0, cam_rot, -1.0).unsqueeze(0)
        for name, azim in (('normal_cano_front', 0.0), ('normal_cano_back', 180.0)):
            R, T = look_at_view_transform(2.0, 0.0, azim)
            p2f = self.rasterize(verts_cano, faces, fov_perspective_camera(R, T), H, W)
            out[name] = self.normal_image(verts_cano, faces, p2f, 1.0, None, 0.0).unsqueeze(0)
        return out
