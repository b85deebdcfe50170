"""CPU: the marshalling of arah_release_b200/images.py (argument order, dtypes, contiguity, workspace sizing, result shapes).

The build container has no GPU, so the product class cannot run here.  This test swaps the loaded C ABI for a shim with the
SAME entry-point signatures whose bodies call the host instantiation of the kernels' arithmetic (tests/native/host_image.cpp),
feeds the product's `FrameImages` methods CPU tensors, and compares with the oracle.  Test infrastructure only: it checks the
Python mirror's plumbing, not the CUDA kernels (tests/test_gpu_zz_images.py does that on the GPU box).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers_images import iso_mesh, load_images_golden, make_camera
from oracle import images_oracle as io
from test_images_host import hi  # noqa: F401  (fixture: builds / loads the host harness)

FP, IP = C.POINTER(C.c_float), C.POINTER(C.c_int32)


def _fp(p):
    return C.cast(p, FP) if p is not None and (not hasattr(p, 'value') or p.value) else None


class ShimLib:
    """Same call signatures as libarah_b200.so's image entry points; host arithmetic from tests/native/host_image.cpp."""
    def __init__(self, h):
        self.h = h
        self.calls = []

    def arah_last_error(self):
        return b'invalid argument (shim)'

    def arah_frame_images_workspace(self, H, W):
        return H * W * 12

    def arah_frame_images(self, rgb, pts, pix, P, H, W, pp, pn, ws, ws_bytes, stream):
        assert ws_bytes >= H * W * 12
        self.calls.append('frame_images')
        n = H * W * 3
        tmp_pp, tmp_pn, zero = (C.c_float * n)(), (C.c_float * n)(), (C.c_float * max(3 * P, 1))()
        self.h.host_frame_images(_fp(rgb) or zero, _fp(pts) or zero, C.cast(pix, IP), P, H, W, _fp(pp) or tmp_pp, _fp(pn) or tmp_pn)
        return 0

    def arah_psnr_workspace(self):
        return 1184 * 8

    def arah_psnr(self, a, b, n, out, ws, ws_bytes, stream):
        if not (a and a.value) or not (b and b.value) or n <= 0:      # the C entry point: ARAH_EINVAL
            return 1
        assert ws_bytes >= 1184 * 8
        x = np.ctypeslib.as_array(C.cast(a, FP), (n,)); y = np.ctypeslib.as_array(C.cast(b, FP), (n,))
        d = (x - y).astype(np.float32)
        mse = np.float32(np.mean((d * d).astype(np.float64)))
        o = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), (2,))
        with np.errstate(divide='ignore'):
            o[0], o[1] = float(mse), -10.0 * np.log(float(mse)) / np.log(10.0)
        return 0

    def arah_ssim_workspace(self):
        return 4096

    def arah_ssim(self, a, b, mask, H, W, out, ws, ws_bytes, stream):
        X = np.ctypeslib.as_array(C.cast(a, FP), (H, W, 3)); Y = np.ctypeslib.as_array(C.cast(b, FP), (H, W, 3))
        m = np.ctypeslib.as_array(C.cast(mask, C.POINTER(C.c_uint8)), (H, W))
        o = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), (5,))
        o[1:] = io.bounding_rect(m)
        try:
            o[0] = io.ssim_metric(X, Y, m)
        except ValueError:
            o[0] = np.nan
        return 0

    def arah_rasterize_mesh_workspace(self, nv, H, W):
        return nv * 12 + 256 + H * W * 8

    def arah_rasterize_mesh(self, verts, nv, faces, nf, cam, H, W, p2f, zbuf, ws, ws_bytes, stream):
        assert ws_bytes >= nv * 12 + H * W * 8
        c = cam._obj
        c16 = np.array(list(c.R) + list(c.T) + [c.fx, c.fy, c.px, c.py], np.float32)
        ndc = np.zeros((max(nv, 1), 3), np.float32)
        self.h.host_project(_fp(verts), nv, c16.ctypes.data_as(FP), ndc.ctypes.data_as(FP))
        tmp = (C.c_float * (H * W))()
        self.h.host_rasterize(ndc.ctypes.data_as(FP), C.cast(faces, IP), nf, nv, H, W, C.cast(p2f, IP), _fp(zbuf) or tmp)
        return 0

    def arah_face_normal_image(self, verts, nv, faces, nf, p2f, H, W, sign, rot, background, image, stream):
        self.h.host_normal_image(_fp(verts), nv, C.cast(faces, IP), nf, C.cast(p2f, IP), H, W, C.c_float(sign), rot, C.c_float(background), _fp(image))
        return 0


@pytest.fixture()
def fi(hi, monkeypatch):  # noqa: F811
    from arah_release_b200 import images

    class HostFrameImages(images.FrameImages):
        def __init__(self):
            self.device = torch.device('cpu')
            self._ws = {}

        @property
        def _stream(self):
            return None

    shim = ShimLib(hi)
    monkeypatch.setattr(images._lib, 'lib', lambda: shim)
    return HostFrameImages()


@pytest.mark.parametrize('seed', [0, 2])
def test_assemble_and_psnr_plumbing(fi, seed):
    g = load_images_golden(seed)
    H, W = int(g['H']), int(g['W'])
    # deliberately awkward inputs: batch dimension, int64 pixel list, non-contiguous rows
    rgb = torch.from_numpy(g['rgb']).view(1, -1, 3)
    pts = torch.from_numpy(np.concatenate([g['points_cam'], g['points_cam']], 1))[:, :3]
    pp, pn = fi.assemble(rgb, pts, torch.from_numpy(g['pix'].astype(np.int64)), H, W)
    assert pp.shape == (H, W, 3) and pn.shape == (H, W, 3)
    o_pp, o_pn = io.frame_images(g['rgb'], g['points_cam'], g['pix'], H, W)
    assert np.array_equal(pp.numpy(), o_pp) and np.array_equal(pn.numpy(), o_pn)
    gt_img, none = fi.assemble(g['gt'], None, g['pix'], H, W, normals=False)
    assert none is None and np.array_equal(gt_img.numpy(), g['ref.rgb_gt'])
    mse, psnr = fi.psnr(rgb, torch.from_numpy(g['gt']))
    assert abs(psnr - float(g['ref.psnr'])) < 1e-5 and mse > 0


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_validation_tail_equals_reference_validation_step(fi, seed):
    """`FrameImages.validation_tail` on the very batch / model outputs the unmodified `validation_step` was run on."""
    g = load_images_golden(seed)
    H, W, P = int(g['H']), int(g['W']), len(g['pix'])
    outputs = {'rgb_values': torch.from_numpy(g['rgb']).view(1, P, 3), 'points_cam': torch.from_numpy(g['points_cam']).view(1, P, 3)}
    batch = {'inputs.img_height': torch.tensor([H]), 'inputs.img_width': torch.tensor([W]), 'inputs.image_mask': torch.from_numpy(g['mask']).view(1, -1),
             'inputs': torch.from_numpy(g['gt']).view(1, P, 3)}
    ev = fi.validation_tail(outputs, batch)
    assert abs(ev['psnr'] - float(g['ref.psnr'])) <= 1e-5
    assert ev['rgb_pred'].shape == (3, H, W)
    assert np.array_equal(ev['rgb_pred'].permute(1, 2, 0).numpy(), g['ref.rgb_pred'])
    assert np.array_equal(ev['rgb_gt'].permute(1, 2, 0).numpy(), g['ref.rgb_gt'])
    assert np.abs(ev['normal_pred'].permute(1, 2, 0).numpy() - g['ref.normal_pred']).max() <= 1.2e-7


def test_normal_maps_plumbing(fi):
    v, f = iso_mesh('torus', 20)
    H, W = 48, 64
    R, T, K = make_camera(H, W)
    T = T + np.array([0, 0, 2.6], np.float32)
    posed = (v + np.float32(0.02)).astype(np.float32)
    maps = fi.normal_maps(torch.from_numpy(v), torch.from_numpy(f.astype(np.int64)), torch.from_numpy(posed), torch.from_numpy(R).view(1, 3, 3),
                          torch.from_numpy(T).view(1, 3), torch.from_numpy(K).view(1, 3, 3), H, W)
    ref = io.normal_maps(v, f, posed, R, T, K, H, W)
    assert set(maps) == {'output_normal', 'normal_cano_front', 'normal_cano_back'}        # the keys models/__init__.py adds
    for k in ref:
        assert maps[k].shape == (1, H, W, 3)
        assert np.array_equal(maps[k][0].numpy(), ref[k])
    p2f, zb = fi.rasterize(v, f, __import__('arah_release_b200.images', fromlist=['x']).opencv_camera(R, T, K, H, W), H, W, zbuf=True)
    o_p2f, o_zb = io.rasterize(io.project(v, io.opencv_camera(R, T, K, H, W)), f, H, W)
    assert p2f.dtype == torch.int32 and np.array_equal(p2f.numpy(), o_p2f) and np.array_equal(zb.numpy(), o_zb)


@pytest.mark.parametrize('name', ['test_frame_images_edge_cases', 'test_psnr_large_and_reproducible', 'test_rasterize_degenerate_inputs',
                                  'test_frame_images_full_size_512', 'test_ssim_matches_oracle', 'test_validation_tail_on_golden_batch',
                                  'test_frame_images_match_reference_validation_step', 'test_rasterize_bit_exact_against_oracle'])
def test_gpu_test_bodies_hold_on_the_host_shim(fi, monkeypatch, name):
    """The assertions of tests/test_gpu_zz_images.py, executed here with the device set to 'cpu' and the shim in place of the
    library: guards the GPU tests' own expectations (edge cases, tolerances) before they ever reach a GPU box."""
    import test_gpu_zz_images as G
    monkeypatch.setattr(G, 'DEV', 'cpu')
    monkeypatch.setattr(G, '_fi', lambda: fi)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    fn = getattr(G, name)
    if name == 'test_frame_images_match_reference_validation_step':
        fn(1)
    elif name == 'test_rasterize_bit_exact_against_oracle':
        fn('hand', 48, 64); fn('sphere', 33, 130)
    else:
        fn()


def test_gpu_full_size_body_holds_on_the_host_shim(fi, monkeypatch):
    import test_gpu_zz_images as G
    monkeypatch.setattr(G, 'DEV', 'cpu')
    monkeypatch.setattr(G, '_fi', lambda: fi)
    G.test_normal_maps_full_size_512()
