"""CPU: the per-element arithmetic of csrc/arah_image.cu (csrc/arah_image_core.h, compiled for the host by
tests/native/host_image.cpp — test infrastructure, never loaded by the product) against the numpy oracle and the reference's
golden validation images.  The CUDA kernels call the same functions with the same index expressions, so this pins their
arithmetic in the build container, which has no GPU; the GPU tests (test_gpu_zz_images.py) then check the kernels themselves.

Bars: integer outputs (pix_to_face) bit-exact against the oracle; floating point bit-exact against the oracle as well (both
round every fp32 operation once, in the same order) and 1.2e-7 against the reference's images (see test_images_oracle.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers_images import handmade_mesh, iso_mesh, load_images_golden, make_camera
from oracle import images_oracle as io

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, 'native', 'host_image.cpp')
CORE = os.path.join(ROOT, 'arah_release_b200', 'csrc', 'arah_image_core.h')
SO = os.path.join(HERE, 'native', 'libarah_image_host.so')
FP, IP = C.POINTER(C.c_float), C.POINTER(C.c_int32)


@pytest.fixture(scope='module')
def hi():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
        subprocess.check_call([cxx, '-O2', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC', '-o', SO, SRC])
    return C.CDLL(SO)


def _f(a):
    return a.ctypes.data_as(FP)


def _i(a):
    return a.ctypes.data_as(IP)


def _cam16(cam):
    return np.concatenate([cam['R'].reshape(9), cam['T'].reshape(3), [cam['fx'], cam['fy'], cam['px'], cam['py']]]).astype(np.float32)


def host_rasterize(hi, verts, faces, cam, H, W):
    v = np.ascontiguousarray(verts, np.float32); f = np.ascontiguousarray(faces, np.int32)
    ndc = np.zeros_like(v)
    c16 = _cam16(cam)
    hi.host_project(_f(v), v.shape[0], _f(c16), _f(ndc))
    p2f, zb = np.zeros((H, W), np.int32), np.zeros((H, W), np.float32)
    hi.host_rasterize(_f(ndc), _i(f), f.shape[0], v.shape[0], H, W, _i(p2f), _f(zb))
    return ndc, p2f, zb


def host_normal_image(hi, verts, faces, p2f, sign, rot, background):
    v = np.ascontiguousarray(verts, np.float32); f = np.ascontiguousarray(faces, np.int32)
    H, W = p2f.shape
    img = np.zeros((H, W, 3), np.float32)
    r = None if rot is None else np.ascontiguousarray(rot, np.float32).reshape(9)
    hi.host_normal_image(_f(v), v.shape[0], _i(f), f.shape[0], _i(np.ascontiguousarray(p2f)), H, W, C.c_float(sign), None if r is None else _f(r),
                         C.c_float(background), _f(img))
    return img


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_frame_images(hi, seed):
    g = load_images_golden(seed)
    H, W = int(g['H']), int(g['W'])
    P = len(g['pix'])
    pp, pn = np.zeros((H, W, 3), np.float32), np.zeros((H, W, 3), np.float32)
    hi.host_frame_images(_f(np.ascontiguousarray(g['rgb'])), _f(np.ascontiguousarray(g['points_cam'])), _i(np.ascontiguousarray(g['pix'])), P, H, W,
                         _f(pp), _f(pn))
    o_pp, o_pn = io.frame_images(g['rgb'], g['points_cam'], g['pix'], H, W)
    assert np.array_equal(pp, o_pp) and np.array_equal(pp, g['ref.rgb_pred'])
    assert np.array_equal(pn, o_pn)                                     # same rounding sequence as the oracle: bit-exact
    assert np.abs(pn - g['ref.normal_pred']).max() <= 1.2e-7            # the unmodified reference


@pytest.mark.parametrize('mesh,H,W', [('hand', 48, 64), ('hand', 64, 40), ('torus', 72, 72), ('two_spheres', 64, 96), ('sphere', 33, 130)])
def test_rasterize_bit_exact_against_oracle(hi, mesh, H, W):
    if mesh == 'hand':
        v, f = handmade_mesh()
        R, T, K = np.eye(3, dtype=np.float32), np.zeros(3, np.float32), make_camera(H, W)[2]
    else:
        v, f = iso_mesh(mesh, 20)
        R, T, K = make_camera(H, W, shift=(1.5, -2.25))
        T = T + np.array([0, 0, 2.6], np.float32)
    for cam in (io.opencv_camera(R, T, K, H, W), io.fov_camera(*io.look_at_view_transform(2.0, 0.0, 0.0)),
                io.fov_camera(*io.look_at_view_transform(2.0, 0.0, 180.0))):
        ndc, p2f, zb = host_rasterize(hi, v, f, cam, H, W)
        o_ndc = io.project(v, cam)
        assert np.array_equal(ndc, o_ndc, equal_nan=True)
        o_p2f, o_zb = io.rasterize(o_ndc, f, H, W)
        assert np.array_equal(p2f, o_p2f)
        assert np.array_equal(zb, o_zb)
        assert (p2f >= 0).any()


def test_pixel_range_is_conservative(hi):
    """A face whose bounding box ends exactly on pixel centres / straddles the image border loses no pixel to the index-range
    shortcut: compare with the oracle, which tests every pixel of the image."""
    H, W = 16, 24
    rng = np.random.default_rng(5)
    cam = io.fov_camera(*io.look_at_view_transform(2.0, 0.0, 0.0))
    xf = io.pix_to_ndc(W - 1 - np.arange(W), W, H); yf = io.pix_to_ndc(H - 1 - np.arange(H), H, W)
    verts, faces = [], []
    for k in range(300):
        # corners snapped to pixel centres in ndc (z = 1 plane in view space <-> world z = 1, x = -ndc / s)
        cx, cy = rng.choice(xf, 3) + rng.choice([0, 0, 1e-7, -1e-7, 0.3], 3), rng.choice(yf, 3) + rng.choice([0, 0, 1e-7, -1e-7, 2.0], 3)
        z = rng.uniform(0.5, 1.5, 3)
        for a in range(3):
            zv = 2.0 - z[a]
            verts.append([-(cx[a] * zv) / cam['fx'], (cy[a] * zv) / cam['fy'], z[a]])
        faces.append([3 * k, 3 * k + 1, 3 * k + 2])
    v, f = np.array(verts, np.float32), np.array(faces, np.int32)
    _, p2f, zb = host_rasterize(hi, v, f, cam, H, W)
    o_p2f, o_zb = io.rasterize(io.project(v, cam), f, H, W)
    assert np.array_equal(p2f, o_p2f) and np.array_equal(zb, o_zb)


def test_normal_maps_bit_exact_against_oracle(hi):
    v, f = iso_mesh('torus', 20)
    H = W = 64
    R, T, K = make_camera(H, W)
    T = T + np.array([0, 0, 2.6], np.float32)
    posed = (v @ np.array([[0.9, 0.1, 0], [-0.1, 0.9, 0.05], [0, -0.05, 1.0]], np.float32) + np.float32(0.03)).astype(np.float32)
    ref = io.normal_maps(v, f, posed, R, T, K, H, W)
    _, p2f, _ = host_rasterize(hi, posed, f, io.opencv_camera(R, T, K, H, W), H, W)
    assert np.array_equal(host_normal_image(hi, posed, f, p2f, -1.0, R, -1.0), ref['output_normal'])
    for name, az in (('normal_cano_front', 0.0), ('normal_cano_back', 180.0)):
        _, p2f, _ = host_rasterize(hi, v, f, io.fov_camera(*io.look_at_view_transform(2.0, 0.0, az)), H, W)
        assert np.array_equal(host_normal_image(hi, v, f, p2f, 1.0, None, 0.0), ref[name])
