"""TEST INFRASTRUCTURE — ctypes front-end of the C oracle (oracle/arah_oracle.c -> libarah_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libarah_oracle.so')
FP = C.POINTER(C.c_float)
U8 = C.POINTER(C.c_uint8)
I32 = C.POINTER(C.c_int32)


class OracleFrame(C.Structure):
    _fields_ = [('sdf_W', FP * 7), ('sdf_b', FP * 7), ('sdf_freq', FP), ('sdf_phase', FP),
                ('skin_W', FP * 5), ('skin_b', FP * 5), ('col_W', FP * 6), ('col_b', FP * 6),
                ('latent', FP), ('latent_dim', C.c_int32), ('beta', C.c_float),
                ('bone_T', FP), ('smpl_verts', FP), ('smpl_w', FP), ('n_verts', C.c_int32),
                ('trans', C.c_float * 3), ('cmin', C.c_float), ('cmax', C.c_float), ('center', C.c_float * 3),
                ('cam_loc', C.c_float * 3), ('pose', C.c_float * 16),
                ('n_steps', C.c_int32), ('near_samples', C.c_int32), ('far_samples', C.c_int32),
                ('cano_view_dirs', C.c_int32), ('render_last_pt', C.c_int32)]


class OracleOut(C.Structure):
    _fields_ = [('points_hat_norm', FP), ('trace_mask', U8), ('dists', FP), ('sampled_pts', FP),
                ('sampled_dists', FP), ('sampled_T', FP), ('sampled_conv', U8), ('rgb', FP), ('vol_mask', U8),
                ('points_cam', FP), ('weights_sum', FP), ('n_trace_evals', I32), ('n_iso_evals', I32),
                ('n_corr_evals', I32), ('n_shaded', I32)]


def build(force: bool = False):
    srcs = [os.path.join(_HERE, f) for f in ('arah_oracle.c', 'mc_oracle.c')]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'libarah_oracle.so'])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()                      # (re)compile when the .so is missing or older than its sources
        _lib = C.CDLL(_SO)
        _lib.arah_oracle_render.argtypes = [C.POINTER(OracleFrame), FP, FP, C.c_int, C.POINTER(OracleOut), C.c_int]
        _lib.arah_oracle_render_train.argtypes = [C.POINTER(OracleFrame), FP, FP, C.c_int, C.POINTER(OracleOut), C.c_int, FP, FP, FP]
        _lib.arah_oracle_sdf.argtypes = [C.POINTER(OracleFrame), FP, C.c_int, FP, FP, FP]
        _lib.arah_oracle_skin.argtypes = [C.POINTER(OracleFrame), FP, C.c_int, FP, FP, FP]
        _lib.arah_oracle_color.argtypes = [C.POINTER(OracleFrame), FP, FP, FP, FP, C.c_int, FP]
        _lib.arah_oracle_knn.argtypes = [C.POINTER(OracleFrame), FP, C.c_int, I32, FP, FP]
        _lib.arah_oracle_mc.argtypes = [FP, C.c_int, C.c_float, C.c_float, FP, FP, C.c_int, I32, C.c_int, I32]
        _lib.arah_oracle_mc_table.argtypes = [C.c_void_p, C.c_void_p]
    return _lib


def _fp(a):
    return a.ctypes.data_as(FP)


def _fold(layer):
    """weight_norm fold in float32 (w = v * (g / ||v||)), same arithmetic as arah_release_b200.synthetic."""
    from arah_release_b200.synthetic import fold_weight_norm
    return fold_weight_norm(layer)


def make_frame(frame):
    """syn.Frame -> (OracleFrame, keepalive list)."""
    keep = []
    def c(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        return _fp(a)
    of = OracleFrame()
    for i in range(7):
        of.sdf_W[i] = c(frame.sdf['W'][i])
        of.sdf_b[i] = c(frame.sdf['b'][i])
    of.sdf_freq = c(frame.sdf['freq'])
    of.sdf_phase = c(frame.sdf['phase'])
    for i in range(5):
        w, b = _fold(frame.skin[i])
        of.skin_W[i] = c(w)
        of.skin_b[i] = c(b)
    from arah_release_b200.synthetic import expand_color_weight
    for i in range(6):
        w, b = _fold(frame.color[i])
        if i in (0, 3):                      # 'no_view_dir' / 'no_normal' nets (decoder.py:101-106): exact zero columns
            w = expand_color_weight(w, getattr(frame, 'color_mode', 'idr'))
        of.col_W[i] = c(w)
        of.col_b[i] = c(b)
    of.latent = c(frame.latent)
    of.latent_dim = int(frame.latent.shape[0])
    of.beta = float(frame.beta)
    of.bone_T = c(frame.bone_transforms.reshape(24, 16))
    of.smpl_verts = c(frame.smpl_verts)
    of.smpl_w = c(frame.smpl_weights)
    of.n_verts = int(frame.smpl_verts.shape[0])
    of.trans[:] = [float(v) for v in frame.trans]
    of.cmin = float(frame.coord_min)
    of.cmax = float(frame.coord_max)
    of.center[:] = [float(v) for v in frame.center]
    of.cam_loc[:] = [float(v) for v in frame.cam_loc]
    of.pose[:] = [float(v) for v in frame.pose.reshape(-1)]
    of.n_steps = frame.n_steps
    of.near_samples = frame.near_samples
    of.far_samples = frame.far_samples
    of.cano_view_dirs = int(frame.cano_view_dirs)
    of.render_last_pt = int(bool(getattr(frame, 'render_last_pt', False)))
    return of, keep


def render(frame, ray_dirs=None, near_far=None, threads: int = 0, stages: bool = True, train_noise=None):
    """Full hot path on CPU.  Returns a dict with the same keys as oracle.ref_harness.run_reference plus counters.
    ``train_noise`` = (u_all [P,S], u_near [P,near+1], u_far [P,far]): training-mode tracing (eval_mode=False) with the
    reference's three torch.rand draws supplied by the caller."""
    of, keep = make_frame(frame)
    rd = np.ascontiguousarray(frame.ray_dirs if ray_dirs is None else ray_dirs, dtype=np.float32)
    nf = np.ascontiguousarray(frame.near_far if near_far is None else near_far, dtype=np.float32)
    P, S = rd.shape[0], frame.n_steps
    o = {
        'trace.points_hat_norm': np.zeros((P, 3), np.float32), 'trace.network_body_mask': np.zeros(P, np.uint8),
        'trace.dists': np.zeros(P, np.float32), 'rgb_values': np.zeros((P, 3), np.float32),
        'network_body_mask': np.zeros(P, np.uint8), 'points_cam': np.zeros((P, 3), np.float32),
        'weights_sum': np.zeros(P, np.float32), 'n_trace_evals': np.zeros(P, np.int32),
        'n_iso_evals': np.zeros(P, np.int32), 'n_corr_evals': np.zeros(P, np.int32), 'n_shaded': np.zeros(P, np.int32),
    }
    if stages:
        o.update({'trace.sampled_pts': np.zeros((P, S, 3), np.float32), 'trace.sampled_dists': np.zeros((P, S), np.float32),
                  'trace.sampled_transforms': np.zeros((P, S, 4, 4), np.float32),
                  'trace.sampler_converge_mask': np.zeros((P, S), np.uint8)})
    oo = OracleOut()
    oo.points_hat_norm = _fp(o['trace.points_hat_norm'])
    oo.trace_mask = o['trace.network_body_mask'].ctypes.data_as(U8)
    oo.dists = _fp(o['trace.dists'])
    if stages:
        oo.sampled_pts = _fp(o['trace.sampled_pts'])
        oo.sampled_dists = _fp(o['trace.sampled_dists'])
        oo.sampled_T = _fp(o['trace.sampled_transforms'])
        oo.sampled_conv = o['trace.sampler_converge_mask'].ctypes.data_as(U8)
    oo.rgb = _fp(o['rgb_values'])
    oo.vol_mask = o['network_body_mask'].ctypes.data_as(U8)
    oo.points_cam = _fp(o['points_cam'])
    oo.weights_sum = _fp(o['weights_sum'])
    for k in ('n_trace_evals', 'n_iso_evals', 'n_corr_evals', 'n_shaded'):
        setattr(oo, k, o[k].ctypes.data_as(I32))
    if train_noise is not None:
        tn = [np.ascontiguousarray(a, np.float32).reshape(P, -1) for a in train_noise]
        assert tn[0].shape[1] == S and tn[1].shape[1] == frame.near_samples + 1 and tn[2].shape[1] == frame.far_samples
        rc = lib().arah_oracle_render_train(C.byref(of), _fp(rd), _fp(nf), P, C.byref(oo), int(threads), _fp(tn[0]), _fp(tn[1]), _fp(tn[2]))
    else:
        rc = lib().arah_oracle_render(C.byref(of), _fp(rd), _fp(nf), P, C.byref(oo), int(threads))
    if rc != 0:
        raise RuntimeError(f'arah_oracle_render failed ({rc})')
    for k in ('trace.network_body_mask', 'network_body_mask', 'trace.sampler_converge_mask'):
        if k in o:
            o[k] = o[k].astype(bool)
    return o


def train_noise(frame, seed, P=None):
    """The three torch.rand draws of ray_sampler in training mode (ray_tracing.py:305), in the reference's order and shapes,
    from torch's CPU generator (the reference draws on the CPU even for CUDA runs: ``torch.rand(shape).to(upper)``)."""
    import torch
    P = frame.P if P is None else P
    torch.manual_seed(seed)
    u_all = torch.rand(1, P, frame.n_steps)
    u_near = torch.rand(1, P, frame.near_samples + 1)
    u_far = torch.rand(1, P, frame.far_samples)
    return u_all[0].numpy(), u_near[0].numpy(), u_far[0].numpy()


def sdf(frame, xn, grad=True, feat=False):
    of, keep = make_frame(frame)
    xn = np.ascontiguousarray(xn, np.float32)
    n = xn.shape[0]
    s = np.zeros(n, np.float32)
    g = np.zeros((n, 3), np.float32) if grad else None
    ft = np.zeros((n, 256), np.float32) if feat else None
    lib().arah_oracle_sdf(C.byref(of), _fp(xn), n, _fp(s), _fp(g) if grad else None, _fp(ft) if feat else None)
    return s, g, ft


def skin(frame, x_hat, jac=True):
    of, keep = make_frame(frame)
    x = np.ascontiguousarray(x_hat, np.float32)
    n = x.shape[0]
    w = np.zeros((n, 24), np.float32)
    xb = np.zeros((n, 3), np.float32)
    J = np.zeros((n, 3, 3), np.float32) if jac else None
    lib().arah_oracle_skin(C.byref(of), _fp(x), n, _fp(w), _fp(xb), _fp(J) if jac else None)
    return w, xb, J


def color(frame, xn, normal, view, feat):
    of, keep = make_frame(frame)
    a = [np.ascontiguousarray(v, np.float32) for v in (xn, normal, view, feat)]
    n = a[0].shape[0]
    rgb = np.zeros((n, 3), np.float32)
    lib().arah_oracle_color(C.byref(of), _fp(a[0]), _fp(a[1]), _fp(a[2]), _fp(a[3]), n, _fp(rgb))
    return rgb


def knn(frame, x):
    of, keep = make_frame(frame)
    x = np.ascontiguousarray(x, np.float32)
    n = x.shape[0]
    idx = np.zeros(n, np.int32)
    xh = np.zeros((n, 3), np.float32)
    T = np.zeros((n, 4, 4), np.float32)
    lib().arah_oracle_knn(C.byref(of), _fp(x), n, idx.ctypes.data_as(I32), _fp(xh), _fp(T))
    return idx, xh, T


def grid_points(N):
    """utils/sdf_meshing.py:20-38 — the N^3 lattice over [-1,1]^3 in the reference's order and fp32 arithmetic."""
    voxel = np.float32(2.0 / (N - 1))
    ax = (np.arange(N, dtype=np.float32) * voxel) + np.float32(-1.0)
    g = np.stack(np.meshgrid(ax, ax, ax, indexing='ij'), -1).reshape(-1, 3)
    return np.ascontiguousarray(g, np.float32), float(voxel)


def sdf_grid(frame, N):
    """sdf_meshing.py:40-58: raw SDF network output on the lattice, [N, N, N]."""
    pts, _ = grid_points(N)
    s, _, _ = sdf(frame, pts, grad=False)
    return s.reshape(N, N, N)


def marching_cubes(vol, level=0.0, voxel=None, origin=(-1.0, -1.0, -1.0)):
    """mc_oracle.c: vertices [nv,3] (origin + lattice position * voxel) and faces [nf,3] of the iso-surface."""
    vol = np.ascontiguousarray(vol, np.float32)
    N = vol.shape[0]
    assert vol.shape == (N, N, N)
    voxel = np.float32(2.0 / (N - 1)) if voxel is None else np.float32(voxel)
    org = np.asarray(origin, np.float32)
    counts = np.zeros(2, np.int32)
    dummy_v, dummy_f = np.zeros((1, 3), np.float32), np.zeros((1, 3), np.int32)
    rc = lib().arah_oracle_mc(_fp(vol), N, float(level), float(voxel), _fp(org), _fp(dummy_v), 0, dummy_f.ctypes.data_as(I32), 0, counts.ctypes.data_as(I32))
    assert rc == 0
    nv, nf = int(counts[0]), int(counts[1])
    v, f = np.zeros((max(nv, 1), 3), np.float32), np.zeros((max(nf, 1), 3), np.int32)
    rc = lib().arah_oracle_mc(_fp(vol), N, float(level), float(voxel), _fp(org), _fp(v), nv, f.ctypes.data_as(I32), nf, counts.ctypes.data_as(I32))
    assert rc == 0
    return v[:nv], f[:nf]


def mc_table():
    tri, ntri = np.zeros((256, 16), np.int8), np.zeros(256, np.uint8)
    rc = lib().arah_oracle_mc_table(tri.ctypes.data, ntri.ctypes.data)
    assert rc == 0
    return tri, ntri


def num_threads():
    return int(lib().arah_oracle_num_threads())
