"""CPU: the hand-written training chain rule (arah_release_b200/csrc/arah_train.h) instantiated with a plain-loop host
backend (tests/native/host_train.cpp — test infrastructure, never loaded by the product) against one training step of the
UNMODIFIED reference (tests/golden/train_*.npz): outputs, loss and the gradient of every parameter tensor.

The CUDA library instantiates the same templates with the CUDA backend; tests/test_gpu_train.py repeats this on the B200."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from helpers_train import TRAIN_CASES, compare_grad, load_train_golden, ref_to_oracle_name

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'native', 'host_train.cpp')
SO = os.path.join(HERE, 'native', 'libarah_train_host.so')
HDR = os.path.join(HERE, '..', 'arah_release_b200', 'csrc', 'arah_train.h')
FP = C.POINTER(C.c_float)


class HostTrainParams(C.Structure):
    _fields_ = [('sdf_W', FP * 7), ('sdf_b', FP * 7), ('sdf_freq', FP), ('sdf_phase', FP), ('skin_W', FP * 5), ('skin_b', FP * 5),
                ('col_W', FP * 6), ('col_b', FP * 6), ('latent', FP), ('latent_dim', C.c_int32), ('bone_T', FP),
                ('cmin', C.c_float), ('cmax', C.c_float), ('center', C.c_float * 3)]


class HostTrainGrads(C.Structure):
    _fields_ = [('sdf_W', FP * 7), ('sdf_b', FP * 7), ('sdf_freq', FP), ('sdf_phase', FP), ('skin_W', FP * 5), ('skin_b', FP * 5),
                ('col_W', FP * 6), ('col_b', FP * 6), ('latent', FP), ('beta', FP)]


@pytest.fixture(scope='module')
def lib():
    deps = [SRC, HDR, os.path.join(os.path.dirname(HDR), 'arah_math.cuh')]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(['g++', '-O2', '-fopenmp', '-fPIC', '-shared', '-std=c++17', '-o', SO, SRC])
    L = C.CDLL(SO)
    L.host_train_shade_forward.restype = C.c_int
    return L


def _fp(a):
    return a.ctypes.data_as(FP)


def build_params(fr):
    """numpy copies in the C ABI's layout (weight-norm applied in float32 like torch._weight_norm)."""
    from arah_release_b200.synthetic import fold_weight_norm
    keep = {}
    p = HostTrainParams()
    def c(name, a):
        a = np.ascontiguousarray(a, np.float32)
        keep[name] = a
        return _fp(a)
    for i in range(7):
        p.sdf_W[i] = c(f'sdf_W{i}', fr.sdf['W'][i]); p.sdf_b[i] = c(f'sdf_b{i}', fr.sdf['b'][i])
    p.sdf_freq = c('sdf_freq', fr.sdf['freq']); p.sdf_phase = c('sdf_phase', fr.sdf['phase'])
    for i in range(5):
        w, b = fold_weight_norm(fr.skin[i]); p.skin_W[i] = c(f'skin_W{i}', w); p.skin_b[i] = c(f'skin_b{i}', b)
    for i in range(6):
        w, b = fold_weight_norm(fr.color[i]); p.col_W[i] = c(f'col_W{i}', w); p.col_b[i] = c(f'col_b{i}', b)
    p.latent = c('latent', fr.latent); p.latent_dim = int(fr.latent.shape[0])
    p.bone_T = c('bone_T', fr.bone_transforms.reshape(24, 16))
    p.cmin, p.cmax = float(fr.coord_min), float(fr.coord_max)
    p.center[:] = [float(v) for v in fr.center]
    return p, keep


def build_grads(keep):
    g = HostTrainGrads()
    gk = {}
    def z(name):
        a = np.zeros_like(keep[name]); gk[name] = a; return _fp(a)
    for i in range(7):
        g.sdf_W[i] = z(f'sdf_W{i}'); g.sdf_b[i] = z(f'sdf_b{i}')
    g.sdf_freq = z('sdf_freq'); g.sdf_phase = z('sdf_phase')
    for i in range(5):
        g.skin_W[i] = z(f'skin_W{i}'); g.skin_b[i] = z(f'skin_b{i}')
    for i in range(6):
        g.col_W[i] = z(f'col_W{i}'); g.col_b[i] = z(f'col_b{i}')
    g.latent = z('latent')
    gk['beta'] = np.zeros(1, np.float32); g.beta = _fp(gk['beta'])
    return g, gk


def to_reference_grads(fr, gk):
    """engine gradients (w.r.t. effective weights, |variance|) -> gradients of the reference's parameter tensors."""
    out = {}
    for l in range(6):
        out[f'sdf.{l}.weights'] = gk[f'sdf_W{l}']; out[f'sdf.{l}.biases'] = gk[f'sdf_b{l}']
        out[f'sdf.{l}.freq'] = gk['sdf_freq'][l]; out[f'sdf.{l}.phase_shift'] = gk['sdf_phase'][l]
    out['sdf.6.weights'] = gk['sdf_W6']; out['sdf.6.biases'] = gk['sdf_b6']
    for net, layers, key in (('skin', fr.skin, 'skin'), ('col', fr.color, 'col')):
        for i, L in enumerate(layers):
            v = torch.tensor(np.asarray(L['v'], np.float32), requires_grad=True)
            g = torch.tensor(np.asarray(L['g'], np.float32).reshape(-1, 1), requires_grad=True)
            torch._weight_norm(v, g, 0).backward(torch.from_numpy(gk[f'{key}_W{i}']))
            out[f'{net}.lin{i}.weight_v'] = v.grad.numpy(); out[f'{net}.lin{i}.weight_g'] = g.grad.numpy()
            out[f'{net}.lin{i}.bias'] = gk[f'{key}_b{i}']
    out['latent'] = gk['latent']
    out['variance'] = gk['beta'] * np.sign(float(fr.beta))          # d ||v|| / d v
    return out


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_host_engine_matches_reference_step(lib, name):
    from oracle import train_oracle as to
    fr, aux, ref, grads, meta = load_train_golden(name)
    P, S = fr.P, fr.n_steps
    p, keep = build_params(fr)
    g, gk = build_grads(keep)
    xn = np.ascontiguousarray(ref['trace.sampled_pts'], np.float32)
    T12 = np.ascontiguousarray(ref['trace.sampled_transforms'][..., :3, :], np.float32)
    z = np.ascontiguousarray(ref['trace.sampled_dists'], np.float32)
    conv = np.ascontiguousarray(ref['trace.sampler_converge_mask']).astype(np.uint8)
    view = np.ascontiguousarray(fr.ray_dirs, np.float32)
    rgb = np.zeros((P, 3), np.float32); ws = np.zeros(P, np.float32)
    M = lib.host_train_shade_forward(C.byref(p), P, S, int(fr.cano_view_dirs), 0, int(meta['train_skinning_net']), C.c_float(float(fr.beta)),
                                     _fp(xn), _fp(T12), _fp(z), conv.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(view), None, _fp(rgb), _fp(ws))
    assert M == int(conv.sum())
    assert np.abs(rgb - ref['out.rgb_values'][0]).max() <= 2e-5
    assert np.abs(ws - ref['out.sdf_output'][0]).max() <= 2e-5
    # auxiliary evaluations
    eik = to.eikonal_points(meta['seed'], fr)
    pts_all = np.ascontiguousarray(np.concatenate([eik, aux['points_uniform']], 0), np.float32)
    n_all = pts_all.shape[0]
    s_all = np.zeros(n_all, np.float32); g_all = np.zeros((n_all, 3), np.float32)
    assert lib.host_train_sdf_forward(C.byref(p), 0, _fp(pts_all), n_all, 1, _fp(s_all), _fp(g_all)) == 0
    pin = np.ascontiguousarray(aux['points_inside'], np.float32)
    s_in = np.zeros(pin.shape[0], np.float32)
    assert lib.host_train_sdf_forward(C.byref(p), 1, _fp(pin), pin.shape[0], 0, _fp(s_in), None) == 0
    psk = np.ascontiguousarray(aux['points_skinning'], np.float32)
    pw = np.zeros((psk.shape[0], 24), np.float32)
    assert lib.host_train_skin_forward(C.byref(p), _fp(psk), psk.shape[0], _fp(pw)) == 0
    assert np.abs(g_all[:1024] - ref['out.grad_theta']).max() <= 1e-4 * max(1.0, np.abs(ref['out.grad_theta']).max())
    assert np.abs(s_all[1024:] - ref['out.off_surface_sdf'].reshape(-1)).max() <= 1e-5
    assert np.abs(s_in - ref['out.inside_sdf'].reshape(-1)).max() <= 1e-5
    assert np.abs(pw - ref['out.pred_weights'][0]).max() <= 1e-5
    # loss (torch restatement of IDHRLoss) on leaf copies of the outputs -> output gradients for the engine's backward
    lw = dict(to.LOSS_WEIGHTS); lw.update(meta['loss_weights'])
    leaf = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).requires_grad_(True)
    o = {'rgb_values': leaf(rgb), 'sdf_output': leaf(ws), 'grad_theta': leaf(g_all[:1024]), 'off_surface_sdf': leaf(s_all[1024:].reshape(-1, 1)),
         'inside_sdf': leaf(s_in.reshape(-1, 1)), 'pred_weights': leaf(pw), 'vol_mask': torch.from_numpy(conv.any(-1))}
    terms = to.loss_terms(o, aux, lw)
    assert abs(float(terms['loss']) - float(ref['loss.loss'])) <= 1e-5 * max(1.0, abs(float(ref['loss.loss'])))
    terms['loss'].backward()
    gz = lambda t: np.ascontiguousarray((t.grad if t.grad is not None else torch.zeros_like(t)).numpy(), np.float32)
    g_rgb, g_ws = gz(o['rgb_values']), gz(o['sdf_output'])
    assert lib.host_train_shade_backward(C.byref(p), C.byref(g), _fp(g_rgb), _fp(g_ws)) == 0
    g_s_all = np.zeros(n_all, np.float32); g_s_all[1024:] = gz(o['off_surface_sdf']).reshape(-1)
    g_g_all = np.zeros((n_all, 3), np.float32); g_g_all[:1024] = gz(o['grad_theta'])
    assert lib.host_train_sdf_backward(C.byref(p), C.byref(g), 0, _fp(g_s_all), _fp(g_g_all)) == 0
    g_in = gz(o['inside_sdf']).reshape(-1).copy()
    assert lib.host_train_sdf_backward(C.byref(p), C.byref(g), 1, _fp(g_in), None) == 0
    g_pw = gz(o['pred_weights'])
    assert lib.host_train_skin_backward(C.byref(p), C.byref(g), _fp(g_pw)) == 0
    ours = to_reference_grads(fr, gk)
    worst, worst_k = 1.0, None
    for k, dig in grads.items():
        st = compare_grad(k, ours[ref_to_oracle_name(k)], dig)
        if st['cos'] < worst:
            worst, worst_k = st['cos'], k
    print(name, 'M', M, 'min gradient cosine vs reference', worst, worst_k)


def test_hierarchical_softmax_vjp(lib):
    from oracle import train_oracle as to
    rng = np.random.default_rng(0)
    for _ in range(20):
        x = torch.tensor(rng.normal(scale=3.0, size=(1, 25)).astype(np.float32), requires_grad=True)
        gp = rng.normal(size=24).astype(np.float32)
        (to.hierarchical_softmax(x)[0] * torch.from_numpy(gp)).sum().backward()
        gx = np.zeros(25, np.float32)
        lib.host_train_hsoftmax_vjp(_fp(np.ascontiguousarray(x.detach().numpy()[0])), _fp(gp), _fp(gx))
        assert np.abs(gx - x.grad.numpy()[0]).max() <= 1e-5 * max(1.0, np.abs(gx).max())
