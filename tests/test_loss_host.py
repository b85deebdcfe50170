"""CPU: the fused training loss without a GPU.

(a) the per-element arithmetic of csrc/arah_loss.cu (csrc/arah_loss_core.h compiled for the host by tests/native/host_loss.cpp —
    test infrastructure, never loaded by the product) against the terms and autograd gradients of the unmodified reference
    `IDHRLoss` (tests/golden/loss_s*.npz);
(b) the product's Python mirror (arah_release_b200/loss.py: marshalling, the 2048-ray cut, result keys / shapes, the autograd
    wrapper) driven on CPU tensors through a shim of the C ABI that forwards `arah_idhr_loss` to (a).
The CUDA kernels themselves are checked on the GPU box (tests/test_gpu_zx_loss.py).  Tolerances: tests/helpers_loss.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from helpers_loss import LOSS_SEEDS, TERMS, check_loss, load_loss_golden

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, 'native', 'host_loss.cpp')
DEPS = [SRC, os.path.join(ROOT, 'arah_release_b200', 'csrc', 'arah_loss_core.h'), os.path.join(ROOT, 'include', 'arah_b200.h')]
SO = os.path.join(HERE, 'native', 'libarah_loss_host.so')
RGB_TYPES = {'l1': 0, 'mse': 1, 'smoothed_l1': 2}


@pytest.fixture(scope='module')
def hl():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in DEPS):
        cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
        subprocess.check_call([cxx, '-O2', '-std=c++17', '-shared', '-fPIC', '-o', SO, SRC])
    return C.CDLL(SO)


def _vp(a):
    return None if a is None or a.size == 0 else C.c_void_p(a.ctypes.data)


def run_host(hl, cfg, cut, with_grads=True):
    from arah_release_b200 import _lib
    c, inp, g = _lib.ArahLossConfig(), _lib.ArahLossInputs(), _lib.ArahLossGrads()
    for k, v in cfg.items():
        if k.endswith('_weight'):
            setattr(c, k, float(v))
    c.rgb_loss_type = RGB_TYPES[cfg['rgb_loss_type']]
    keep = {k: np.ascontiguousarray(cut[k], np.float32) for k in ('rgb_values', 'rgb_gt', 'sdf_output', 'grad_theta', 'off_surface_sdf', 'inside_sdf',
                                                                   'pred_weights', 'sampled_weights')}
    keep.update({k: np.ascontiguousarray(cut[k]).astype(np.uint8) for k in ('network_body_mask', 'body_mask', 'off_surface_mask')})
    for k, a in keep.items():
        setattr(inp, k, _vp(a))
    params = [np.ascontiguousarray(p, np.float32) for p in cut['sdf_params']]
    inp.n_rays, inp.n_eikonal, inp.n_off, inp.n_inside = keep['body_mask'].size, keep['grad_theta'].shape[0], keep['off_surface_sdf'].size, keep['inside_sdf'].size
    inp.n_skin, inp.n_joints, inp.n_param_tensors = keep['pred_weights'].shape[0], keep['pred_weights'].shape[1], len(params)
    grads = {k: np.full_like(keep[k], np.nan) for k in ('rgb_values', 'sdf_output', 'grad_theta', 'off_surface_sdf', 'inside_sdf', 'pred_weights')}
    pg = [np.full_like(p, np.nan) for p in params]
    for k, a in grads.items():
        setattr(g, k, _vp(a))
    for i, (p, q) in enumerate(zip(params, pg)):
        inp.sdf_params[i], inp.sdf_params_count[i], g.sdf_params[i] = _vp(p), p.size, _vp(q)
    terms = np.zeros(9, np.float32)
    rc = hl.host_idhr_loss(C.byref(c), C.byref(inp), terms.ctypes.data_as(C.POINTER(C.c_float)), C.byref(g) if with_grads else None)
    assert rc == 0
    grads['sdf_params'] = pg
    return dict(zip(TERMS, terms.tolist())), grads


@pytest.mark.parametrize('seed', LOSS_SEEDS)
def test_host_core_matches_reference(hl, seed):
    cfg, cut, full, ref = load_loss_golden(seed)
    terms, grads = run_host(hl, cfg, cut)
    for k, a in grads.items():
        if k != 'sdf_params':
            assert not np.isnan(a).any(), k + ': gradient buffer not fully written'
    check_loss(terms, grads, ref, full)
    t2, _ = run_host(hl, cfg, cut, with_grads=False)
    assert t2 == terms


# ------------------------------------------------------------------------------------------------ (b) the Python mirror
class ShimLib:
    def __init__(self, h):
        self.h = h

    def arah_last_error(self):
        return b'invalid argument (shim)'

    def arah_idhr_loss_workspace(self):
        return 9472

    def arah_idhr_loss(self, cfg, inp, terms, grads, ws, ws_bytes, stream):
        assert ws_bytes >= 9472
        return -1 if self.h.host_idhr_loss(cfg, inp, C.cast(terms, C.POINTER(C.c_float)), grads) != 0 else 0


@pytest.fixture()
def criterion_factory(hl, monkeypatch):
    from arah_release_b200 import loss as L
    shim = ShimLib(hl)
    monkeypatch.setattr(L._lib, 'lib', lambda: shim)

    class HostIDHRLoss(L.IDHRLoss):            # lifts exactly the two things that need a GPU: the device check and the stream
        def _stream(self, dev):
            return None

        @staticmethod
        def _require_cuda(dev):
            pass

    def make(cfg):
        return HostIDHRLoss(rgb_loss_type=cfg['rgb_loss_type'], **{k: v for k, v in cfg.items() if k.endswith('_weight')})
    return make


@pytest.mark.parametrize('seed', LOSS_SEEDS)
def test_mirror_matches_reference_through_shim(criterion_factory, seed):
    cfg, cut, full, ref = load_loss_golden(seed)
    crit = criterion_factory(cfg)
    t = lambda a, rg=False: torch.from_numpy(np.ascontiguousarray(a)).unsqueeze(0).requires_grad_(rg)
    leaves = {k: t(full[k], True) for k in ('rgb_values', 'sdf_output', 'pred_weights')}
    leaves.update({k: torch.from_numpy(full[k]).requires_grad_(True) for k in ('grad_theta', 'off_surface_sdf', 'inside_sdf')})
    params = [t(p, True) for p in full['sdf_params']]
    mo = {'rgb_values': leaves['rgb_values'], 'sdf_output': leaves['sdf_output'], 'network_body_mask': t(full['network_body_mask']),
          'body_mask': t(full['body_mask']), 'off_surface_mask': t(full['off_surface_mask']), 'surface_normals': None, 'grad_theta': leaves['grad_theta'],
          'off_surface_sdf': leaves['off_surface_sdf'], 'inside_sdf': leaves['inside_sdf'], 'pred_weights': leaves['pred_weights'], 'sdf_params': params}
    out = crit(mo, {'rgb': t(full['rgb_gt']), 'sampled_weights': t(full['sampled_weights'])})
    assert tuple(out) == TERMS
    assert tuple(out['loss'].shape) == tuple(ref['loss_shape'])
    for k in TERMS[1:]:
        w = cfg[k.replace('_loss', '_weight').replace('sdf_params', 'params')]
        assert tuple(out[k].shape) == (() if w > 0 else (1,))
    out['loss'].sum().backward()
    terms = {k: float(out[k].detach().reshape(-1)[0]) for k in TERMS}
    grads = {k: (v.grad.numpy().reshape(np.asarray(full[k]).shape) if v.grad is not None else None) for k, v in leaves.items()}
    grads['sdf_params'] = [p.grad.numpy().reshape(-1) if p.grad is not None else np.zeros(p.numel(), np.float32) for p in params]
    check_loss(terms, grads, ref, full)
    # rows beyond the 2048-ray cut receive exactly zero
    if full['rgb_values'].shape[0] > 2048 and grads['rgb_values'] is not None:
        assert not grads['rgb_values'][2048:].any()


def test_mirror_rejects_row_count_mismatch(criterion_factory):
    """Every per-ray tensor must have body_mask's row count: the kernels index all of them up to n_rays (a mismatch
    that raises an indexing error in the reference must not become an out-of-bounds device access)."""
    from arah_release_b200 import _lib
    cfg, cut, full, ref = load_loss_golden(LOSS_SEEDS[0])
    crit = criterion_factory(cfg)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).unsqueeze(0)

    def outputs(**short):
        mo = {'rgb_values': t(full['rgb_values']), 'sdf_output': t(full['sdf_output']), 'network_body_mask': t(full['network_body_mask']),
              'body_mask': t(full['body_mask']), 'off_surface_mask': t(full['off_surface_mask']), 'surface_normals': None,
              'grad_theta': torch.from_numpy(full['grad_theta']), 'off_surface_sdf': torch.from_numpy(full['off_surface_sdf']),
              'inside_sdf': torch.from_numpy(full['inside_sdf']), 'pred_weights': t(full['pred_weights']), 'sdf_params': [t(p) for p in full['sdf_params']]}
        for k, n in short.items():
            mo[k] = mo[k][:, :n]
        return mo
    gt = {'rgb': t(full['rgb_gt']), 'sampled_weights': t(full['sampled_weights'])}
    n = full['body_mask'].shape[0]
    for key in ('rgb_values', 'network_body_mask', 'off_surface_mask'):
        with pytest.raises(_lib.ArahError):
            crit(outputs(**{key: n - 3}), gt)
    with pytest.raises(_lib.ArahError):
        crit(outputs(), {'rgb': t(full['rgb_gt'])[:, :n - 3], 'sampled_weights': gt['sampled_weights']})


@pytest.mark.parametrize('seed', LOSS_SEEDS)
def test_gpu_test_body_holds_on_the_host_shim(hl, monkeypatch, seed):
    """The assertions of tests/test_gpu_zx_loss.py::test_fused_loss_matches_reference, executed with the device set to 'cpu' and the
    shim in place of the library: guards the GPU test's own expectations before it reaches a GPU box."""
    import test_gpu_zx_loss as G
    from arah_release_b200 import loss as L
    monkeypatch.setattr(L._lib, 'lib', lambda: ShimLib(hl))
    monkeypatch.setattr(L.IDHRLoss, '_require_cuda', staticmethod(lambda dev: None))
    monkeypatch.setattr(L.IDHRLoss, '_stream', lambda self, dev: None)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)
    monkeypatch.setattr(G, 'DEV', 'cpu')
    G.test_fused_loss_matches_reference(seed)
