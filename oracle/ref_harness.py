"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference hot path (/root/reference, Python/PyTorch) on CPU.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``oracle/gen_golden.py`` to produce the fixtures under ``tests/golden/`` that pin the C oracle
(``oracle/arah_oracle.c``) and, through it, the CUDA path.  Nothing in the product imports this.

How the reference is made importable: ``sys.modules`` stubs for packages that are absent
here and that never touch the hot path's arithmetic (pytorch_lightning, kornia, lpips, imageio, skimage,
plyfile, trimesh, igl, wandb ...) plus an exact CPU 1-NN for ``pytorch3d.ops.knn_points``
(/root/reference/im2mesh/metaavatar_render/renderer/ray_tracing.py:386,407 — brute force in float32 with the
same ``sum((x-v)^2)`` form the CUDA kernel and the C oracle use; a cKDTree variant is kept for timing runs).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

REF_ROOT = os.environ.get('ARAH_REFERENCE_ROOT', '/root/reference')

_STUBBED = ('pytorch3d', 'pytorch_lightning', 'kornia', 'lpips', 'imageio', 'skimage', 'plyfile', 'trimesh', 'igl',
            'wandb', 'torchmetrics', 'cv2_stub_never')


class _Anything:
    """Attribute sink: any attribute access / call returns another sink (never reached by hot-path arithmetic)."""
    def __init__(self, name='stub'):
        self.__name__ = name
    def __getattr__(self, k):
        if k.startswith('__') and k.endswith('__'):
            raise AttributeError(k)
        return _Anything(k)
    def __call__(self, *a, **k):
        return _Anything('call')
    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__') and k.endswith('__'):
            raise AttributeError(k)
        return _Anything(k)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in _STUBBED:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None
    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m
    def exec_module(self, module):
        pass


_KNN = namedtuple('KNN', ['dists', 'idx', 'knn'])
KNN_MODE = 'brute'          # 'brute' (bit-faithful fp32) | 'kdtree' (fast, float64 tree; timing only)


def _knn_points(p1, p2, K=1, **kw):
    """Exact 1-NN, role of pytorch3d.ops.knn_points (squared distances, int64 idx)."""
    assert K == 1 and p1.shape[0] == 1 and p2.shape[0] == 1
    a = p1[0].detach().float()
    b = p2[0].detach().float()
    if KNN_MODE == 'kdtree':
        from scipy.spatial import cKDTree
        d, i = cKDTree(b.numpy()).query(a.numpy(), k=1, workers=-1)
        idx = torch.from_numpy(i.astype(np.int64))
        d2 = torch.from_numpy((d * d).astype(np.float32))
    else:
        idx = torch.empty(a.shape[0], dtype=torch.int64)
        d2 = torch.empty(a.shape[0], dtype=torch.float32)
        for s in range(0, a.shape[0], 4096):
            diff = a[s:s + 4096, None, :] - b[None, :, :]
            dd = diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1] + diff[..., 2] * diff[..., 2]
            m, i = dd.min(dim=1)
            idx[s:s + 4096] = i
            d2[s:s + 4096] = m
    return _KNN(dists=d2.view(1, -1, 1), idx=idx.view(1, -1, 1), knn=None)


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f'reference tree not found at {REF_ROOT}; the harness only runs in the build container')
    sys.meta_path.insert(0, _StubFinder())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import pytorch3d  # noqa: F401  (stub)
    import pytorch3d.ops as ops
    ops.knn_points = _knn_points
    pytorch3d.ops = ops
    # im2mesh.utils.libmesh is a Cython extension that is not built; off the hot path
    lm = _StubModule('im2mesh.utils.libmesh')
    lm.__path__ = []
    sys.modules['im2mesh.utils.libmesh'] = lm
    _installed = True


# --------------------------------------------------------------------------------------
# build reference modules from a synthetic Frame
# --------------------------------------------------------------------------------------

def build_reference_modules(frame):
    """Instantiate the reference's own leaf networks + IDHRNetwork and load the frame's weights into them."""
    install()
    import im2mesh.metaavatar_render  # noqa: F401  (import order matters: the reference has an import cycle)
    from im2mesh import hyperlayers
    from im2mesh.metaavatar.models.siren_modules import Sine
    from im2mesh.metaavatar.models.decoder import Deformer
    from im2mesh.metaavatar_render.models.decoder import RenderingNetwork, SingleVarianceNetwork
    from im2mesh.metaavatar_render.models.skinning_model import SkinningModel
    from im2mesh.metaavatar_render.renderer.ray_tracing import BodyRayTracing
    from im2mesh.metaavatar_render.renderer.implicit_differentiable_renderer import IDHRNetwork

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()
    # SDF network exactly as HyperFCFiLM.forward assembles it (hyperlayers.py:270-285):
    layers = []
    for i in range(6):
        film = hyperlayers.BatchLinearFiLM(weights=t(frame.sdf['W'][i]).unsqueeze(0),
                                           biases=t(frame.sdf['b'][i]).view(1, 1, -1),
                                           freq=t(frame.sdf['freq'][i]).view(1, -1),
                                           phase_shift=t(frame.sdf['phase'][i]).view(1, -1))
        layers.append(nn.Sequential(film, Sine()))
    layers.append(hyperlayers.BatchLinear(weights=t(frame.sdf['W'][6]).unsqueeze(0),
                                          biases=t(frame.sdf['b'][6]).view(1, 1, -1)))
    sdf_network = nn.Sequential(*layers)

    # skinning net: configs/arah-zju/ZJUMOCAP-377_4gpus.yaml:40
    deformer = Deformer(d_in=3, d_out=25, d_hidden=128, n_layers=4, skip_in=[], cond_in=[], multires=0, bias=1.0,
                        geometric_init=False, weight_norm=True)
    with torch.no_grad():
        for i, L in enumerate(frame.skin):
            lin = getattr(deformer, f'lin{i}')
            lin.weight_v.copy_(t(L['v']))
            lin.weight_g.copy_(t(L['g']))
            lin.bias.copy_(t(L['b']))
    skinning_model = SkinningModel(skinning_decoder_fwd=deformer)

    # colour net: yaml:39 + config.py:96-133 (pose_encoder 'latent' -> d_feature 256+128)
    cmode = getattr(frame, 'color_mode', 'idr')         # mono configs: configs/arah-zju/ZJUMOCAP-39x-mono_4gpus.yaml:36
    rend = RenderingNetwork(d_feature=256 + 128, mode=cmode, d_in=6 if cmode != 'idr' else 9, d_out=3, d_hidden=256, n_layers=5,
                            weight_norm=True, multires=0, multires_view=0 if cmode == 'no_view_dir' else 4, skips=[3],
                            squeeze_out=True, pose_encoder='latent')
    with torch.no_grad():
        for i, L in enumerate(frame.color):
            lin = getattr(rend, f'lin{i}')
            lin.weight_v.copy_(t(L['v']))
            lin.weight_g.copy_(t(L['g']))
            lin.bias.copy_(t(L['b']))
    dev = SingleVarianceNetwork(float(frame.beta))
    tracer = BodyRayTracing(root_finding_threshold=1e-5, n_steps=frame.n_steps,
                            near_surface_vol_samples=frame.near_samples, far_surface_vol_samples=frame.far_samples,
                            sample_bg_pts=0, low_vram=False)
    idhr = IDHRNetwork(dev, rend, skinning_model, tracer, cano_view_dirs=frame.cano_view_dirs,
                       train_skinning_net=False, render_last_pt=False, low_vram=False)
    idhr.eval()
    return idhr, sdf_network


def reference_inputs(frame, sdf_network):
    """The dict IDHRNetwork.forward reads (implicit_differentiable_renderer.py:52-71), batch = 1."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float()
    P = frame.P
    return {
        'ray_dirs': t(frame.ray_dirs).view(1, P, 3),
        'cam_loc': t(frame.cam_loc).view(1, 3),
        'pose': t(frame.pose).view(1, 4, 4),
        'body_mask': torch.ones(1, P, dtype=torch.bool),
        'body_bounds_intersections': t(frame.near_far).view(1, P, 2),
        'loc': torch.zeros(1, 1, 3),
        'sc_factor': torch.ones(1, 1, 1),
        'smpl_verts': t(frame.smpl_verts).view(1, -1, 3),
        'skinning_weights': t(frame.smpl_weights).view(1, -1, 24),
        'vol_feat': torch.empty(1, 0),
        'bone_transforms': t(frame.bone_transforms).view(1, 24, 4, 4),
        'trans': t(frame.trans).view(1, 1, 3),
        'coord_min': t(np.array([frame.coord_min])).view(1, 1, 1),
        'coord_max': t(np.array([frame.coord_max])).view(1, 1, 1),
        'center': t(frame.center).view(1, 1, 3),
        'minimal_shape': t(frame.minimal_shape).view(1, -1, 3),
        'sdf_network': sdf_network,
        'pose_cond': {'latent_code': t(frame.latent).view(1, 128)},
    }


class Counters:
    """Iteration counters gathered by wrapping the reference's broyden() (utils/broyden.py:4)."""
    def __init__(self):
        self.calls = []

    def wrap(self):
        import im2mesh.utils.broyden as bmod
        import im2mesh.utils.root_finding_utils as rfu
        orig = bmod.broyden
        me = self

        def counted(g, x_init, T_init, J_inv_init, *a, **k):
            n_eval = [0]
            def g2(x, mask=None):
                n_eval[0] += int(mask.sum())
                return g(x, mask=mask)
            out = orig(g2, x_init, T_init, J_inv_init, *a, **k)
            me.calls.append({'n': int(x_init.shape[0]), 'dim': int(x_init.shape[1]), 'g_evals': n_eval[0],
                             'converged': int(out['valid_ids'].sum())})
            return out
        rfu.broyden = counted
        return lambda: setattr(rfu, 'broyden', orig)


def run_reference(frame, *, stages=True, threads=None, counters=None):
    """Run the reference eval forward on CPU.  Returns dict of numpy arrays."""
    install()
    if threads:
        torch.set_num_threads(threads)
    idhr, sdf_network = build_reference_modules(frame)
    inputs = reference_inputs(frame, sdf_network)
    out = {}
    undo = counters.wrap() if counters is not None else None
    try:
        if stages:
            with torch.no_grad():
                tr = idhr.ray_tracer(sdf_network, idhr.skinning_model,
                                     cam_loc=inputs['cam_loc'], ray_directions=inputs['ray_dirs'],
                                     body_bounds_intersections=inputs['body_bounds_intersections'],
                                     loc=inputs['loc'], sc_factor=inputs['sc_factor'],
                                     smpl_verts=inputs['smpl_verts'], smpl_verts_cano=inputs['minimal_shape'],
                                     skinning_weights=inputs['skinning_weights'], vol_feat=inputs['vol_feat'],
                                     bone_transforms=inputs['bone_transforms'], trans=inputs['trans'],
                                     coord_min=inputs['coord_min'], coord_max=inputs['coord_max'],
                                     center=inputs['center'], eval_mode=True)
            names = ['points_hat_norm', 'network_body_mask', 'dists', 'sampled_pts', 'sampled_dists',
                     'sampled_transforms', 'sampler_converge_mask']
            for n, v in zip(names, tr):
                out['trace.' + n] = v[0].numpy().copy()
        res = idhr(inputs)
        out['rgb_values'] = res['rgb_values'][0].detach().numpy().copy()
        out['network_body_mask'] = res['network_body_mask'][0].numpy().copy()
        out['points_cam'] = res['points_cam'][0].detach().numpy().copy()
    finally:
        if undo:
            undo()
    return out


# --------------------------------------------------------------------------------------
# training mode (BASELINE configs[2])
# --------------------------------------------------------------------------------------

LOSS_WEIGHTS = dict(rgb_weight=1.0, perceptual_weight=0.0, eikonal_weight=0.1, mask_weight=1.0, off_surface_weight=0.01,
                    inside_weight=0.01, params_weight=0.0, skinning_weight=10.0)


def run_reference_train(frame, aux, *, seed=0, train_skinning_net=True, threads=None, loss_weights=None):
    """One training forward + backward of the UNMODIFIED reference IDHRNetwork (self.training == True) on CPU.

    ``aux`` = arah_release_b200.synthetic.train_aux_points(frame).  torch.manual_seed(seed) is set right before the forward,
    so the reference's own torch.rand calls (three in ray_sampler, ray_tracing.py:305; one for the eikonal points,
    implicit_differentiable_renderer.py:126) are reproducible: the host mirror draws the same numbers in the same order.
    Returns numpy dict: tracer 7-tuple (train mode), model outputs, loss terms, gradients of every parameter tensor.
    """
    install()
    if threads:
        torch.set_num_threads(threads)
    idhr, sdf_network = build_reference_modules(frame)
    idhr.train_skinning_net = bool(train_skinning_net)
    idhr.train()
    from im2mesh.metaavatar_render.renderer.loss import IDHRLoss
    # leaves: hypernetwork outputs (SDF weights), latent code
    sdf_leaves = {}
    for l in range(6):
        film = sdf_network[l][0]
        for n in ('weights', 'biases', 'freq', 'phase_shift'):
            v = getattr(film, n).clone().requires_grad_(True)
            setattr(film, n, v)
            sdf_leaves[f'sdf.{l}.{n}'] = v
    for n in ('weights', 'biases'):
        v = getattr(sdf_network[6], n).clone().requires_grad_(True)
        setattr(sdf_network[6], n, v)
        sdf_leaves[f'sdf.6.{n}'] = v
    inputs = reference_inputs(frame, sdf_network)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    latent = inputs['pose_cond']['latent_code'].clone().requires_grad_(True)
    inputs['pose_cond']['latent_code'] = latent
    inputs['body_mask'] = t(aux['body_mask']).view(1, -1)
    inputs['points_uniform'] = t(aux['points_uniform']).float().view(1, -1, 3)
    inputs['points_skinning'] = t(aux['points_skinning']).float().view(1, -1, 3)
    inputs['points_inside'] = t(aux['points_inside']).float().view(1, -1, 3)
    out = {}
    torch.manual_seed(seed)
    state = torch.get_rng_state()
    with torch.no_grad():
        tr = idhr.ray_tracer(sdf_network, idhr.skinning_model, cam_loc=inputs['cam_loc'], ray_directions=inputs['ray_dirs'],
                             body_bounds_intersections=inputs['body_bounds_intersections'], loc=inputs['loc'],
                             sc_factor=inputs['sc_factor'], smpl_verts=inputs['smpl_verts'],
                             smpl_verts_cano=inputs['minimal_shape'], skinning_weights=inputs['skinning_weights'],
                             vol_feat=inputs['vol_feat'], bone_transforms=inputs['bone_transforms'], trans=inputs['trans'],
                             coord_min=inputs['coord_min'], coord_max=inputs['coord_max'], center=inputs['center'],
                             eval_mode=False)
    names = ['points_hat_norm', 'network_body_mask', 'dists', 'sampled_pts', 'sampled_dists', 'sampled_transforms',
             'sampler_converge_mask']
    for n, v in zip(names, tr):
        out['trace.' + n] = v[0].numpy().copy()
    torch.set_rng_state(state)           # replay: the full forward draws the same jitter, then the eikonal points
    res = idhr(inputs)
    for k in ('rgb_values', 'sdf_output', 'off_surface_sdf', 'grad_theta', 'pred_weights', 'inside_sdf'):
        out['out.' + k] = res[k].detach().numpy().copy()
    out['out.network_body_mask'] = res['network_body_mask'][0].numpy().copy()
    lw = dict(LOSS_WEIGHTS)
    lw.update(loss_weights or {})
    crit = IDHRLoss(rgb_loss_type='l1', **lw)
    gt = {'rgb': t(aux['rgb_gt']).float().view(1, -1, 3), 'sampled_weights': t(aux['sampled_weights']).float().view(1, -1, 24)}
    res['sdf_params'] = None
    losses = crit(res, gt)
    for k, v in losses.items():
        out['loss.' + k] = np.asarray(float(v.reshape(-1)[0]) if torch.is_tensor(v) else float(v), np.float64)
    losses['loss'].backward()
    for k, v in sdf_leaves.items():
        out['grad.' + k] = v.grad.numpy().copy()
    out['grad.latent'] = latent.grad.numpy().copy()
    for k, p in idhr.named_parameters():
        out['grad.' + k] = (p.grad.numpy().copy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
    return out
