"""TEST INFRASTRUCTURE — golden fixtures for the training loss (tests/golden/loss_s*.npz).

Runs the UNMODIFIED `IDHRLoss` (/root/reference/im2mesh/metaavatar_render/renderer/loss.py) on seeded synthetic model outputs
with torch autograd and stores the nine terms and d loss / d input for every differentiable input.

    python -m oracle.gen_golden_loss            (build container only; needs /root/reference)
"""
import json
import os

import numpy as np
import torch

from oracle import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = [
    # configs/default.yaml:62-69
    dict(seed=0, P=2048, ne=3000, no=1024, ni=512, ns=1024, patch=False, rgb_loss_type='l1',
         w=dict(rgb_weight=30.0, perceptual_weight=0.0, eikonal_weight=50.0, mask_weight=3000.0, off_surface_weight=100.0, inside_weight=0.0,
                params_weight=100.0, skinning_weight=0.0)),
    # configs/arah-zju/ZJUMOCAP-313_4gpus.yaml:54-56 on top of the defaults; patch-style body mask (value 100 = border, loss.py:52-54)
    dict(seed=1, P=2100, ne=2500, no=700, ni=900, ns=1536, patch=True, rgb_loss_type='smoothed_l1',
         w=dict(rgb_weight=30.0, perceptual_weight=0.0, eikonal_weight=50.0, mask_weight=0.0, off_surface_weight=100.0, inside_weight=10.0,
                params_weight=100.0, skinning_weight=10.0)),
    # everything on, tiny and degenerate: no ray hit, no eikonal points, exact zeros under abs / norm
    dict(seed=2, P=7, ne=0, no=3, ni=2, ns=5, patch=False, rgb_loss_type='mse', degenerate=True,
         w=dict(rgb_weight=1.0, perceptual_weight=0.0, eikonal_weight=2.0, mask_weight=3.0, off_surface_weight=4.0, inside_weight=5.0,
                params_weight=6.0, skinning_weight=7.0)),
    dict(seed=3, P=300, ne=64, no=33, ni=17, ns=40, patch=False, rgb_loss_type='l1', degenerate=True,
         w=dict(rgb_weight=1.0, perceptual_weight=0.0, eikonal_weight=2.0, mask_weight=3.0, off_surface_weight=4.0, inside_weight=5.0,
                params_weight=6.0, skinning_weight=7.0)),
]
PARAM_SIZES = [768, 65536, 65536, 65536, 65536, 65536, 256]            # siren_modules.py:310-314: the seven weight matrices


def synth(c):
    rng = np.random.default_rng(c['seed'])
    P = c['P']
    f = lambda *s: rng.random(s).astype(np.float32)
    d = {'rgb_values': f(P, 3), 'rgb_gt': f(P, 3), 'network_body_mask': rng.random(P) < 0.6, 'off_surface_mask': rng.random(P) < 0.5,
         'body_mask': (rng.random(P) < 0.5).astype(np.uint8), 'sdf_output': f(P),
         'grad_theta': (rng.normal(size=(c['ne'], 3)) * 0.8).astype(np.float32),
         'off_surface_sdf': (rng.normal(size=c['no']) * 0.02).astype(np.float32), 'inside_sdf': (rng.normal(size=c['ni']) * 4e-4).astype(np.float32),
         'pred_weights': f(c['ns'], 24), 'sampled_weights': f(c['ns'], 24),
         'sdf_params': [(rng.normal(size=n) * 0.05).astype(np.float32) for n in PARAM_SIZES]}
    if c['patch']:
        d['body_mask'][rng.random(P) < 0.1] = 100
    if c.get('degenerate'):
        if c['seed'] == 2:
            d['network_body_mask'][:] = False
        d['rgb_values'][:2] = d['rgb_gt'][:2]                     # sign(0)
        d['sdf_output'][:2] = d['body_mask'][:2]
        d['pred_weights'][0] = d['sampled_weights'][0]
        if c['ne']:
            d['grad_theta'][0] = 0.0                              # norm backward at the origin
            d['grad_theta'][1] = [1.0, 0.0, 0.0]                  # |n| - 1 == 0
    return d


def run_reference(c, d):
    rh.install()
    import im2mesh.metaavatar_render  # noqa: F401
    from im2mesh.metaavatar_render.renderer.loss import IDHRLoss
    crit = IDHRLoss(rgb_loss_type=c['rgb_loss_type'], perceptual_loss_fn=None, **c['w'])
    t = lambda a, rg=False: torch.from_numpy(np.ascontiguousarray(a)).unsqueeze(0).requires_grad_(rg)
    leaves = {k: t(d[k], True) for k in ('rgb_values', 'sdf_output', 'pred_weights')}
    leaves.update({k: torch.from_numpy(d[k]).requires_grad_(True) for k in ('grad_theta', 'off_surface_sdf', 'inside_sdf')})
    params = [t(p, True) for p in d['sdf_params']]
    # 'sdf_output' is [1, P], as the renderer returns it (implicit_differentiable_renderer.py:229-237)
    mo = {'rgb_values': leaves['rgb_values'], 'sdf_output': leaves['sdf_output'], 'network_body_mask': t(d['network_body_mask']),
          'body_mask': t(d['body_mask']), 'off_surface_mask': t(d['off_surface_mask']), 'surface_normals': None, 'grad_theta': leaves['grad_theta'],
          'off_surface_sdf': leaves['off_surface_sdf'], 'inside_sdf': leaves['inside_sdf'], 'pred_weights': leaves['pred_weights'], 'sdf_params': params}
    out = crit(mo, {'rgb': t(d['rgb_gt']), 'sampled_weights': t(d['sampled_weights'])})
    out['loss'].sum().backward()
    res = {'terms.' + k: np.float64(v.detach().reshape(-1)[0]) for k, v in out.items()}
    res['loss_shape'] = np.array(out['loss'].shape, np.int64)
    for k, v in leaves.items():
        res['grad.' + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).detach().numpy().reshape(np.asarray(d[k]).shape)
    for i, p in enumerate(params):
        res[f'grad.sdf_params.{i}'] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().numpy().reshape(-1)
    return res


def main():
    for c in CASES:
        d = synth(c)
        res = run_reference(c, d)
        path = os.path.join(OUT, f"loss_s{c['seed']}.npz")
        # inputs are regenerated from the recipe by the tests (oracle.gen_golden_loss.synth); only outputs ship — except the big
        # parameter gradients, stored as a 512-entry head + sum + norm
        small = {}
        for k, v in res.items():
            if k.startswith('grad.sdf_params.') and v.size > 4096:
                small[k + '.head'] = v[:512]; small[k + '.sum'] = np.float64(v.astype(np.float64).sum()); small[k + '.norm'] = np.float64(np.linalg.norm(v.astype(np.float64)))
            else:
                small[k] = v
        np.savez_compressed(path, meta=json.dumps({k: v for k, v in c.items()}), **small)
        print(path, os.path.getsize(path) // 1024, 'KB', {k[6:]: float(v) for k, v in res.items() if k.startswith('terms.')})


if __name__ == '__main__':
    main()
