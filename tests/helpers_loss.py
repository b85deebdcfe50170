"""Shared pieces of the training-loss tests: fixture loading (inputs are rebuilt from the seeded recipe, outputs come from the
unmodified reference, oracle/gen_golden_loss.py) and the comparison with its tolerances.

Tolerances (floating point): terms 2e-6 relative (the reference sums fp32 in torch's reduction order; oracle and kernels
accumulate in fp64), gradients 2e-6 of the largest entry of each gradient tensor + exact zeros where the reference has zeros."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RAY_KEYS = ('rgb_values', 'rgb_gt', 'network_body_mask', 'body_mask', 'off_surface_mask')
TERMS = ('loss', 'rgb_loss', 'perceptual_loss', 'eikonal_loss', 'mask_loss', 'off_surface_loss', 'inside_loss', 'sdf_params_loss', 'skinning_loss')
LOSS_SEEDS = (0, 1, 2, 3)


def load_loss_golden(seed):
    from oracle.gen_golden_loss import synth
    z = np.load(os.path.join(GOLDEN, f'loss_s{seed}.npz'))
    case = json.loads(str(z['meta']))
    full = synth(case)
    cut = dict(full)
    for k in RAY_KEYS:                                   # IDHRLoss.forward cuts these to the first 2048 rays (loss.py:124-127,132) ...
        cut[k] = full[k][:2048]
    if case['w']['mask_weight'] <= 0:                    # ... but not sdf_output (:143); it only matters when the mask term is on
        cut['sdf_output'] = full['sdf_output'][:2048]
    cfg = dict(case['w'], rgb_loss_type=case['rgb_loss_type'])
    return cfg, cut, full, {k: z[k] for k in z.files if k != 'meta'}


def check_loss(terms, grads, ref, full, rtol=2e-6):
    """terms: name -> float; grads: name -> array (d loss / d input over the rows the loss saw; missing = all zero)."""
    for k in TERMS:
        r = float(ref['terms.' + k])
        assert abs(float(terms[k]) - r) <= rtol * max(abs(r), 1e-12) + 1e-30, (k, float(terms[k]), r)
    for k in ('rgb_values', 'sdf_output', 'grad_theta', 'off_surface_sdf', 'inside_sdf', 'pred_weights'):
        r = ref['grad.' + k]
        mine = np.zeros_like(r)
        if grads.get(k) is not None:
            g = np.asarray(grads[k], np.float32)
            if k in ('rgb_values', 'sdf_output'):
                mine[:g.shape[0]] = g.reshape((-1,) + r.shape[1:])
            else:
                mine = g.reshape(r.shape)
        if r.size == 0:
            continue
        scale = float(np.abs(r).max())
        assert np.abs(mine - r).max() <= rtol * scale + 1e-30, (k, float(np.abs(mine - r).max()), scale)
        assert ((mine == 0) == (r == 0)).all(), k + ': zero pattern (masks, sign(0), norm at the origin) differs'
    for i in range(len(full['sdf_params'])):
        key = f'grad.sdf_params.{i}'
        g = None if grads.get('sdf_params') is None else np.asarray(grads['sdf_params'][i], np.float32).reshape(-1)
        if key in ref:
            r = ref[key]
            g = np.zeros_like(r) if g is None else g
            assert np.abs(g - r).max() <= rtol * float(np.abs(r).max()) + 1e-30, key
        else:
            g = np.zeros(full['sdf_params'][i].size, np.float32) if g is None else g
            assert np.abs(g[:512] - ref[key + '.head']).max() <= rtol * float(np.abs(ref[key + '.head']).max()) + 1e-30, key
            assert abs(float(g.astype(np.float64).sum()) - float(ref[key + '.sum'])) <= 1e-4 * abs(float(ref[key + '.norm'])) + 1e-30, key
            assert abs(float(np.linalg.norm(g.astype(np.float64))) - float(ref[key + '.norm'])) <= 1e-5 * float(ref[key + '.norm']) + 1e-30, key
