"""TEST INFRASTRUCTURE — numpy restatement of the reference's training loss and of its gradient.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product never does.

Follows /root/reference/im2mesh/metaavatar_render/renderer/loss.py: `IDHRLoss.forward` (:122-200) and the term functions
(:46-120).  The gradients are the closed forms torch autograd produces for those expressions (abs -> sign with sign(0) = 0,
2-norm -> x / |x| with 0 at the origin, L1 / MSE / SmoothL1(beta = 0.1) with reduction 'sum').
PARITY PINNED: tests/golden/loss_s*.npz hold the terms and the autograd gradients of the unmodified `IDHRLoss`
(oracle/gen_golden_loss.py).  The perceptual term (LPIPS, a VGG network) is outside this path: weight must be 0.
"""
import numpy as np

TERMS = ('loss', 'rgb_loss', 'perceptual_loss', 'eikonal_loss', 'mask_loss', 'off_surface_loss', 'inside_loss', 'sdf_params_loss', 'skinning_loss')
SMOOTH_BETA = 0.1


def idhr_loss(cfg, inp, want_grads=True):
    """cfg: dict of the eight weights + rgb_loss_type.  inp: numpy arrays named as in model_outputs / ground_truth, batch dim
    dropped and already cut to the first 2048 rays (:124-127,132): rgb_values [N,3], rgb_gt [N,3], network_body_mask [N] bool,
    body_mask [N] (uint8; 100 marks patch borders), off_surface_mask [N] bool, sdf_output [N], grad_theta [ne,3],
    off_surface_sdf [no], inside_sdf [ni], pred_weights [ns,J], sampled_weights [ns,J], sdf_params list of 1-D arrays.
    -> (terms dict of float64, grads dict of float32 arrays = d loss / d input)."""
    f64 = lambda a: np.asarray(a, np.float64)
    N = int(np.asarray(inp['body_mask']).size)
    t = {k: 0.0 for k in TERMS}
    g = {}
    body = np.asarray(inp['body_mask'])
    if cfg['rgb_weight'] > 0:                                                   # get_rgb_loss :46-61
        m = np.asarray(inp['network_body_mask']).astype(bool)
        if body.size and body.max() > 1:
            m = m & (body != 100)
        d = f64(inp['rgb_values']) - f64(inp['rgb_gt'])
        typ = cfg.get('rgb_loss_type', 'l1')
        if typ == 'l1':
            v, dv = np.abs(d), np.sign(d)
        elif typ == 'mse':
            v, dv = d * d, 2 * d
        else:
            small = np.abs(d) < SMOOTH_BETA
            v, dv = np.where(small, 0.5 * d * d / SMOOTH_BETA, np.abs(d) - 0.5 * SMOOTH_BETA), np.where(small, d / SMOOTH_BETA, np.sign(d))
        t['rgb_loss'] = float((v * m[:, None]).sum() / N) if m.any() else 0.0
        g['rgb_values'] = (cfg['rgb_weight'] * dv * m[:, None] / N).astype(np.float32)
    if cfg['mask_weight'] > 0:                                                  # get_mask_loss_vol_sdf :96-105
        # model_outputs['sdf_output'] is [1, P] (implicit_differentiable_renderer.py:229-237): `weights_output[off] - gt` is 1-D and
        # torch.norm(dim=-1) is the 2-norm of the whole vector (:100-101) — one norm, not a sum over rays
        off = np.asarray(inp['off_surface_mask']).astype(bool)
        w_all = f64(inp['sdf_output']).reshape(-1)
        assert w_all.size == N, 'the reference does not cut sdf_output to 2048 rays (:143): shapes must agree'
        d = (w_all - body.astype(np.float64)) * off
        norm = np.sqrt((d * d).sum())
        t['mask_loss'] = float(norm / N) if off.any() else 0.0
        g['sdf_output'] = (cfg['mask_weight'] * (d / norm if norm > 0 else 0 * d) / N).astype(np.float32)
    if cfg['eikonal_weight'] > 0:                                               # get_eikonal_loss :88-94
        gt_ = f64(inp['grad_theta']).reshape(-1, 3)
        n = np.sqrt((gt_ * gt_).sum(-1))
        t['eikonal_loss'] = float(np.abs(n - 1).sum() / N) if gt_.shape[0] else 0.0
        with np.errstate(divide='ignore', invalid='ignore'):
            unit = np.where(n[:, None] > 0, gt_ / n[:, None], 0.0)
        g['grad_theta'] = (cfg['eikonal_weight'] * np.sign(n - 1)[:, None] * unit / N).astype(np.float32)
    if cfg['off_surface_weight'] > 0:                                           # get_off_surface_loss :107-109
        e = np.exp(-1e2 * f64(inp['off_surface_sdf']).reshape(-1))
        t['off_surface_loss'] = float(e.sum() / N)
        g['off_surface_sdf'] = (cfg['off_surface_weight'] * -1e2 * e / N).astype(np.float32)
    if cfg['inside_weight'] > 0:                                                # get_inside_loss :119-120
        s = 1.0 / (1.0 + np.exp(-5e3 * f64(inp['inside_sdf']).reshape(-1)))
        t['inside_loss'] = float(s.sum() / N)
        g['inside_sdf'] = (cfg['inside_weight'] * 5e3 * s * (1 - s) / N).astype(np.float32)
    if cfg['params_weight'] > 0:                                                # get_sdf_params_loss :111-115
        ps = [f64(p).reshape(-1) for p in inp['sdf_params']]
        n_params = sum(p.size for p in ps)
        norm = np.sqrt(sum(float((p * p).sum()) for p in ps))
        t['sdf_params_loss'] = norm / n_params
        g['sdf_params'] = [((cfg['params_weight'] * p / (norm * n_params)) if norm > 0 else 0 * p).astype(np.float32) for p in ps]
    if cfg['skinning_weight'] > 0:                                              # get_skinning_loss :117-118
        d = f64(inp['pred_weights']) - f64(inp['sampled_weights'])
        ns = d.reshape(-1, d.shape[-1]).shape[0]
        t['skinning_loss'] = float(np.abs(d).sum() / ns)
        g['pred_weights'] = (cfg['skinning_weight'] * np.sign(d) / ns).astype(np.float32)
    t['loss'] = sum(cfg[k] * t[v] for k, v in (('rgb_weight', 'rgb_loss'), ('eikonal_weight', 'eikonal_loss'), ('mask_weight', 'mask_loss'),
                                               ('off_surface_weight', 'off_surface_loss'), ('inside_weight', 'inside_loss'),
                                               ('params_weight', 'sdf_params_loss'), ('skinning_weight', 'skinning_loss')))
    return (t, g) if want_grads else t
