"""Shared parity helpers: golden loading, tolerances, PSNR (formula of /root/reference/im2mesh/utils/eval.py:6-9)."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
GOLDEN_CASES = ['zju377_24x24_s0', 'cano_20x20_s1', 'n32_16x16_s2', 'h36m_n160_12x12_s4', 'mono_noview_16x16_s8']

# Tolerances (fp32 path).  The reference is batched MKL fp32, ours sequential-k FMA fp32.
TOL = dict(
    ray_mask_mismatch=0.002,      # fraction of rays whose hit mask may differ (threshold 1e-5 borderline cases)
    sample_mask_mismatch=0.001,   # fraction of samples whose converged flag may differ
    depth_linf=1e-4,              # metres, on rays both sides call converged
    pts_linf=3e-4,                # normalised canonical coords, on commonly converged samples
    T_linf=5e-4,                  # forward transforms on commonly converged samples
    rgb_psnr_min=60.0,            # dB, PSNR(ours, reference) over all rays
    dpsnr_max=0.05,               # dB, |PSNR(ours, GT) - PSNR(ref, GT)| against a fixed pseudo ground truth
)


def psnr(a, b):
    mse = np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2)
    return float(-10.0 * np.log10(max(mse, 1e-30)))


def pseudo_gt(ref_rgb, seed=123, sigma=0.03):
    """Fixed pseudo ground truth = reference image + seeded noise (PSNR(ref, GT) ~ 30 dB, like real data)."""
    rng = np.random.default_rng(seed)
    return np.clip(ref_rgb + rng.normal(scale=sigma, size=ref_rgb.shape), 0, 1).astype(np.float32)


def load_golden(name):
    from arah_release_b200 import synthetic as syn
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    meta = json.loads(str(z['meta']))
    fr = syn.make_frame(**meta['make_frame'])
    nd = meta['n_degenerate']
    if nd:
        fr.ray_dirs = np.concatenate([fr.ray_dirs, fr.ray_dirs[:nd]], 0)
        fr.near_far = np.concatenate([fr.near_far, np.repeat(fr.near_far[:nd, 1:2], 2, axis=1)], 0)
        fr.pix = np.concatenate([fr.pix, fr.pix[:nd]])
    ref = {k.replace('__', '.'): z[k] for k in z.files if k != 'meta'}
    assert fr.P == meta['P']
    return fr, ref, meta


def check_render(out, ref, label='', tol=TOL, stages=True):
    """Compare a render result dict (keys as oracle.oracle.render) with reference outputs. Returns stats dict."""
    st = {}
    P = ref['rgb_values'].shape[0]
    m_ref, m_out = ref['trace.network_body_mask'].astype(bool), np.asarray(out['trace.network_body_mask']).astype(bool)
    st['ray_mask_mismatch'] = float((m_ref != m_out).mean())
    assert st['ray_mask_mismatch'] <= tol['ray_mask_mismatch'], (label, st)
    both = m_ref & m_out
    st['depth_linf'] = float(np.abs(ref['trace.dists'][both] - out['trace.dists'][both]).max()) if both.any() else 0.0
    assert st['depth_linf'] <= tol['depth_linf'], (label, st)
    # rays neither side converged keep dists = near exactly
    neither = ~m_ref & ~m_out
    if neither.any():
        assert np.abs(ref['trace.dists'][neither] - out['trace.dists'][neither]).max() <= 1e-6, label
    st['pts_hat_linf'] = float(np.abs(ref['trace.points_hat_norm'][both] - out['trace.points_hat_norm'][both]).max()) if both.any() else 0.0
    assert st['pts_hat_linf'] <= tol['pts_linf'], (label, st)
    assert (ref['network_body_mask'].astype(bool) != np.asarray(out['network_body_mask']).astype(bool)).mean() <= tol['ray_mask_mismatch'], label
    if stages and 'trace.sampler_converge_mask' in out:
        c_ref = ref['trace.sampler_converge_mask'].astype(bool)
        c_out = np.asarray(out['trace.sampler_converge_mask']).astype(bool)
        st['sample_mask_mismatch'] = float((c_ref != c_out).mean())
        assert st['sample_mask_mismatch'] <= tol['sample_mask_mismatch'], (label, st)
        cb = c_ref & c_out & both[:, None] | (c_ref & c_out & neither[:, None])
        st['sampled_dists_linf'] = float(np.abs(ref['trace.sampled_dists'] - out['trace.sampled_dists'])[cb].max())
        assert st['sampled_dists_linf'] <= tol['depth_linf'], (label, st)
        st['sampled_pts_linf'] = float(np.abs(ref['trace.sampled_pts'] - out['trace.sampled_pts'])[cb].max())
        assert st['sampled_pts_linf'] <= tol['pts_linf'], (label, st)
        if 'trace.sampled_transforms_head' in ref and 'trace.sampled_transforms' in out:
            n = ref['trace.sampled_transforms_head'].shape[0]
            d = np.abs(ref['trace.sampled_transforms_head'] - out['trace.sampled_transforms'][:n])[cb[:n]]
            st['T_linf'] = float(d.max()) if d.size else 0.0
            assert st['T_linf'] <= tol['T_linf'], (label, st)
    st['rgb_psnr'] = psnr(out['rgb_values'], ref['rgb_values'])
    assert st['rgb_psnr'] >= tol['rgb_psnr_min'], (label, st)
    gt = pseudo_gt(ref['rgb_values'])
    st['dpsnr'] = abs(psnr(out['rgb_values'], gt) - psnr(ref['rgb_values'], gt))
    assert st['dpsnr'] <= tol['dpsnr_max'], (label, st)
    pc_both = both
    st['points_cam_linf'] = float(np.abs(ref['points_cam'][pc_both] - out['points_cam'][pc_both]).max()) if pc_both.any() else 0.0
    assert st['points_cam_linf'] <= 2 * tol['depth_linf'], (label, st)
    return st


EXTRA_WSUM_CASES = ['wsum_zju377_24x24_s0', 'wsum_cano_20x20_s1']
EXTRA_LASTPT_CASE = 'lastpt_16x16_s3'


def load_extra(name):
    """Fixtures of oracle/gen_golden_extra.py: (frame, reference arrays, meta); frame.render_last_pt is set from the recipe."""
    from arah_release_b200 import synthetic as syn
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    meta = json.loads(str(z['meta']))
    fr = syn.make_frame(**meta['make_frame'])
    fr.render_last_pt = bool(meta['render_last_pt'])
    assert fr.P == meta['P']
    return fr, {k: z[k] for k in z.files if k != 'meta'}, meta


def check_weights_sum(ws, ref, label='', atol=1e-4, max_outliers=0.002):
    """Eval weights_sum (implicit_differentiable_renderer.py:392) against the reference at 1e-4 on rays both sides put in the volume
    mask; rays whose converged-sample set differs by a borderline sample may deviate and must be rare."""
    m = ref['network_body_mask'].astype(bool)
    d = np.abs(np.asarray(ws, np.float64)[m] - ref['weights_sum'].astype(np.float64)[m])
    frac = float((d > atol).mean()) if d.size else 0.0
    assert frac <= max_outliers, (label, frac, float(d.max()))
    return {'wsum_linf_inliers': float(d[d <= atol].max()) if (d <= atol).any() else 0.0, 'wsum_outlier_frac': frac}
