"""CPU: the C oracle (oracle/arah_oracle.c) against fixtures produced by the unmodified reference
(oracle/gen_golden.py).  This is what pins the oracle; the GPU tests then compare the CUDA path with both."""
import numpy as np
import pytest

from helpers import EXTRA_LASTPT_CASE, EXTRA_WSUM_CASES, GOLDEN_CASES, check_render, check_weights_sum, load_extra, load_golden, psnr


@pytest.mark.parametrize('name', GOLDEN_CASES)
def test_oracle_matches_reference(name):
    from oracle import oracle as orc
    fr, ref, meta = load_golden(name)
    out = orc.render(fr, threads=0)
    st = check_render(out, ref, label=name)
    # iteration counters agree with the reference's own (wrapped broyden) to within borderline cases
    iso_ref = meta['iso_calls'][0]['g_evals']
    corr_ref = meta['corr_calls'][0]['g_evals']
    n_pts = meta['corr_calls'][0]['n']
    assert abs(int(out['n_iso_evals'].sum()) - iso_ref) <= max(8, 0.05 * iso_ref)   # non-converging rays iterate chaotically
    corr_ours = int(out['n_corr_evals'].sum()) - n_pts          # ours also counts the J-init evaluation
    assert abs(corr_ours - corr_ref) <= 0.02 * corr_ref
    print(name, st)


def test_degenerate_rays_are_background():
    from oracle import oracle as orc
    fr, ref, meta = load_golden('n32_16x16_s2')
    nd = meta['n_degenerate']
    out = orc.render(fr)
    assert not out['trace.network_body_mask'][-nd:].any()
    np.testing.assert_allclose(out['trace.dists'][-nd:], fr.near_far[-nd:, 0])


def test_oracle_unit_functions_selfconsistent():
    """finite-difference checks of the analytic derivatives the oracle (and the kernels) rely on."""
    from oracle import oracle as orc
    fr, _, _ = load_golden('n32_16x16_s2')
    rng = np.random.default_rng(0)
    xn = rng.uniform(-0.5, 0.5, size=(64, 3)).astype(np.float32)
    s, g, _ = orc.sdf(fr, xn)
    h = 1e-3
    for k in range(3):
        e = np.zeros(3, np.float32); e[k] = h
        sp, _, _ = orc.sdf(fr, xn + e, grad=False)
        sm, _, _ = orc.sdf(fr, xn - e, grad=False)
        np.testing.assert_allclose((sp - sm) / (2 * h), g[:, k], atol=2e-2, rtol=5e-2)
    xh = rng.uniform(-0.4, 0.4, size=(64, 3)).astype(np.float32)
    w, xb, J = orc.skin(fr, xh)
    np.testing.assert_allclose(w.sum(1), 1.0, atol=1e-5)
    for k in range(3):
        e = np.zeros(3, np.float32); e[k] = h
        _, xp, _ = orc.skin(fr, xh + e, jac=False)
        _, xm, _ = orc.skin(fr, xh - e, jac=False)
        np.testing.assert_allclose((xp - xm) / (2 * h), J[:, :, k], atol=3e-2, rtol=5e-2)


def test_oracle_matches_reference_config0_64x64():
    """BASELINE configs[0] — the reference's own CPU-runnable case, a 64x64 frame (final outputs + per-ray tracer outputs; the
    fixture keeps no per-sample stage tensors)."""
    from oracle import oracle as orc
    fr, ref, meta = load_golden('zju377_64x64_s0')
    assert (fr.H, fr.W) == (64, 64) and fr.P == ref['rgb_values'].shape[0] and fr.P > 1000
    out = orc.render(fr, threads=0, stages=False)
    st = check_render(out, ref, label='config0', stages=False)
    iso_ref, corr_ref, n_pts = meta['iso_calls'][0]['g_evals'], meta['corr_calls'][0]['g_evals'], meta['corr_calls'][0]['n']
    assert abs(int(out['n_iso_evals'].sum()) - iso_ref) <= max(8, 0.05 * iso_ref)
    assert abs(int(out['n_corr_evals'].sum()) - n_pts - corr_ref) <= 0.02 * corr_ref
    print('config0', st)


def test_oracle_matches_reference_no_normal_colour_mode():
    """RenderingNetwork mode 'no_normal' (metaavatar_render/models/decoder.py:104-106; 414 inputs): the oracle and the product reach it
    by inserting three zero weight columns where the 'idr' layout has the normal — exact."""
    import torch
    from arah_release_b200 import renderer as R, synthetic as syn
    from oracle import oracle as orc
    fr, ref, meta = load_golden('nonormal_16x16_s9')
    assert fr.color_mode == 'no_normal'
    out = orc.render(fr, threads=0)
    print('no_normal', check_render(out, ref, label='no_normal'))
    # the product's expansion (torch) is the oracle's (numpy), column for column
    for mode, width in (('no_normal', 414), ('no_view_dir', 390)):
        W = np.random.default_rng(0).normal(size=(5, width)).astype(np.float32)
        assert np.array_equal(R._expand_color_weight(torch.from_numpy(W), mode).numpy(), syn.expand_color_weight(W, mode))


@pytest.mark.parametrize('name', EXTRA_WSUM_CASES)
def test_oracle_eval_weights_sum_matches_reference(name):
    """Eval weights_sum, the second output of get_rbg_value_vol_sdf (SURVEY §7 minimum slice: 1e-4)."""
    from oracle import oracle as orc
    fr, ref, meta = load_extra(name)
    out = orc.render(fr, threads=0, stages=False)
    assert (out['network_body_mask'].astype(bool) != ref['network_body_mask'].astype(bool)).mean() <= 0.002
    print(name, check_weights_sum(out['weights_sum'], ref, label=name))


def test_oracle_render_last_pt_matches_reference():
    """IDHRNetwork(render_last_pt=True): the last converged sample of a ray gets the interval 1e10 (:380-381)."""
    from oracle import oracle as orc
    fr, ref, meta = load_extra(EXTRA_LASTPT_CASE)
    assert fr.render_last_pt
    out = orc.render(fr, threads=0, stages=False)
    assert psnr(out['rgb_values'], ref['rgb_values']) >= 60.0
    print('last_pt', check_weights_sum(out['weights_sum'], ref, label='last_pt'))
    fr.render_last_pt = False                               # and the switch matters on this frame
    off = orc.render(fr, threads=0, stages=False)
    assert np.abs(off['weights_sum'] - out['weights_sum']).max() > 1e-3
