"""CPU: the ray set-up oracle (oracle/rays_oracle.py) against
  (a) fixtures produced by the reference's own functions (tests/golden/rays_s*.npz, oracle/gen_golden_rays.py) and
  (b) cv2 itself for the fillPoly / line restatement (cv2 is a dependency of the reference, importable here)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(seed):
    return dict(np.load(os.path.join(GOLDEN, f'rays_s{seed}.npz')))


def compare_rays(out, g, label=''):
    """pixel list bit-exact; directions / near / far to fp32 rounding (the reference's dots go through BLAS)."""
    H, W = int(g['H']), int(g['W'])
    pix, rd, nf = out['pix'], out['ray_dirs'], out['near_far']
    if 'sample_idx' in g:
        assert pix.shape[0] == int(g['n_rays']) and int(pix.astype(np.int64).sum()) == int(g['pix_sum']), label
        idx = g['sample_idx']
        pix, rd, nf = pix[idx], rd[idx], nf[idx]
    assert np.array_equal(pix, g['pix']), label
    np.testing.assert_allclose(rd, g['ray_dirs'], atol=2e-7, rtol=0, err_msg=label)
    np.testing.assert_allclose(nf, g['near_far'], atol=0, rtol=2e-6, err_msg=label)


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_oracle_matches_reference_ray_setup(seed):
    from oracle import rays_oracle as ro
    g = load(seed)
    H, W = int(g['H']), int(g['W'])
    out = ro.gen_rays(g['K'], g['R'], g['T'], g['bounds'], H, W)
    ref_mask = np.unpackbits(g['bound_mask_bits'])[:H * W].reshape(H, W)
    assert np.array_equal(out['bound_mask'], ref_mask), 'bounding-box mask differs from get_bound_2d_mask (cv2.fillPoly)'
    assert int(out['bound_mask'].sum()) == int(g['n_bound'])
    compare_rays(out, g, f'oracle s{seed}')
    np.testing.assert_allclose(out['cam_loc'], g['cam_loc'], atol=1e-6)


def test_fill_poly_and_line_match_cv2():
    cv2 = pytest.importorskip('cv2')
    from oracle import rays_oracle as ro
    rng = np.random.default_rng(0)
    H, W = 72, 88
    for _ in range(1500):                                   # lines, incl. end points far outside the image
        p0, p1 = rng.integers(-50, 130, 2).tolist(), rng.integers(-50, 130, 2).tolist()
        a, b = np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint8)
        cv2.line(a, tuple(p0), tuple(p1), 1)
        ro.draw_line(b, p0, p1)
        assert np.array_equal(a, b), (p0, p1)
    bad_inside = bad_clipped = n_inside = n_clipped = 0
    for t in range(1500):                                   # quadrilaterals in the reference's index patterns
        ang = np.sort(rng.uniform(0, 2 * np.pi, 4)); c = rng.uniform(20, 60, 2); r = rng.uniform(3, 40, 2)
        q = np.round(np.stack([c[0] + r[0] * np.cos(ang), c[1] + r[1] * np.sin(ang)], 1)).astype(np.int32)
        pts = [q[[0, 1, 2, 3, 0]], q[[0, 1, 3, 2, 1]], q][t % 3]
        inside = (q >= 0).all() and (q[:, 0] < W).all() and (q[:, 1] < H).all()
        a, b = np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint8)
        cv2.fillPoly(a, [pts.reshape(-1, 1, 2)], 1)
        ro.fill_poly(b, pts.tolist())
        same = np.array_equal(a, b)
        if inside:
            n_inside += 1; bad_inside += not same
        else:
            n_clipped += 1; bad_clipped += not same
    print(f'fillPoly restatement: {bad_inside}/{n_inside} differing polygons inside the image, {bad_clipped}/{n_clipped} partially outside')
    assert bad_inside == 0 and n_inside > 300
    assert bad_clipped <= 0.01 * n_clipped                 # degenerate slivers of self-overlapping, clipped polygons


def test_pose_smpl_oracle_shapes_and_margin():
    from arah_release_b200 import synthetic as syn
    from oracle import rays_oracle as ro
    p = syn.make_smpl_pose_inputs(0)
    verts, bounds = ro.pose_smpl(**p)
    assert verts.shape == (6890, 3) and verts.dtype == np.float32 and bounds.shape == (2, 3)
    assert np.allclose(bounds[0], verts.min(0) - 0.05, atol=1e-6) and np.allclose(bounds[1], verts.max(0) + 0.05, atol=1e-6)
    p2 = dict(p); p2['pose_feature'] = np.zeros_like(p['pose_feature'])
    v2, _ = ro.pose_smpl(**p2)
    assert np.abs(v2 - verts).max() > 1e-4                  # the pose blend shapes matter
