"""CPU: the numpy oracle of the image-space tail (oracle/images_oracle.py).

* validation images + PSNR: against the golden outputs of the UNMODIFIED reference `validation_step`
  (tests/golden/images_s*.npz, oracle/gen_golden_images.py).  Tolerance: images 1.2e-7 absolute (1 ulp of the [0,1] colour:
  torch's vector norm sums the three squares in another order), scattered pixels bit-exact, PSNR 1e-9 dB.
* rasteriser (pytorch3d is absent: parity unpinned): cross-checked against an independent float64 ray caster — same face at
  every pixel whose hit is not within rounding of a triangle edge or of a second surface; depth 2e-6 relative.
"""
import numpy as np
import pytest

from helpers_images import handmade_mesh, iso_mesh, load_images_golden, make_camera, raycast_pix_to_face
from oracle import images_oracle as io


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_frame_images_match_reference_validation_step(seed):
    g = load_images_golden(seed)
    H, W = int(g['H']), int(g['W'])
    pred_pixels, pred_normals = io.frame_images(g['rgb'], g['points_cam'], g['pix'], H, W)
    assert np.array_equal(pred_pixels, g['ref.rgb_pred'])
    assert np.array_equal(io.scatter_rows(g['gt'], g['pix'], H, W), g['ref.rgb_gt'])
    assert np.abs(pred_normals - g['ref.normal_pred']).max() <= 1.2e-7
    # the fixtures exercise the special values: background (0/0 -> NaN -> 0), x/0 -> inf -> NaN -> 0 with the other lanes 0.5
    assert (g['ref.normal_pred'] == 0).any() and (g['ref.normal_pred'] == 0.5).any()
    mse, psnr = io.psnr_metric(g['rgb'], g['gt'])
    assert abs(psnr - float(g['ref.psnr'])) <= 1e-9


def test_look_at_matches_closed_form():
    R, T = io.look_at_view_transform(2.0, 0.0, 0.0)
    np.testing.assert_allclose(R, np.diag([-1.0, 1.0, -1.0]), atol=1e-6)
    np.testing.assert_allclose(T, [0, 0, 2], atol=1e-6)
    R, T = io.look_at_view_transform(2.0, 0.0, 180.0)
    np.testing.assert_allclose(R, np.eye(3), atol=1e-6)
    np.testing.assert_allclose(T, [0, 0, 2], atol=1e-6)


def test_opencv_camera_maps_pixel_centres():
    """A point that OpenCV projects to (u, v) must land on the NDC centre of pixel (u - 0.5, v - 0.5)."""
    H, W = 40, 56
    R, T, K = make_camera(H, W, shift=(2.5, -1.25))
    cam = io.opencv_camera(R, T, K, H, W)
    for (x, y) in [(0, 0), (W - 1, H - 1), (13, 29)]:
        d = np.linalg.inv(K.astype(np.float64)) @ np.array([x + 0.5, y + 0.5, 1.0])
        Xw = R.astype(np.float64).T @ (d * 2.7 - T.astype(np.float64))
        ndc = io.project(Xw[None].astype(np.float32), cam)[0]
        assert abs(ndc[0] - io.pix_to_ndc(W - 1 - x, W, H)) < 2e-5 and abs(ndc[1] - io.pix_to_ndc(H - 1 - y, H, W)) < 2e-5
        assert abs(ndc[2] - 2.7) < 1e-5


@pytest.mark.parametrize('mesh,H,W', [('hand', 48, 64), ('hand', 64, 40), ('torus', 72, 72), ('two_spheres', 64, 96)])
def test_rasterize_against_float64_ray_caster(mesh, H, W):
    if mesh == 'hand':
        v, f = handmade_mesh()
        R, T, K = np.eye(3, dtype=np.float32), np.zeros(3, np.float32), make_camera(H, W)[2]
    else:
        v, f = iso_mesh(mesh, 20)
        R, T, K = make_camera(H, W)
        T = T + np.array([0, 0, 2.6], np.float32)
    p2f, zbuf = io.rasterize(io.project(v, io.opencv_camera(R, T, K, H, W)), f, H, W)
    rf, margin, gap = raycast_pix_to_face(v, f, R, T, K, H, W)
    if mesh == 'hand':
        rf = np.where(rf == 8, 0, rf)                 # face 8 duplicates face 0: the rasteriser keeps the lower index
        assert 0 in p2f and 2 in p2f and 4 in p2f and 5 not in p2f and 6 not in p2f and 7 not in p2f and 8 not in p2f
    clear = ((margin > 1e-4) & (gap > 1e-4)) | (rf < 0)
    edge_px = ~clear
    assert clear.mean() > 0.9
    # away from edges / depth ties the two algorithms must agree exactly ...
    bg_near_edge = (rf < 0) & (p2f >= 0)              # ... a background pixel can only be claimed by a hair-line edge hit
    assert (p2f[clear & (rf >= 0)] == rf[clear & (rf >= 0)]).all()
    assert bg_near_edge.sum() <= 2
    assert (p2f >= 0).sum() > 0.05 * H * W
    fg = clear & (rf >= 0)
    v_cam = v.astype(np.float64) @ R.astype(np.float64).T + T
    # depth: the ray caster's t is the camera-space z of the hit
    ys, xs = np.nonzero(fg)
    a, b, c = (v_cam[f[rf[fg], k]] for k in range(3))
    n = np.cross(b - a, c - a)
    d = np.stack([(xs + 0.5 - K[0, 2]) / K[0, 0], (ys + 0.5 - K[1, 2]) / K[1, 1], np.ones(len(xs))], -1)
    t = np.einsum('ij,ij->i', a, n) / np.einsum('ij,ij->i', d, n)
    np.testing.assert_allclose(zbuf[fg], t, rtol=2e-5)
    assert (zbuf[p2f < 0] == -1).all()
    assert edge_px.sum() < 0.1 * H * W


def test_normal_maps_of_a_sphere():
    """Front / back canonical views of a sphere: the decoded normal at a pixel is the outward normal of the visible hemisphere,
    i.e. colour = (n + 1) / 2 with n ~ the unit vector from the centre to the visible surface point; background 0.5 / 0."""
    v, f = iso_mesh('sphere', 24)
    R, T, K = make_camera(64, 64)
    T = T + np.array([0, 0, 2.6], np.float32)
    maps = io.normal_maps(v, f, v, R, T, K, 64, 64)
    front, back, posed = maps['normal_cano_front'], maps['normal_cano_back'], maps['output_normal']
    assert front.shape == (64, 64, 3) and front.dtype == np.float32
    for img, zsign in ((front, 1.0), (back, -1.0)):
        assert np.all(img[0, 0] == 0.5) and np.all(img[-1, -1] == 0.5)               # background 0 -> 0.5
        c = img[32, 32] * 2 - 1                                                        # centre pixel looks along -/+ z
        assert c[2] * zsign > 0.97
    # the front camera sits at +z: view x = -world x and pytorch3d's +X is the image's left, so world +x is on the right
    assert front[32, 20, 0] < 0.4 and front[32, 44, 0] > 0.6
    assert back[32, 20, 0] > 0.6 and back[32, 44, 0] < 0.4
    assert front[20, 32, 1] > 0.6 and front[44, 32, 1] < 0.4                           # +y up
    # posed view: normals negated then rotated into the camera frame: the centre pixel's outward normal points to the camera
    # (-z in OpenCV camera coordinates), negated -> +z -> blue channel high; background -1 -> 0
    assert np.all(posed[0, 0] == 0.0)
    assert posed[32, 32, 2] > 0.95


def test_ssim_oracle_against_brute_force_and_cv2():
    """SSIM (skimage is absent: unpinned against it): the scipy-based restatement against a filter-free brute force, the bounding
    rectangle against cv2.boundingRect itself, and the textbook properties (identical images -> 1, symmetric)."""
    import cv2
    from numpy.lib.stride_tricks import sliding_window_view as sw
    rng = np.random.default_rng(0)
    H, W = 40, 52
    a = rng.random((H, W, 3)).astype(np.float32)
    b = np.clip(a + 0.1 * rng.standard_normal(a.shape), 0, 1).astype(np.float32)
    mask = np.zeros((H, W), bool)
    mask[5:33, 9:47] = rng.random((28, 38)) < 0.7
    mask[5, 20] = mask[32, 46] = mask[17, 9] = True
    assert io.bounding_rect(mask) == tuple(cv2.boundingRect(mask.astype(np.uint8)))
    assert io.bounding_rect(np.zeros((4, 4), bool)) == tuple(cv2.boundingRect(np.zeros((4, 4), np.uint8)))
    x, y, w, h = io.bounding_rect(mask)

    def brute(X, Y):
        X, Y = X.astype(np.float64), Y.astype(np.float64)
        m = lambda A: sw(A, (7, 7)).mean((-1, -2))
        ux, uy, uxx, uyy, uxy = m(X), m(Y), m(X * X), m(Y * Y), m(X * Y)
        cn = 49 / 48
        vx, vy, vxy = cn * (uxx - ux * ux), cn * (uyy - uy * uy), cn * (uxy - ux * uy)
        C1, C2 = 0.02 ** 2, 0.06 ** 2
        return (((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))).mean()
    ca, cb = a[y:y + h, x:x + w], b[y:y + h, x:x + w]
    ref = float(np.mean([brute(ca[..., c], cb[..., c]) for c in range(3)]))
    assert abs(io.ssim_metric(a, b, mask) - ref) <= 1e-12
    assert abs(io.ssim_metric(a, a, mask) - 1.0) <= 1e-12
    assert abs(io.ssim_metric(a, b, mask) - io.ssim_metric(b, a, mask)) <= 1e-12
    with pytest.raises(ValueError):
        io.ssim_metric(a, b, np.pad(np.ones((3, 9), bool), ((0, H - 3), (0, W - 9))))
