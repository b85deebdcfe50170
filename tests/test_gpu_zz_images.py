"""GPU parity for the image-space tail (SURVEY §8 rows f4 / f1 after the renderer): arah_frame_images, arah_psnr,
arah_rasterize_mesh, arah_face_normal_image through the C ABI (arah_release_b200/images.py) against

* the golden outputs of the unmodified reference `validation_step` (tests/golden/images_s*.npz): scattered pixels bit-exact,
  normal map 1.2e-7 absolute (1 ulp of the [0,1] colour), PSNR 1e-5 dB / MSE 1e-6 relative (the kernel accumulates in fp64, numpy
  pairwise in fp32);
* the numpy oracle (oracle/images_oracle.py): pix_to_face (integer) bit-exact, depth buffer and normal images bit-exact too —
  kernel and oracle round every fp32 operation once in the same order (csrc/arah_image_core.h, pinned on the host by
  tests/test_images_host.py);
* at full size (512 x 512, a 128^3-lattice mesh) additionally size-independent properties: bit-identical re-runs, invariance
  under a permutation of the face list (the oracle itself is cross-checked against a float64 ray caster on the CPU,
  tests/test_images_oracle.py).
"""
import numpy as np
import pytest
import torch

from helpers_images import handmade_mesh, iso_mesh, load_images_golden, make_camera

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _fi():
    from arah_release_b200.images import FrameImages
    return FrameImages(DEV)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_frame_images_match_reference_validation_step(seed):
    from oracle import images_oracle as io
    g = load_images_golden(seed)
    H, W = int(g['H']), int(g['W'])
    fi = _fi()
    pp, pn = fi.assemble(_t(g['rgb']), _t(g['points_cam']), _t(g['pix']), H, W)
    gt_img, _ = fi.assemble(_t(g['gt']), None, _t(g['pix']), H, W, normals=False)
    mse, psnr = fi.psnr(_t(g['rgb']), _t(g['gt']))
    torch.cuda.synchronize()
    pp, pn, gt_img = pp.cpu().numpy(), pn.cpu().numpy(), gt_img.cpu().numpy()
    assert np.array_equal(pp, g['ref.rgb_pred']) and np.array_equal(gt_img, g['ref.rgb_gt'])
    o_pp, o_pn = io.frame_images(g['rgb'], g['points_cam'], g['pix'], H, W)
    assert np.array_equal(pn, o_pn)
    assert np.abs(pn - g['ref.normal_pred']).max() <= 1.2e-7
    o_mse, o_psnr = io.psnr_metric(g['rgb'], g['gt'])
    assert abs(mse - o_mse) <= 1e-6 * o_mse
    assert abs(psnr - float(g['ref.psnr'])) <= 1e-5


def test_frame_images_edge_cases():
    """Empty ray list, a full-image mask, a single-pixel image, PSNR of identical lists (mse 0 -> +inf, as numpy)."""
    from oracle import images_oracle as io
    fi = _fi()
    rng = np.random.default_rng(3)
    pp, pn = fi.assemble(torch.empty(0, 3, device=DEV), torch.empty(0, 3, device=DEV), torch.empty(0, dtype=torch.int32, device=DEV), 6, 7)
    e = np.zeros((0, 3), np.float32)
    o_pp, o_pn = io.frame_images(e, e, np.zeros(0, np.int32), 6, 7)           # all background: NaN -> -1 -> 0, except the last
    assert np.array_equal(pp.cpu().numpy(), o_pp) and np.array_equal(pn.cpu().numpy(), o_pn)      # pixel: (0, 0, 1) -> (.5, .5, 1)
    assert pp.abs().max().item() == 0 and pn[:-1].abs().max().item() == 0 and pn[-1, -1].tolist() == [0.5, 0.5, 1.0]
    H, W = 9, 5
    rgb, pts = rng.random((H * W, 3)).astype(np.float32), rng.normal(size=(H * W, 3)).astype(np.float32)
    pix = np.arange(H * W, dtype=np.int32)
    pp, pn = fi.assemble(_t(rgb), _t(pts), _t(pix), H, W)
    o_pp, o_pn = io.frame_images(rgb, pts, pix, H, W)
    assert np.array_equal(pp.cpu().numpy(), o_pp) and np.array_equal(pn.cpu().numpy(), o_pn)
    pp, pn = fi.assemble(_t(rgb[:1]), _t(pts[:1]), _t(pix[:1]), 1, 1)
    assert np.array_equal(pn.cpu().numpy(), io.frame_images(rgb[:1], pts[:1], pix[:1], 1, 1)[1])
    mse, psnr = fi.psnr(_t(rgb), _t(rgb))
    assert mse == 0.0 and psnr == float('inf')
    from arah_release_b200 import _lib
    with pytest.raises(_lib.ArahError):
        fi.psnr(_t(rgb), _t(rgb[:-1]))
    with pytest.raises(_lib.ArahError):
        fi.psnr(torch.empty(0, device=DEV), torch.empty(0, device=DEV))


def test_frame_images_full_size_512():
    """BASELINE-size frame (512 x 512, ~1.3e5 bbox rays): the assembled images bit-exact against the oracle, plus size-independent
    properties — scatter -> gather round trip, untouched background, value range, bit-identical re-runs."""
    from oracle import images_oracle as io
    H = W = 512
    rng = np.random.default_rng(21)
    yy, xx = np.mgrid[0:H, 0:W]
    mask = ((yy - 250) / 230.0) ** 2 + ((xx - 260) / 140.0) ** 2 < 1.0
    pix = np.flatnonzero(mask.reshape(-1)).astype(np.int32)
    P = len(pix)
    rgb = rng.random((P, 3)).astype(np.float32)
    z = (3.0 + 0.3 * np.sin(xx / 17.0) * np.cos(yy / 23.0)).astype(np.float32)
    pts = np.stack([(xx - 256) / 540.0 * z, (yy - 256) / 540.0 * z, z], -1).astype(np.float32).reshape(-1, 3)[pix]
    pts[rng.random(P) < 0.2] = 0.0                                        # rays without a surface
    fi = _fi()
    pp, pn = fi.assemble(_t(rgb), _t(pts), _t(pix), H, W)
    pp2, pn2 = fi.assemble(_t(rgb), _t(pts), _t(pix), H, W)
    assert torch.equal(pp, pp2) and torch.equal(pn, pn2)
    o_pp, o_pn = io.frame_images(rgb, pts, pix, H, W)
    pp, pn = pp.cpu().numpy(), pn.cpu().numpy()
    assert np.array_equal(pp, o_pp) and np.array_equal(pn, o_pn)
    assert np.array_equal(pp.reshape(-1, 3)[pix], rgb)                    # round trip
    assert not pp.reshape(-1, 3)[~mask.reshape(-1)].any()                 # background untouched
    assert pn.min() >= 0.0 and pn.max() <= 1.0 and np.isfinite(pn).all()


def test_psnr_large_and_reproducible():
    """786 432 floats (a 512 x 512 ray list): against float64 numpy, and bit-identical across runs (fixed reduction tree)."""
    fi = _fi()
    rng = np.random.default_rng(11)
    a = rng.random((512 * 512, 3)).astype(np.float32)
    b = np.clip(a + 0.03 * rng.standard_normal(a.shape), 0, 1).astype(np.float32)
    ta, tb = _t(a), _t(b)
    r1 = fi.psnr_device(ta, tb).clone(); r2 = fi.psnr_device(ta, tb).clone()
    torch.cuda.synchronize()
    assert torch.equal(r1, r2)
    d = (a - b).astype(np.float32)
    mse64 = float(np.mean((d * d).astype(np.float64)))
    assert abs(r1[0].item() - mse64) <= 1e-7 * mse64 + 1e-12
    assert abs(r1[1].item() + 10 * np.log10(mse64)) <= 1e-5


def test_ssim_matches_oracle(light=False):
    """arah_ssim against the oracle's restatement of skimage's structural_similarity (float64: 1e-10), the bounding rectangle exactly,
    a crop smaller than the window -> ValueError like skimage.  (`light`: the subset the CPU emulator run executes.)"""
    from oracle import images_oracle as io
    fi = _fi()
    rng = np.random.default_rng(4)
    H, W = 72, 90
    a = rng.random((H, W, 3)).astype(np.float32)
    b = np.clip(a + 0.1 * rng.standard_normal(a.shape), 0, 1).astype(np.float32)
    mask = np.zeros((H, W), bool)
    mask[7:61, 11:80] = rng.random((54, 69)) < 0.6
    mask[7, 30] = mask[60, 79] = mask[33, 11] = True
    out = fi.ssim_device(_t(a), _t(b), _t(mask)).cpu().numpy()
    assert tuple(int(v) for v in out[1:]) == io.bounding_rect(mask)
    assert abs(out[0] - io.ssim_metric(a, b, mask)) <= 1e-10
    small = np.zeros((H, W), bool); small[10:14, 10:40] = True
    with pytest.raises(ValueError):
        fi.ssim(_t(a), _t(b), _t(small))
    if light:
        return
    assert abs(fi.ssim(_t(a), _t(a), _t(mask)) - 1.0) <= 1e-12
    assert torch.equal(fi.ssim_device(_t(a), _t(b), _t(mask)), fi.ssim_device(_t(a), _t(b), _t(mask)))       # fixed reduction trees
    full = np.ones((H, W), bool)
    assert abs(fi.ssim(_t(a), _t(b), _t(full)) - io.ssim_metric(a, b, full)) <= 1e-10
    with pytest.raises(ValueError):
        fi.ssim(_t(a), _t(b), _t(np.zeros((H, W), bool)))
    # BASELINE-size images (512 x 512), an elliptic bounding-box mask
    yy, xx = np.mgrid[0:512, 0:512]
    big_mask = ((yy - 250) / 230.0) ** 2 + ((xx - 260) / 140.0) ** 2 < 1.0
    A = rng.random((512, 512, 3)).astype(np.float32)
    B = np.clip(A + 0.05 * rng.standard_normal(A.shape), 0, 1).astype(np.float32)
    assert abs(fi.ssim(_t(A), _t(B), _t(big_mask)) - io.ssim_metric(A, B, big_mask)) <= 1e-10


def test_validation_tail_on_golden_batch(seed=2):
    """`FrameImages.validation_tail` on the batch / model outputs the unmodified `validation_step` was run on."""
    from oracle import images_oracle as io
    fi = _fi()
    g = load_images_golden(seed)
    Hg, Wg, P = int(g['H']), int(g['W']), len(g['pix'])
    outputs = {'rgb_values': _t(g['rgb']).view(1, P, 3), 'points_cam': _t(g['points_cam']).view(1, P, 3)}
    batch = {'inputs.img_height': torch.tensor([Hg]), 'inputs.img_width': torch.tensor([Wg]), 'inputs.image_mask': _t(g['mask']).view(1, Hg, Wg),
             'inputs': _t(g['gt']).view(1, P, 3)}
    ev = fi.validation_tail(outputs, batch)
    assert abs(ev['psnr'] - float(g['ref.psnr'])) <= 1e-5
    assert abs(ev['ssim'] - io.ssim_metric(g['ref.rgb_pred'], g['ref.rgb_gt'], g['mask'])) <= 1e-10
    assert np.array_equal(ev['rgb_pred'].permute(1, 2, 0).cpu().numpy(), g['ref.rgb_pred'])
    assert np.abs(ev['normal_pred'].permute(1, 2, 0).cpu().numpy() - g['ref.normal_pred']).max() <= 1.2e-7


CASES = [('hand', 48, 64), ('hand', 64, 40), ('torus', 72, 72), ('two_spheres', 64, 96), ('sphere', 33, 130)]


@pytest.mark.parametrize('mesh,H,W', CASES)
def test_rasterize_bit_exact_against_oracle(mesh, H, W):
    from arah_release_b200 import images
    from oracle import images_oracle as io
    if mesh == 'hand':
        v, f = handmade_mesh()
        R, T, K = np.eye(3, dtype=np.float32), np.zeros(3, np.float32), make_camera(H, W)[2]
    else:
        v, f = iso_mesh(mesh, 20)
        R, T, K = make_camera(H, W, shift=(1.5, -2.25))
        T = T + np.array([0, 0, 2.6], np.float32)
    fi = _fi()
    cams = [(images.opencv_camera(R, T, K, H, W), io.opencv_camera(R, T, K, H, W))]
    for az in (0.0, 180.0):
        cams.append((images.fov_perspective_camera(*images.look_at_view_transform(2.0, 0.0, az)), io.fov_camera(*io.look_at_view_transform(2.0, 0.0, az))))
    for cam, ocam in cams:
        p2f, zb = fi.rasterize(_t(v), _t(f), cam, H, W, zbuf=True)
        o_p2f, o_zb = io.rasterize(io.project(v, ocam), f, H, W)
        assert np.array_equal(p2f.cpu().numpy(), o_p2f)
        assert np.array_equal(zb.cpu().numpy(), o_zb)
        img = fi.normal_image(_t(v), _t(f), p2f, -1.0, R, -1.0)
        assert np.array_equal(img.cpu().numpy(), io.normal_image(v, f, o_p2f, -1.0, R, -1.0))
        img = fi.normal_image(_t(v), _t(f), p2f, 1.0, None, 0.0)
        assert np.array_equal(img.cpu().numpy(), io.normal_image(v, f, o_p2f, 1.0, None, 0.0))


def test_rasterize_degenerate_inputs():
    """No faces / no vertices -> all background; out-of-range vertex indices and non-finite vertices are skipped, not read."""
    from arah_release_b200 import images
    fi = _fi()
    cam = images.fov_perspective_camera(*images.look_at_view_transform(2.0, 0.0, 0.0))
    p2f = fi.rasterize(torch.empty(0, 3, device=DEV), torch.empty(0, 3, dtype=torch.int32, device=DEV), cam, 8, 8)
    assert (p2f == -1).all()
    v = np.array([[-0.5, -0.5, 0], [0.5, -0.5, 0], [0, 0.5, 0], [np.nan, 0, 0], [np.inf, 0, 0]], np.float32)
    f = np.array([[0, 1, 7], [0, 1, 3], [0, 4, 2], [-1, 1, 2], [0, 1, 2]], np.int32)
    p2f = fi.rasterize(_t(v), _t(f), cam, 16, 16).cpu().numpy()
    assert set(np.unique(p2f)) == {-1, 4}
    img = fi.normal_image(_t(v), _t(f), torch.from_numpy(p2f).to(DEV), 1.0, None, 0.0).cpu().numpy()
    assert np.isfinite(img).all() and img.min() >= 0 and img.max() <= 1


def test_normal_maps_full_size_512():
    """BASELINE-size normal maps (512 x 512, the reference's fixed raster size, models/__init__.py:242-244) of a 128^3-lattice
    mesh: the three images against the oracle, plus size-independent properties."""
    from arah_release_b200 import images
    from oracle import images_oracle as io
    v, f = iso_mesh('two_spheres', 128)
    assert f.shape[0] > 20000
    H = W = 512
    R, T, K = make_camera(H, W)
    T = T + np.array([0, 0, 2.6], np.float32)
    posed = (v @ np.array([[0.9, 0.1, 0], [-0.1, 0.9, 0.05], [0, -0.05, 1.0]], np.float32) + np.float32(0.03)).astype(np.float32)
    fi = _fi()
    maps = fi.normal_maps(_t(v), _t(f), _t(posed), R, T, K, H, W)
    maps2 = fi.normal_maps(_t(v), _t(f), _t(posed), R, T, K, H, W)
    ref = io.normal_maps(v, f, posed, R, T, K, H, W)
    for k in ('output_normal', 'normal_cano_front', 'normal_cano_back'):
        assert maps[k].shape == (1, H, W, 3)
        assert torch.equal(maps[k], maps2[k])                                   # atomics, but an order-independent minimum
        assert np.array_equal(maps[k][0].cpu().numpy(), ref[k])
    # face order does not matter (up to the index relabelling) where depths are distinct
    cam = images.opencv_camera(R, T, K, H, W)
    p2f, zb = fi.rasterize(_t(posed), _t(f), cam, H, W, zbuf=True)
    perm = np.random.default_rng(0).permutation(f.shape[0])
    p2f_p, zb_p = fi.rasterize(_t(posed), _t(f[perm]), cam, H, W, zbuf=True)
    assert torch.equal(zb, zb_p)
    back = torch.from_numpy(perm.astype(np.int32)).to(DEV)
    fg = p2f >= 0
    same = back[p2f_p[fg].long()] == p2f[fg]
    assert same.float().mean().item() > 0.999                                   # exact depth ties (shared edges) may relabel
    assert fg.float().mean().item() > 0.03


def test_gen_cano_mesh_end_to_end():
    """extract_canonical_mesh -> skinned vertices -> the three normal maps, all on the GPU (models/__init__.py:203-311)."""
    from helpers import load_golden
    from tools import ref_layout as rl
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    from oracle import images_oracle as io
    fr, _, _ = load_golden('zju377_24x24_s0')
    dev, rend, skin, sdf = rl.modules_from_frame(fr, DEV)
    net = IDHRNetwork(dev, rend, skin, BodyRayTracing(n_steps=fr.n_steps), cano_view_dirs=fr.cano_view_dirs).eval()
    inputs = rl.inputs_from_frame(fr, sdf, DEV)
    verts, faces, posed = net.extract_canonical_mesh(inputs, N=64)
    assert faces.shape[0] > 1000
    R, T, K = make_camera(128, 128)
    T = np.array([0, 0, 3.0], np.float32) - R @ np.asarray(fr.trans, np.float32).reshape(3)
    inputs.update({'cam_rot': _t(R).view(1, 3, 3), 'cam_trans': _t(T).view(1, 3), 'intrinsics': _t(K).view(1, 3, 3)})
    maps = net.render_normal_maps(inputs, image_size=(128, 128), mesh=(verts, faces, posed))
    ref = io.normal_maps(verts.cpu().numpy(), faces.cpu().numpy(), posed.cpu().numpy(), R, T, K, 128, 128)
    for k in ref:
        assert np.array_equal(maps[k][0].cpu().numpy(), ref[k])
    assert (maps['normal_cano_front'][0] != 0.5).any() and (maps['output_normal'][0] != 0).any()
