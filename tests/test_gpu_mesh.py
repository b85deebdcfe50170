"""GPU parity for the canonical-mesh row: arah_sdf_grid + arah_marching_cubes through the C ABI against the
CPU oracle (oracle/oracle.py::sdf_grid, oracle/mc_oracle.c).  Integer/index work (faces, vertex order) must be bit-exact;
the lattice SDF is floating point: tolerance 2e-5 (fp32 FFMA tiles) / 1e-4 (3xTF32 tensor-core tiles) on raw network output,
written below."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from helpers_mesh import analytic_volumes, mesh_report, noise_volume

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _net(fr, root_mode):
    from tools import ref_layout as rl
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    dev, rend, skin, sdf = rl.modules_from_frame(fr, DEV)
    tracer = BodyRayTracing(n_steps=fr.n_steps, near_surface_vol_samples=fr.near_samples, far_surface_vol_samples=fr.far_samples)
    net = IDHRNetwork(dev, rend, skin, tracer, cano_view_dirs=fr.cano_view_dirs, shade_mode='fp32' if root_mode == 'fp32' else 'tf32',
                      root_mode=root_mode).eval()
    return net, rl.inputs_from_frame(fr, sdf, DEV)


def _renderer():
    from arah_release_b200.renderer import ArahRenderer
    return ArahRenderer(DEV, max_rays=1024)


@pytest.mark.parametrize('root_mode,atol', [('fp32', 2e-5), ('3xtf32', 1e-4)])
def test_sdf_grid_matches_oracle(root_mode, atol):
    from oracle import oracle as orc
    fr, _, _ = load_golden('zju377_24x24_s0')
    net, inputs = _net(fr, root_mode)
    r = net._prepare(inputs)
    for N in (17, 48):                       # 17^3 = 4913: ragged last tile
        g = r.sdf_grid(N)
        torch.cuda.synchronize()
        ref = orc.sdf_grid(fr, N)
        np.testing.assert_allclose(g.cpu().numpy(), ref, atol=atol, rtol=0)
        assert (np.sign(g.cpu().numpy()) != np.sign(ref)).mean() < 2e-3


@pytest.mark.parametrize('N,level', [(17, 0.0), (48, 0.0), (256, 0.0), (128, 0.03)])
def test_banded_lattice_gives_the_full_lattice_mesh(N, level):
    """arah_sdf_grid_banded: fp16 pass + split precision near the level.  Marching cubes must not be able to tell it from the
    full split-precision lattice (bit-identical vertices and faces), the run-time bound check must stay silent, and the refined
    points must carry exactly arah_sdf_grid's values."""
    from arah_release_b200.renderer import BAND_EPS
    fr, _, _ = load_golden('zju377_24x24_s0')
    net, inputs = _net(fr, '3xtf32')
    r = net._prepare(inputs)
    full = r.sdf_grid(N)
    vol, stats = r.sdf_grid_banded(N, level=level)
    torch.cuda.synchronize()
    n_ref, n_bad = (int(v) for v in stats.tolist())
    diff = (vol - full).abs()
    print(f'banded lattice N={N} level={level}: refined {n_ref} of {N ** 3} points ({n_ref / N ** 3:.2%}), '
          f'max |fp16 pass - split precision| = {float(diff.max()):.2e} (eps {BAND_EPS})')
    assert n_bad == 0
    assert float(diff.max()) <= 0.5 * BAND_EPS                 # the assumed bound holds with a factor of two to spare
    assert n_ref == int((diff == 0).sum()) or n_ref <= int((diff == 0).sum())     # refined points are bit-equal to arah_sdf_grid
    near = (full - level).abs() <= 0.25 * BAND_EPS
    assert bool((diff[near] == 0).all())                        # everything close to the level was refined
    if N >= 128:
        assert n_ref < 0.10 * N ** 3
    v0, f0 = r.marching_cubes(full, level=level)
    v1, f1 = r.marching_cubes(vol, level=level)
    torch.cuda.synchronize()
    assert torch.equal(f0, f1) and torch.equal(v0, v1)
    v2, f2 = r.canonical_mesh(N, level=level)
    assert torch.equal(f0, f2) and torch.equal(v0, v2) and r.last_band_stats[1] == 0


def test_banded_lattice_rejects_what_it_cannot_do():
    """Error behaviour of arah_sdf_grid_banded: before set_frame, without the fp16 images (root_mode fp32), bad arguments."""
    from arah_release_b200 import _lib
    r0 = _renderer()
    with pytest.raises(_lib.ArahError):
        r0.sdf_grid_banded(17)                                   # no frame yet
    fr, _, _ = load_golden('zju377_24x24_s0')
    net, inputs = _net(fr, 'fp32')
    r = net._prepare(inputs)
    with pytest.raises(_lib.ArahError):
        r.sdf_grid_banded(17)                                    # fp32 root mode packs no fp16 images
    v, f = r.canonical_mesh(17)                                  # ... and the mirror takes the full-precision lattice by itself
    assert v.shape[1] == 3 and f.shape[1] == 3
    net2, inputs2 = _net(fr, '3xtf32')
    r2 = net2._prepare(inputs2)
    for bad in (dict(N=1), dict(N=2000), dict(N=17, eps=-1.0), dict(N=17, eps=float('nan'))):
        with pytest.raises(_lib.ArahError):
            r2.sdf_grid_banded(**bad)


def test_banded_lattice_bound_check_fires():
    """With eps = 0 the coarse values are (almost) never within the bound: the run-time check must report it."""
    fr, _, _ = load_golden('zju377_24x24_s0')
    net, inputs = _net(fr, '3xtf32')
    r = net._prepare(inputs)
    vol, stats = r.sdf_grid_banded(48, eps=0.0)
    torch.cuda.synchronize()
    n_ref, n_bad = (int(v) for v in stats.tolist())
    assert n_ref > 0 and n_bad > 0.5 * n_ref


@pytest.mark.parametrize('N', [2, 3, 31, 64])
def test_marching_cubes_bit_exact_vs_oracle(N):
    from oracle import oracle as orc
    r = _renderer()
    vols = {'noise': noise_volume(N, seed=N)}
    if N >= 31:
        vols.update({k: v[0] for k, v in analytic_volumes(N).items()})
    for name, vol in vols.items():
        for level in (0.0, 0.05):
            v, f = r.marching_cubes(torch.from_numpy(vol).to(DEV), level=level)
            torch.cuda.synchronize()
            vo, fo = orc.marching_cubes(vol, level=level)
            assert v.shape[0] == vo.shape[0] and f.shape[0] == fo.shape[0], (name, N, v.shape, vo.shape, f.shape, fo.shape)
            assert np.array_equal(f.cpu().numpy(), fo), (name, N)
            assert np.array_equal(v.cpu().numpy(), vo), (name, N, np.abs(v.cpu().numpy() - vo).max())


def test_marching_cubes_empty_and_regrow():
    r = _renderer()
    v, f = r.marching_cubes(torch.ones(16, 16, 16, device=DEV))
    assert v.shape == (0, 3) and f.shape == (0, 3)
    vol = noise_volume(40, seed=3)                       # far more faces than the default buffers hold: exercises the regrow path
    v, f = r.marching_cubes(torch.from_numpy(vol).to(DEV), max_verts=100, max_faces=100)
    rep = mesh_report(v.cpu().numpy(), f.cpu().numpy())
    assert rep['open_or_nonmanifold_edges'] == 0 and rep['repeated_directed_edges'] == 0, rep


def test_canonical_mesh_full_size_256():
    """BASELINE-size lattice (256^3, models/__init__.py:205): size-independent properties + the oracle on the same lattice."""
    from arah_release_b200.renderer import create_mesh_vertices_and_faces
    from oracle import oracle as orc
    fr, _, _ = load_golden('zju377_24x24_s0')
    net, inputs = _net(fr, '3xtf32')
    r = net._prepare(inputs)
    vol = r.sdf_grid(256)
    vol2 = r.sdf_grid(256)
    torch.cuda.synchronize()
    assert torch.equal(vol, vol2)                        # deterministic
    # a 4096-point random subset against the CPU oracle
    rng = np.random.default_rng(0)
    sel = rng.choice(256 ** 3, size=4096, replace=False)
    pts, _ = orc.grid_points(256)
    so, _, _ = orc.sdf(fr, pts[sel], grad=False)
    np.testing.assert_allclose(vol.view(-1)[torch.from_numpy(sel).to(DEV)].cpu().numpy(), so, atol=1e-4, rtol=0)
    v, f = r.marching_cubes(vol)
    torch.cuda.synchronize()
    vh, fh = v.cpu().numpy(), f.cpu().numpy()
    vo, fo = orc.marching_cubes(vol.cpu().numpy())
    assert np.array_equal(fh, fo) and np.array_equal(vh, vo)
    rep = mesh_report(vh, fh)
    print('256^3 canonical mesh', rep)
    assert rep['n_faces'] > 10000 and rep['repeated_directed_edges'] == 0 and rep['degenerate_faces'] == 0
    assert rep['open_or_nonmanifold_edges'] == 0 and rep['signed_volume'] > 0
    assert vh.min() >= -1.0 and vh.max() <= 1.0
    mv, mf = create_mesh_vertices_and_faces(r, N=256)
    assert np.array_equal(mv, vh) and np.array_equal(mf, fh)


def test_extract_canonical_mesh_posed_vertices():
    from oracle import oracle as orc
    fr, _, _ = load_golden('cano_20x20_s1')
    net, inputs = _net(fr, 'fp32')
    verts, faces, bar = net.extract_canonical_mesh(inputs, N=40)
    torch.cuda.synchronize()
    vh = verts.cpu().numpy()
    d = np.float32(fr.coord_max - fr.coord_min)
    hat = (vh / np.float32(2.0) + np.float32(0.5)) * np.float32(1.1) * d + np.float32(fr.coord_min) - np.float32(0.05) * d + fr.center.reshape(1, 3).astype(np.float32)
    _, xb, _ = orc.skin(fr, hat, jac=False)
    np.testing.assert_allclose(bar.cpu().numpy(), xb + fr.trans.reshape(1, 3), atol=5e-5, rtol=0)
    assert faces.shape[0] > 100 and int(faces.max()) == verts.shape[0] - 1
