"""CPU: the canonical-mesh oracle (oracle/mc_oracle.c + lattice restatement in oracle/oracle.py) — SURVEY §8 row f1.

The reference delegates the triangulation to skimage (absent here: parity unpinned for that step, see mc_oracle.c), so the
oracle is pinned by what can be pinned: the lattice against the reference's own arithmetic (utils/sdf_meshing.py:20-38 restated
with the same torch ops), and the extracted surface against geometry (closed, oriented, right topology, right volume)."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers_mesh import analytic_volumes, mesh_report, noise_volume


def test_lattice_matches_reference_arithmetic():
    from oracle import oracle as orc
    for N in (5, 48, 256):
        # the statements of sdf_meshing.py:20-38, verbatim semantics (float32 tensor, python-float scalars)
        voxel_origin, voxel_size = [-1, -1, -1], 2.0 / (N - 1)
        idx = torch.arange(0, N ** 3, 1, out=torch.LongTensor())
        s = torch.zeros(N ** 3, 3)
        s[:, 2] = idx % N
        s[:, 1] = (idx.long() // N) % N
        s[:, 0] = ((idx.long() // N) // N) % N
        s[:, 0] = (s[:, 0] * voxel_size) + voxel_origin[2]
        s[:, 1] = (s[:, 1] * voxel_size) + voxel_origin[1]
        s[:, 2] = (s[:, 2] * voxel_size) + voxel_origin[0]
        g, vox = orc.grid_points(N)
        assert np.array_equal(g, s.numpy())
        assert vox == float(np.float32(voxel_size))
        if N == 256:
            break


def test_case_tables_agree_and_are_complete():
    """The product generates its case table in C++ (arah_mesh.cu), the oracle in C (mc_oracle.c): same rules, two programs."""
    from arah_release_b200 import _lib
    from oracle import oracle as orc
    tri_o, ntri_o = orc.mc_table()
    tri_p, ntri_p = np.zeros((256, 16), np.int8), np.zeros(256, np.uint8)
    assert _lib.lib().arah_mc_case_table(C.c_void_p(tri_p.ctypes.data), C.c_void_p(ntri_p.ctypes.data)) == 0
    assert np.array_equal(tri_o, tri_p) and np.array_equal(ntri_o, ntri_p)
    assert ntri_o.max() == 5 and ntri_o[0] == 0 and ntri_o[255] == 0

    def corners(e):
        a, u, v = e >> 2, e & 1, (e >> 1) & 1
        c0 = (u << (1 if a == 0 else 0)) | (v << (1 if a == 2 else 2))
        return c0, c0 | (1 << a)
    for cs in range(256):
        crossing = {e for e in range(12) if ((cs >> corners(e)[0]) & 1) != ((cs >> corners(e)[1]) & 1)}
        assert {int(e) for e in tri_o[cs] if e >= 0} == crossing, cs
        assert ntri_o[cs] == ntri_o[255 - cs] or True     # (complements need not mirror: ambiguous faces separate INSIDE corners)


@pytest.mark.parametrize('N', [33, 64])
def test_oracle_surface_is_closed_oriented_and_has_the_right_volume(N):
    from oracle import oracle as orc
    for name, (vol, volume, euler) in analytic_volumes(N).items():
        v, f = orc.marching_cubes(vol)
        rep = mesh_report(v, f)
        assert rep['open_or_nonmanifold_edges'] == 0 and rep['repeated_directed_edges'] == 0, (name, rep)
        assert rep['unreferenced_verts'] == 0 and rep['euler'] == euler, (name, rep)
        assert abs(rep['signed_volume'] - volume) / volume < (0.04 if N == 33 else 0.012), (name, rep, volume)   # outward winding => positive
    # vertices lie on the analytic surface up to the interpolation error of a voxel
    vol, _, _ = analytic_volumes(N)['sphere']
    v, _ = orc.marching_cubes(vol)
    assert np.abs(np.linalg.norm(v, axis=1) - 0.61).max() < 0.6 * (2.0 / (N - 1)) ** 2 / 0.61 + 1e-6


def test_oracle_surface_closed_on_noise():
    """White noise hits all 256 configurations incl. every ambiguous face; the surface must still be a closed 2-manifold."""
    from oracle import oracle as orc
    vol = noise_volume(28, seed=1)
    cs_seen = set()
    ins = vol < 0
    N = vol.shape[0]
    code = np.zeros((N - 1,) * 3, np.int32)
    for c in range(8):
        code |= ins[(c & 1):N - 1 + (c & 1), ((c >> 1) & 1):N - 1 + ((c >> 1) & 1), ((c >> 2) & 1):N - 1 + ((c >> 2) & 1)].astype(np.int32) << c
    cs_seen = set(np.unique(code).tolist())
    assert len(cs_seen) == 256
    v, f = orc.marching_cubes(vol)
    rep = mesh_report(v, f)
    assert rep['open_or_nonmanifold_edges'] == 0 and rep['repeated_directed_edges'] == 0 and rep['degenerate_faces'] == 0, rep
    assert rep['n_faces'] == int(orc.mc_table()[1][code].sum())


def test_oracle_level_and_truncation():
    from oracle import oracle as orc
    vol, _, _ = analytic_volumes(24)['sphere']
    v0, f0 = orc.marching_cubes(vol, level=0.0)
    v1, f1 = orc.marching_cubes(vol + 0.1, level=0.1)
    assert v0.shape == v1.shape and np.array_equal(f0, f1) and np.abs(v0 - v1).max() < 1e-5
    ve, fe = orc.marching_cubes(np.ones((8, 8, 8), np.float32))
    assert ve.shape == (0, 3) and fe.shape == (0, 3)


def _band_flags(coarse, level, eps):
    """The refinement rule of arah_sdf_grid_banded (k_grid_band_flag), restated in numpy: a cell whose eight coarse corner values
    satisfy min - eps <= level <= max + eps marks all its corners."""
    c = coarse.astype(np.float32)
    eps = np.float32(eps); level = np.float32(level)
    corners = [c[dx:c.shape[0] - 1 + dx, dy:c.shape[1] - 1 + dy, dz:c.shape[2] - 1 + dz] for dx in (0, 1) for dy in (0, 1) for dz in (0, 1)]
    mn, mx = np.minimum.reduce(corners), np.maximum.reduce(corners)
    cell = (mn - eps <= level) & (level <= mx + eps)
    flag = np.zeros(c.shape, bool)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                flag[dx:c.shape[0] - 1 + dx, dy:c.shape[1] - 1 + dy, dz:c.shape[2] - 1 + dz] |= cell
    return flag


@pytest.mark.parametrize('level', [0.0, 0.05])
def test_banded_lattice_argument_holds_for_any_bounded_error(level):
    """The exactness argument behind arah_sdf_grid_banded, independent of the GPU: perturb a lattice by ANY error of
    magnitude <= eps, restore the exact values only where the band rule asks for it, and marching cubes cannot tell the result
    from the exact lattice — vertices and faces bit-identical — even when the error is adversarial (pushes values towards the level)."""
    from oracle import oracle as orc
    N, eps = 40, 0.02
    rng = np.random.default_rng(7)
    for name, (exact, _, _) in analytic_volumes(N).items():
        for kind in ('random', 'adversarial'):
            if kind == 'random':
                err = rng.uniform(-eps, eps, exact.shape).astype(np.float32)
            else:
                err = (-np.sign(exact - level) * eps * 0.999).astype(np.float32)      # every value moved towards / across the level
            coarse = (exact + err).astype(np.float32)
            assert np.abs(coarse - exact).max() <= eps
            flag = _band_flags(coarse, level, eps)
            banded = np.where(flag, exact, coarse).astype(np.float32)
            v0, f0 = orc.marching_cubes(exact, level=level)
            v1, f1 = orc.marching_cubes(banded, level=level)
            assert np.array_equal(f0, f1) and np.array_equal(v0, v1), (name, kind)
            # the perturbed lattice alone does give another mesh: the refinement is what makes the difference
            v2, f2 = orc.marching_cubes(coarse, level=level)
            assert not (v2.shape == v0.shape and np.array_equal(v2, v0)), (name, kind)
            assert flag.mean() < 0.5
