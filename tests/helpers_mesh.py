"""Mesh checks that depend on neither the CUDA kernels nor the oracle's triangulation code (numpy only)."""
import numpy as np


def mesh_report(verts, faces):
    """Closedness / orientation / topology of a triangle mesh: every undirected edge must be used by exactly two triangles
    and every directed edge exactly once (consistent winding)."""
    f = np.asarray(faces, np.int64)
    a = np.concatenate([f[:, 0], f[:, 1], f[:, 2]])
    b = np.concatenate([f[:, 1], f[:, 2], f[:, 0]])
    nv = int(verts.shape[0])
    und = np.minimum(a, b) * nv + np.maximum(a, b)
    _, cu = np.unique(und, return_counts=True)
    _, cd = np.unique(a * nv + b, return_counts=True)
    v = np.asarray(verts, np.float64)
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    vol = float(np.einsum('ij,ij->i', p0, np.cross(p1, p2)).sum() / 6.0)
    return {'n_verts': nv, 'n_faces': int(f.shape[0]), 'open_or_nonmanifold_edges': int((cu != 2).sum()),
            'repeated_directed_edges': int((cd != 1).sum()), 'euler': nv - int(cu.shape[0]) + int(f.shape[0]),
            'signed_volume': vol, 'unreferenced_verts': nv - int(np.unique(f).shape[0]),
            'degenerate_faces': int(((f[:, 0] == f[:, 1]) | (f[:, 1] == f[:, 2]) | (f[:, 0] == f[:, 2])).sum())}


def analytic_volumes(N):
    """[-1,1]^3 lattice test bodies: name -> (sdf [N,N,N] float32, enclosed volume, euler characteristic)."""
    ax = np.linspace(-1.0, 1.0, N)
    x, y, z = np.meshgrid(ax, ax, ax, indexing='ij')
    sph = np.sqrt(x * x + y * y + z * z) - 0.61
    q = np.sqrt(x * x + y * y) - 0.55
    tor = np.sqrt(q * q + z * z) - 0.21
    two = np.minimum(np.sqrt((x - 0.4) ** 2 + y * y + z * z) - 0.3, np.sqrt((x + 0.4) ** 2 + y * y + z * z) - 0.33)
    return {'sphere': (sph.astype(np.float32), 4 / 3 * np.pi * 0.61 ** 3, 2),
            'torus': (tor.astype(np.float32), 2 * np.pi ** 2 * 0.55 * 0.21 ** 2, 0),
            'two_spheres': (two.astype(np.float32), 4 / 3 * np.pi * (0.3 ** 3 + 0.33 ** 3), 4)}


def noise_volume(N, seed=0):
    """White noise (every ambiguous configuration occurs) with an 'outside' shell so that the surface is closed."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((N, N, N)).astype(np.float32)
    v[[0, -1]] = 1.0; v[:, [0, -1]] = 1.0; v[:, :, [0, -1]] = 1.0
    return v
