"""Shared inputs for the image-tail tests (CPU oracle, host harness, GPU): seeded meshes + cameras, and a rasteriser-independent
float64 ray caster used to cross-check `pix_to_face`."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_images_golden(seed):
    g = np.load(os.path.join(GOLDEN, f'images_s{seed}.npz'))
    return {k: g[k] for k in g.files}


def iso_mesh(name, N):
    """Closed small-triangle mesh of an analytic body (oracle marching cubes: test infrastructure)."""
    from helpers_mesh import analytic_volumes
    from oracle import oracle as orc
    v, f = orc.marching_cubes(analytic_volumes(N)[name][0])
    return np.ascontiguousarray(v, np.float32), np.ascontiguousarray(f, np.int32)


def handmade_mesh():
    """Few large triangles: two overlapping quads at different depths, a triangle leaving the image, a zero-area face, a face
    behind the camera (camera looks down +z from the origin in the test camera below), and a duplicate face (depth tie)."""
    v = np.array([[-0.5, -0.5, 3.0], [0.5, -0.5, 3.0], [0.5, 0.5, 3.0], [-0.5, 0.5, 3.0],             # far quad
                  [-0.2, -0.3, 2.0], [0.6, -0.3, 2.2], [0.6, 0.4, 2.4], [-0.2, 0.4, 2.1],              # nearer, slanted quad
                  [0.3, 0.1, 2.5], [4.0, 0.2, 2.5], [0.3, 3.0, 2.5],                                   # leaves the image
                  [0.0, 0.0, 1.5], [0.1, 0.1, 1.5], [0.2, 0.2, 1.5],                                   # zero area
                  [-0.1, -0.1, -1.0], [0.1, -0.1, -1.0], [0.0, 0.1, -1.0],                             # behind the camera
                  [-0.6, -0.6, 0.5], [-0.4, -0.6, -0.5], [-0.5, -0.4, 0.5]], np.float32)               # crosses the camera plane
    f = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [8, 9, 10], [11, 12, 13], [14, 15, 16], [17, 18, 19],
                  [0, 1, 2], [6, 5, 4]], np.int32)                                                      # duplicate + reversed winding
    return v, f


def make_camera(H, W, focal=None, shift=(0.0, 0.0)):
    """OpenCV camera (cam_rot, cam_trans, K) looking down +z with a small rotation; world == camera up to that."""
    focal = 0.9 * max(H, W) if focal is None else focal
    a, b = 0.05, -0.08
    Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
    Ry = np.array([[np.cos(b), 0, np.sin(b)], [0, 1, 0], [-np.sin(b), 0, np.cos(b)]])
    R = (Rx @ Ry).astype(np.float32)
    T = np.array([0.02, -0.03, 0.1], np.float32)
    K = np.array([[focal, 0, W / 2 + shift[0]], [0, focal, H / 2 + shift[1]], [0, 0, 1]], np.float32)
    return R, T, K


def raycast_pix_to_face(verts, faces, cam_rot, cam_trans, K, H, W):
    """Nearest triangle hit by the ray through every pixel centre (u, v) = (x + 0.5, y + 0.5) of an OpenCV pinhole camera —
    Moeller-Trumbore in float64, no projection, no barycentric rasterisation.  -> face [H,W] (-1 none), margin [H,W] = smallest
    barycentric coordinate of the hit (how far inside the triangle: small = near an edge), depth gap to the runner-up."""
    v = np.asarray(verts, np.float64) @ np.asarray(cam_rot, np.float64).T + np.asarray(cam_trans, np.float64)
    f = np.asarray(faces, np.int64)
    K = np.asarray(K, np.float64)
    ys, xs = np.mgrid[0:H, 0:W]
    d = np.stack([(xs + 0.5 - K[0, 2]) / K[0, 0], (ys + 0.5 - K[1, 2]) / K[1, 1], np.ones_like(xs, np.float64)], -1).reshape(-1, 3)
    best_t = np.full(d.shape[0], np.inf); second_t = np.full(d.shape[0], np.inf)
    best_f = np.full(d.shape[0], -1, np.int64); margin = np.zeros(d.shape[0])
    for i, (a, b, c) in enumerate(f):
        p0, e1, e2 = v[a], v[b] - v[a], v[c] - v[a]
        if min(v[a][2], v[b][2], v[c][2]) <= 0:
            continue
        pv = np.cross(d, e2)
        det = pv @ e1
        with np.errstate(divide='ignore', invalid='ignore'):
            inv = 1.0 / det
            tv = -p0
            u = (pv @ tv) * inv
            qv = np.cross(tv, e1)
            w = (d @ qv) * inv
            t = (e2 @ qv) * inv
        hit = (np.abs(det) > 1e-14) & (u > 0) & (w > 0) & (u + w < 1) & (t > 0)
        better = hit & (t < best_t)
        second_t = np.where(better, best_t, np.where(hit & (t < second_t), t, second_t))
        margin = np.where(better, np.minimum(np.minimum(u, w), 1 - u - w), margin)
        best_f = np.where(better, i, best_f)
        best_t = np.where(better, t, best_t)
    gap = second_t - best_t
    return best_f.reshape(H, W), margin.reshape(H, W), gap.reshape(H, W)
