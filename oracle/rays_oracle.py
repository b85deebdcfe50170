"""TEST INFRASTRUCTURE — numpy restatement of the per-frame ray set-up of the reference's datasets.

Follows /root/reference/im2mesh:
  data/zju_mocap_odp.py:250-283   posed minimally-clothed SMPL vertices (pose blend shapes + LBS + trans)       -> pose_smpl
  data/zju_mocap_odp.py:285-293   bounding box of the posed body (+ box_margin), its 2-D mask                    -> bound_mask
  utils/utils.py:17-52            project / get_bound_corners / get_bound_2d_mask (six cv2.fillPoly calls)       -> bound_mask
  data/zju_mocap_odp.py:137-178   pixel grid, normalize_vectors, get_camera_location, get_camera_rays           -> gen_rays
  data/zju_mocap_odp.py:295-315   uv = homo_2d . K_inv^T, rays = normalise(uv . R), near / far, mask_at_box      -> gen_rays
  utils/utils.py:54-73            get_near_far                                                                   -> near_far

cv2.fillPoly is third-party (OpenCV 4.13 in this image; not part of /root/reference).  `fill_poly` below restates its
published algorithm (modules/imgproc/src/drawing.cpp: CollectPolyEdges + FillEdgeCollection + Line/clipLine, 16.16 fixed
point, 8-connected Bresenham drawn left to right, spans ceil(x_left) .. floor(x_right), edges of partially visible
polygons re-derived from their clipped end points) and is PINNED EMPIRICALLY against cv2 itself in tests/test_rays_oracle.py
(cv2 is importable wherever the tests run) and against the reference's own get_bound_2d_mask in the golden fixtures
(oracle/gen_golden_rays.py): identical on every box that projects inside the image, 1 differing frame in 1500 when corners
leave the image (a degenerate self-overlapping sliver).  Only tests/ and bench.py's CPU leg may import this module.
"""
import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT
F32 = np.float32
FACES = ([0, 1, 3, 2, 0], [4, 5, 7, 6, 5], [0, 1, 5, 4, 0], [2, 3, 7, 6, 2], [0, 2, 6, 4, 0], [1, 3, 7, 5, 1])   # utils/utils.py:47-52 (sic)


# ------------------------------------------------------------------------------------------------ cv2.fillPoly
def _trunc(v):
    return int(v)            # C cast double -> int64: towards zero


def _cdiv(a, b):             # C++ integer division: towards zero
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def clip_line(W, H, x1, y1, x2, y2):
    """cv::clipLine(Size, Point&, Point&) — returns (visible, x1, y1, x2, y2); the points are modified even when invisible."""
    right, bottom = W - 1, H - 1
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += _trunc(float(a - y1) * (x2 - x1) / (y2 - y1)); y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += _trunc(float(a - y2) * (x2 - x1) / (y2 - y1)); y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += _trunc(float(a - x1) * (y2 - y1) / (x2 - x1)); x1 = a; c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += _trunc(float(a - x2) * (y2 - y1) / (x2 - x1)); x2 = a; c2 = 0
    return (c1 | c2) == 0, x1, y1, x2, y2


def draw_line(img, p0, p1):
    """cv::line, thickness 1, LINE_8: clip, then Bresenham from the left end point (initial error adx - 2 ady, step when < 0)."""
    H, W = img.shape
    ok, x0, y0, x1, y1 = clip_line(W, H, int(p0[0]), int(p0[1]), int(p1[0]), int(p1[1]))
    if not ok:
        return
    dx, dy = x1 - x0, y1 - y0
    if dx < 0:
        x0, y0, x1, y1, dx, dy = x1, y1, x0, y0, -dx, -dy
    adx, ady, sy = abs(dx), abs(dy), (1 if dy >= 0 else -1)
    x, y = x0, y0
    if adx >= ady:
        err = adx - 2 * ady
        for _ in range(adx + 1):
            img[y, x] = 1
            if err < 0:
                err += 2 * adx - 2 * ady; y += sy
            else:
                err -= 2 * ady
            x += 1
    else:
        err = ady - 2 * adx
        for _ in range(ady + 1):
            img[y, x] = 1
            if err < 0:
                err += 2 * ady - 2 * adx; x += 1
            else:
                err -= 2 * adx
            y += sy


def poly_edges(pts, W, H):
    """CollectPolyEdges: [(y0, y1, x_fixed_at_y0, dx_fixed)] of the non-horizontal edges (closing edge first)."""
    edges = []
    p0 = pts[-1]
    for p1 in pts:
        x0, y0, x1, y1 = int(p0[0]), int(p0[1]), int(p1[0]), int(p1[1])
        c0x, c0y, c1x, c1y = x0 << XY_SHIFT, y0, x1 << XY_SHIFT, y1
        if not (0 <= x0 < W and 0 <= x1 < W and 0 <= y0 < H and 0 <= y1 < H):
            _, tx0, ty0, tx1, ty1 = clip_line(W, H, x0, y0, x1, y1)       # "use clipped endpoints to create a more accurate PolyEdge"
            if ty0 != ty1:
                c0y, c1y, c0x, c1x = ty0, ty1, tx0 << XY_SHIFT, tx1 << XY_SHIFT
        if y0 != y1:
            dx = _cdiv(c1x - c0x, c1y - c0y)
            if y0 < y1:
                edges.append((y0, y1, c0x + (y0 - c0y) * dx, dx))
            else:
                edges.append((y1, y0, c1x + (y1 - c1y) * dx, dx))
        p0 = p1
    return edges


def fill_poly(img, pts):
    """cv2.fillPoly(img, [pts], 1) for one polygon with integer vertices (uint8 image, in place)."""
    H, W = img.shape
    p0 = pts[-1]
    for p1 in pts:
        draw_line(img, p0, p1)
        p0 = p1
    edges = poly_edges(pts, W, H)
    if not edges:
        return
    ymin, ymax = min(e[0] for e in edges), min(max(e[1] for e in edges), H)
    for y in range(max(ymin, 0), ymax):
        xs = sorted(e[2] + (y - e[0]) * e[3] for e in edges if e[0] <= y < e[1])
        for k in range(0, len(xs) - 1, 2):
            xa, xb = (xs[k] + XY_ONE - 1) >> XY_SHIFT, xs[k + 1] >> XY_SHIFT
            if xa < W and xb >= 0:
                xa, xb = max(xa, 0), min(xb, W - 1)
                if xb >= xa:
                    img[y, xa:xb + 1] = 1


# ------------------------------------------------------------------------------------------------ utils/utils.py
def bound_corners(bounds):
    (mnx, mny, mnz), (mxx, mxy, mxz) = bounds[0], bounds[1]
    return np.array([[mnx, mny, mnz], [mnx, mny, mxz], [mnx, mxy, mnz], [mnx, mxy, mxz],
                     [mxx, mny, mnz], [mxx, mny, mxz], [mxx, mxy, mnz], [mxx, mxy, mxz]])        # :28-39


def project(xyz, K, RT):
    xyz = np.dot(xyz, RT[:, :3].T) + RT[:, 3:].T                                                   # :23-26
    xyz = np.dot(xyz, K.T)
    return xyz[:, :2] / xyz[:, 2:]


def corners_2d(bounds, K, pose):
    return np.round(project(bound_corners(bounds), K, pose)).astype(int)                           # :43-45


def bound_mask(bounds, K, pose, H, W):
    c2 = corners_2d(bounds, K, pose)
    mask = np.zeros((H, W), np.uint8)
    for idx in FACES:
        fill_poly(mask, c2[idx].tolist())
    return mask


def near_far(bounds, ray_o, ray_d):
    norm_d = np.linalg.norm(ray_d, axis=-1, keepdims=True)                                         # :54-73
    viewdir = ray_d / norm_d
    viewdir[(viewdir < 1e-5) & (viewdir > -1e-10)] = 1e-5
    viewdir[(viewdir > -1e-5) & (viewdir < 1e-10)] = -1e-5
    tmin = (bounds[:1] - ray_o[:1]) / viewdir
    tmax = (bounds[1:2] - ray_o[:1]) / viewdir
    t1, t2 = np.minimum(tmin, tmax), np.maximum(tmin, tmax)
    near, far = np.max(t1, axis=-1), np.min(t2, axis=-1)
    hit = near < far
    return near / norm_d[..., 0], far / norm_d[..., 0], hit


# ------------------------------------------------------------------------------------------------ data/zju_mocap_odp.py
def gen_rays(K, R, T, bounds, H, W, mask=None):
    """zju_mocap_odp.py:216,231,286-315 — K (3x3, already rescaled), R (3x3), T (3,) float32; bounds [2,3] float32.
    Returns dict(pix [P] int32 (y*W+x, row-major order), ray_dirs [P,3], near_far [P,2], image_mask [H,W] bool, cam_loc [3])."""
    K, R, T = np.asarray(K, F32), np.asarray(R, F32), np.asarray(T, F32).ravel()
    cam_loc = np.dot(-R.T, T)                                                                      # :171-173
    K_inv = np.linalg.inv(K)                                                                       # :231
    if mask is None:
        mask = bound_mask(bounds, K, np.concatenate([R, T.reshape(3, 1)], axis=-1), H, W)          # :291
    y_inds, x_inds = np.where(mask != 0)                                                           # :292
    Y, X = np.meshgrid(np.arange(H, dtype=F32), np.arange(W, dtype=F32), indexing='ij')           # :137-151
    homo = np.stack([X, Y, np.ones_like(X)], axis=-1)
    uv = np.dot(homo[y_inds, x_inds].reshape(-1, 3), K_inv.T)                                      # :298
    rays = np.dot(uv, R)                                                                           # :175-178
    rays = rays / (np.linalg.norm(rays, ord=2, axis=1, keepdims=True) + 1e-12)                     # :165-169
    near, far, hit = near_far(np.asarray(bounds, F32), np.broadcast_to(cam_loc, rays.shape), rays)  # :302
    image_mask = np.zeros((H, W), bool)
    image_mask[y_inds[hit], x_inds[hit]] = True                                                    # :314-315
    return {'pix': (y_inds[hit] * W + x_inds[hit]).astype(np.int32), 'ray_dirs': rays[hit].astype(F32),
            'near_far': np.stack([near[hit], far[hit]], axis=-1).astype(F32), 'image_mask': image_mask,
            'cam_loc': cam_loc.astype(F32), 'K_inv': K_inv.astype(F32), 'bound_mask': mask}


def pose_smpl(minimal_shape, posedirs, pose_feature, skinning_weights, bone_transforms, trans, box_margin=0.05):
    """zju_mocap_odp.py:268-289: pose blend shapes, LBS, translation; bounds of the posed body with margin.
    Dtypes as the reference sees them with float32 model files: the pose feature is float64 (scipy), so the blend-shape product
    is float64 and added into the float32 shape; everything after is float32."""
    ms = np.asarray(minimal_shape, F32).copy()
    pf = np.asarray(pose_feature, np.float64).reshape(207, 1)
    ms += np.dot(np.asarray(posedirs, F32).reshape(-1, 207), pf).reshape(-1, 3)                    # :270-272
    w = np.asarray(skinning_weights, F32)
    T = np.dot(w, np.asarray(bone_transforms, F32).reshape(-1, 16)).reshape(-1, 4, 4)              # :276
    homo = np.concatenate([ms, np.ones((ms.shape[0], 1), F32)], axis=-1).reshape(-1, 4, 1)
    verts = (np.matmul(T, homo)[:, :3, 0].astype(F32) + np.asarray(trans, F32)).astype(F32)        # :278-280
    mn, mx = verts.min(0) - F32(box_margin), verts.max(0) + F32(box_margin)                        # :286-289
    return verts, np.stack([mn, mx], axis=0).astype(F32)
