"""TEST INFRASTRUCTURE — numpy restatement of the image-space tail of the reference's validation / test step.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this; the product never does.

(1) `frame_images`, `psnr_metric`: what `LightningModel.validation_step` does with the renderer's outputs
    (/root/reference/im2mesh/metaavatar_render/lightning_model.py:176-205 — masked_scatter_ of rgb / points_cam into the image,
    finite-difference normal map of the depth image; :218-221 + im2mesh/utils/eval.py:6-9 — PSNR over the ray list).
    PARITY PINNED: tests/golden/images_s*.npz hold the outputs of the unmodified `validation_step` run by
    oracle/gen_golden_images.py (only `ssim_metric` / `lpips_metric`, which need skimage / lpips, are replaced by constants).

(2) `rasterize`, `face_normals`, `normal_image`, `normal_maps`: the three normal maps `MetaAvatarRender.forward(gen_cano_mesh=True)`
    renders from the extracted mesh (/root/reference/im2mesh/metaavatar_render/models/__init__.py:226-311).  The rasteriser
    itself is third party: pytorch3d 0.6.1 (environment.yml:17), ABSENT from /root/reference and from this image — PARITY
    UNPINNED for `pix_to_face`.  What is restated is its published algorithm (pytorch3d/csrc/rasterize_meshes/
    rasterize_meshes.cu `CheckPixelInsideFace`, csrc/utils/geometry_utils.cuh, renderer/mesh/rasterizer.py `transform`,
    renderer/cameras.py `look_at_view_transform` / `FoVPerspectiveCameras`, utils/camera_conversions.py
    `_cameras_from_opencv_projection`) for the settings the reference uses: faces_per_pixel = 1, blur_radius = 0,
    perspective_correct = True (both cameras are perspective), no back-face culling, no z clipping in range.
    The glue around it (sign of the normals, rotation into the camera frame, background, (n + 1) / 2 clip) is reference code.
"""
import numpy as np

F = np.float32
K_EPS = F(1e-8)          # pytorch3d csrc/utils/float_math.cuh kEpsilon


# --------------------------------------------------------------------------------------------------------------------
# (1) validation images — lightning_model.py:176-205
# --------------------------------------------------------------------------------------------------------------------
def scatter_rows(rows, pix, H, W, fill=0.0):
    """`img.masked_scatter_(image_mask.view(1,H,W,1), rows[:image_mask.sum()])` (:178-181, :185-188): the k-th true pixel of the
    mask in row-major order receives row k; `pix[k]` (np.where order, what arah_frame_rays emits) is that pixel."""
    img = np.full((H * W, 3), fill, np.float32)
    img[np.asarray(pix, np.int64)] = np.asarray(rows, np.float32)[:len(pix)]
    return img.reshape(H, W, 3)


def depth_normals(pred_points):
    """:190-205.  pred_points [H,W,3] camera-space surface points (zero where nothing was hit)."""
    p = np.asarray(pred_points, np.float32)
    H, W, _ = p.shape
    xs, ys, zs = p[:, :, 0], p[:, :, 1], p[:, :, 2]
    with np.errstate(divide='ignore', invalid='ignore'):
        zy = (zs[1:, :] - zs[:-1, :]) / (ys[1:, :] - ys[:-1, :])
        zx = (zs[:, 1:] - zs[:, :-1]) / (xs[:, 1:] - xs[:, :-1])
        n = np.zeros((H, W, 3), np.float32)
        n[:-1, :, 1] = -zy
        n[:, :-1, 0] = -zx
        n[:, :, 2] = 1
        nx, ny, nz = n[:, :, 0], n[:, :, 1], n[:, :, 2]
        norm = np.sqrt((nx * nx + ny * ny) + nz * nz)[..., None]          # torch.linalg.norm(dim=-1), fp32
        n = n / norm
    n[np.isnan(n)] = -1
    return np.clip((n + F(1)) / F(2), F(0), F(1)).astype(np.float32)


def frame_images(rgb, points_cam, pix, H, W):
    pred_pixels = scatter_rows(rgb, pix, H, W)
    pred_normals = depth_normals(scatter_rows(points_cam, pix, H, W))
    return pred_pixels, pred_normals


def psnr_metric(img_pred, img_gt):
    """im2mesh/utils/eval.py:6-9 on the float32 ray lists (:218-221)."""
    d = np.asarray(img_pred, np.float32) - np.asarray(img_gt, np.float32)
    mse = np.mean(d ** 2)
    return float(mse), float(-10 * np.log(mse) / np.log(10))


# --------------------------------------------------------------------------------------------------------------------
# (2) cameras — pytorch3d conventions: row vectors, X_view = X_world R + T, +X left, +Y up, +Z into the screen
# --------------------------------------------------------------------------------------------------------------------
def _normalize(v, eps=1e-5):
    v = np.asarray(v, np.float32)
    return v / np.maximum(np.linalg.norm(v).astype(np.float32), F(eps))


def look_at_view_transform(dist, elev_deg, azim_deg):
    """pytorch3d.renderer.cameras.look_at_view_transform(dist, elev, azim) with at = 0, up = +Y (models/__init__.py:265,291)."""
    elev, azim = F(np.pi / 180.0) * F(elev_deg), F(np.pi / 180.0) * F(azim_deg)
    C = np.array([F(dist) * np.cos(elev) * np.sin(azim), F(dist) * np.sin(elev), F(dist) * np.cos(elev) * np.cos(azim)], np.float32)
    z = _normalize(-C)
    x = _normalize(np.cross(np.array([0, 1, 0], np.float32), z))
    y = _normalize(np.cross(z, x))
    if np.allclose(x, 0.0, atol=5e-3):
        x = _normalize(np.cross(y, z))
    R = np.stack([x, y, z], axis=1).astype(np.float32)                  # columns are the camera axes
    T = -(C @ R).astype(np.float32)
    return R, T


def fov_camera(R, T, fov_deg=60.0):
    """FoVPerspectiveCameras(R, T) defaults: fov 60 deg, aspect 1 -> x_ndc = s X / Z, s = 1 / tan(fov / 2)."""
    s = F(1.0) / np.tan(F(np.pi / 180.0) * F(fov_deg) / F(2), dtype=np.float32)
    return {'R': np.asarray(R, np.float32), 'T': np.asarray(T, np.float32), 'fx': F(s), 'fy': F(s), 'px': F(0), 'py': F(0)}


def opencv_camera(cam_rot, cam_trans, K, H, W):
    """pytorch3d.utils.cameras_from_opencv_projection(R, tvec, K, image_size=(H, W)) (models/__init__.py:247-252)."""
    K = np.asarray(K, np.float32).reshape(3, 3)
    scale = F(min(H, W)) / F(2)
    c0 = np.array([F(W) / F(2), F(H) / F(2)], np.float32)
    focal = np.array([K[0, 0], K[1, 1]], np.float32) / scale
    p0 = -(np.array([K[0, 2], K[1, 2]], np.float32) - c0) / scale
    R = np.asarray(cam_rot, np.float32).reshape(3, 3).T.copy()
    T = np.asarray(cam_trans, np.float32).reshape(3).copy()
    R[:, :2] *= -1
    T[:2] *= -1
    return {'R': R, 'T': T, 'fx': focal[0], 'fy': focal[1], 'px': p0[0], 'py': p0[1]}


def project(verts, cam):
    """MeshRasterizer.transform: view = X R + T; ndc xy = ([x y z 1] P)[:2] / z with P the perspective matrix; z stays view z."""
    v = np.asarray(verts, np.float32)
    R, T = cam['R'], cam['T']
    view = np.empty_like(v)
    for c in range(3):
        view[:, c] = ((v[:, 0] * R[0, c] + v[:, 1] * R[1, c]) + v[:, 2] * R[2, c]) + T[c]
    out = np.empty_like(v)
    with np.errstate(divide='ignore', invalid='ignore'):
        out[:, 0] = (cam['fx'] * view[:, 0] + cam['px'] * view[:, 2]) / view[:, 2]
        out[:, 1] = (cam['fy'] * view[:, 1] + cam['py'] * view[:, 2]) / view[:, 2]
    out[:, 2] = view[:, 2]
    return out


def pix_to_ndc(i, S1, S2):
    """PixToNonSquareNdc (rasterization_utils.cuh): centre of pixel i along an axis of S1 pixels, S2 the other axis."""
    rng = F(2) * F(S1) / F(S2) if S1 > S2 else F(2)
    off = rng / F(2)
    return -off + (rng * np.asarray(i, np.float32) + off) / F(S1)


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def rasterize(verts_ndc, faces, H, W):
    """Nearest face per pixel.  -> pix_to_face [H,W] int32 (-1 = background), zbuf [H,W] float32 (-1 = background)."""
    v = np.asarray(verts_ndc, np.float32)
    faces = np.asarray(faces, np.int64)
    xf = pix_to_ndc(W - 1 - np.arange(W), W, H).astype(np.float32)       # pixel column xi -> ndc x (decreasing)
    yf = pix_to_ndc(H - 1 - np.arange(H), H, W).astype(np.float32)
    best_z = np.full((H, W), np.inf, np.float32)
    best_f = np.full((H, W), -1, np.int32)
    for f, (i0, i1, i2) in enumerate(faces):
        x0, y0, z0 = v[i0]; x1, y1, z1 = v[i1]; x2, y2, z2 = v[i2]
        if not (np.isfinite([x0, y0, z0, x1, y1, z1, x2, y2, z2]).all()):
            continue
        if min(z0, z1, z2) < K_EPS:                                      # any vertex at / behind the camera plane: not drawn
            continue
        area = _edge(x0, y0, x1, y1, x2, y2)                             # face_area = EdgeFunction(v0, v1, v2)
        if -K_EPS <= area <= K_EPS:
            continue
        xmin, xmax, ymin, ymax = min(x0, x1, x2), max(x0, x1, x2), min(y0, y1, y2), max(y0, y1, y2)
        cols = np.nonzero((xf >= xmin) & (xf <= xmax))[0]
        rows = np.nonzero((yf >= ymin) & (yf <= ymax))[0]
        if len(cols) == 0 or len(rows) == 0:
            continue
        px, py = np.meshgrid(xf[cols], yf[rows])
        barea = _edge(x2, y2, x0, y0, x1, y1) + K_EPS                    # BarycentricCoordsForward
        w0 = _edge(px, py, x1, y1, x2, y2) / barea
        w1 = _edge(px, py, x2, y2, x0, y0) / barea
        w2 = _edge(px, py, x0, y0, x1, y1) / barea
        inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
        t0, t1, t2 = w0 * z1 * z2, z0 * w1 * z2, z0 * z1 * w2            # BarycentricPerspectiveCorrectionForward
        den = np.maximum((t0 + t1) + t2, K_EPS)
        pz = ((t0 / den) * z0 + (t1 / den) * z1) + (t2 / den) * z2
        ok = inside & (pz >= 0)
        rr, cc = np.nonzero(ok)
        for r, c in zip(rr, cc):
            y, x = rows[r], cols[c]
            if pz[r, c] < best_z[y, x]:                                  # ties keep the lower face index
                best_z[y, x] = pz[r, c]
                best_f[y, x] = f
    zbuf = np.where(best_f >= 0, best_z, F(-1)).astype(np.float32)
    return best_f, zbuf


def face_normals(verts, faces):
    """Meshes.faces_normals_packed(): cross(v1 - v0, v2 - v0) / max(norm, 1e-6)."""
    v = np.asarray(verts, np.float32)
    f = np.asarray(faces, np.int64)
    a, b = v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]
    c = np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1)
    n = np.sqrt((c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]) + c[:, 2] * c[:, 2])
    return (c / np.maximum(n, F(1e-6))[:, None]).astype(np.float32)


def normal_image(verts, faces, pix_to_face, sign=1.0, rot=None, background=0.0):
    """models/__init__.py:256-263 (posed: sign -1, rot = cam_rot, background -1) and :279-286 (canonical: +1, none, 0)."""
    H, W = pix_to_face.shape
    n = F(sign) * face_normals(verts, faces)
    if rot is not None:
        Rm = np.asarray(rot, np.float32).reshape(3, 3)
        n = np.stack([(Rm[i, 0] * n[:, 0] + Rm[i, 1] * n[:, 1]) + Rm[i, 2] * n[:, 2] for i in range(3)], 1)
    img = np.full((H, W, 3), F(background), np.float32)
    fg = pix_to_face >= 0
    img[fg] = n[pix_to_face[fg]]
    return np.clip((img + F(1)) / F(2), F(0), F(1)).astype(np.float32)


def normal_maps(verts_cano, faces, verts_posed, cam_rot, cam_trans, K, H=512, W=512):
    """The three images `forward(gen_cano_mesh=True)` adds to its outputs (models/__init__.py:226-309)."""
    out = {}
    p2f, _ = rasterize(project(verts_posed, opencv_camera(cam_rot, cam_trans, K, H, W)), faces, H, W)
    out['output_normal'] = normal_image(verts_posed, faces, p2f, -1.0, cam_rot, -1.0)
    for name, azim in (('normal_cano_front', 0.0), ('normal_cano_back', 180.0)):
        R, T = look_at_view_transform(2.0, 0.0, azim)
        p2f, _ = rasterize(project(verts_cano, fov_camera(R, T)), faces, H, W)
        out[name] = normal_image(verts_cano, faces, p2f, 1.0, None, 0.0)
    return out


# --------------------------------------------------------------------------------------------------------------------
# (3) SSIM — lightning_model.py:222, im2mesh/utils/eval.py:11-19
# --------------------------------------------------------------------------------------------------------------------
def bounding_rect(mask):
    """cv2.boundingRect(mask.astype(np.uint8)): (x, y, w, h) of the non-zero pixels of a 2-D mask; (0, 0, 0, 0) if there are none."""
    ys, xs = np.nonzero(np.asarray(mask))
    if len(ys) == 0:
        return 0, 0, 0, 0
    return int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)


def structural_similarity(X, Y, win_size=7, data_range=2.0, K1=0.01, K2=0.03):
    """skimage.metrics.structural_similarity (0.18.1, environment.yml; ABSENT here — PARITY UNPINNED against skimage itself) for one
    2-D channel with its defaults as the reference calls it: uniform 7x7 window, sample covariance, float64 arithmetic and — the
    images being float32 without an explicit data_range — data_range = dtype range of float = 2.  Restated from its source with
    the same scipy primitive (scipy.ndimage.uniform_filter, mode 'reflect'); the mean is taken over the image cropped by 3."""
    from scipy.ndimage import uniform_filter
    X, Y = np.asarray(X, np.float64), np.asarray(Y, np.float64)
    if min(X.shape) < win_size:
        raise ValueError('win_size exceeds image extent')
    NP = win_size ** 2
    cov_norm = NP / (NP - 1.0)
    f = lambda a: uniform_filter(a, size=win_size)
    ux, uy, uxx, uyy, uxy = f(X), f(Y), f(X * X), f(Y * Y), f(X * Y)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:-pad, pad:-pad].mean(dtype=np.float64))


def ssim_metric(img_pred, img_gt, mask_at_box):
    """im2mesh/utils/eval.py:11-19: crop both [H,W,3] float32 images to the bounding rectangle of the mask, then
    structural_similarity(multichannel=True) = mean of the per-channel values."""
    x, y, w, h = bounding_rect(mask_at_box)
    a, b = np.asarray(img_pred)[y:y + h, x:x + w], np.asarray(img_gt)[y:y + h, x:x + w]
    return float(np.mean([structural_similarity(a[..., c], b[..., c]) for c in range(a.shape[-1])]))
