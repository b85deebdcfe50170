"""Seeded synthetic "ZJU-377-like" frames for the ARAH hot path (numpy only).

No dataset, SMPL file or checkpoint is reachable in this environment, so the
parity tests, the CPU oracle, the reference harness and bench.py all consume the frames built here:

* a 24-joint capsule-chain body that follows the SMPL kinematic tree
  (``ktree_parents`` — /root/reference/im2mesh/metaavatar/models/siren_modules.py:204-205),
  6890 surface vertices, <=4 non-zero skinning weights per vertex;
* seeded joint rotations composed down the tree -> ``bone_transforms`` (canonical -> posed, without the
  global translation; same meaning as /root/reference/im2mesh/metaavatar_render/lightning_model.py:564);
* a pinhole camera and the bbox rays + near/far bounds in the layout the reference's datasets emit
  (/root/reference/im2mesh/data/zju_mocap_odp.py:286-315, im2mesh/utils/utils.py:56-73);
* network weights in the reference's layouts: FiLM-SIREN SDF (hyperlayers.py:391-415), weight-normed
  skinning MLP (metaavatar/models/decoder.py:133-233), weight-normed colour MLP
  (metaavatar_render/models/decoder.py:10-124), scalar beta (decoder.py:127-133).

The SDF and skinning nets are *fitted* offline to the analytic body (tools/make_synthetic_assets.py) and
shipped as ``data/synthetic_nets_v1.npz``; everything else is generated from the seed.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

F32 = np.float32

KTREE_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21],
                         dtype=np.int32)
N_JOINTS = 24
N_VERTS = 6890

# rough SMPL-like T-pose joint locations (metres): x = left(+)/right(-), y = up, z = front
JOINTS_CANO = np.array([
    [0.00, 0.00, 0.00],    # 0 pelvis
    [0.08, -0.09, 0.00],   # 1 l_hip
    [-0.08, -0.09, 0.00],  # 2 r_hip
    [0.00, 0.11, -0.02],   # 3 spine1
    [0.12, -0.47, 0.00],   # 4 l_knee
    [-0.12, -0.47, 0.00],  # 5 r_knee
    [0.00, 0.25, 0.00],    # 6 spine2
    [0.14, -0.87, -0.03],  # 7 l_ankle
    [-0.14, -0.87, -0.03], # 8 r_ankle
    [0.00, 0.31, 0.02],    # 9 spine3
    [0.15, -0.93, 0.09],   # 10 l_foot
    [-0.15, -0.93, 0.09],  # 11 r_foot
    [0.00, 0.51, -0.01],   # 12 neck
    [0.08, 0.42, 0.00],    # 13 l_collar
    [-0.08, 0.42, 0.00],   # 14 r_collar
    [0.00, 0.60, 0.03],    # 15 head
    [0.18, 0.45, -0.01],   # 16 l_shoulder
    [-0.18, 0.45, -0.01],  # 17 r_shoulder
    [0.44, 0.45, -0.03],   # 18 l_elbow
    [-0.44, 0.45, -0.03],  # 19 r_elbow
    [0.69, 0.45, -0.03],   # 20 l_wrist
    [-0.69, 0.45, -0.03],  # 21 r_wrist
    [0.78, 0.44, -0.03],   # 22 l_hand
    [-0.78, 0.44, -0.03],  # 23 r_hand
], dtype=np.float64)

# capsules: (owning joint, end point A, end point B, radius).  Each capsule moves rigidly with its joint.
def _capsules():
    J = JOINTS_CANO
    caps = []
    def add(j, a, b, r):
        caps.append((j, np.asarray(a, np.float64), np.asarray(b, np.float64), float(r)))
    add(0, J[0] + [0, -0.03, 0], J[3], 0.125)
    add(3, J[3], J[6], 0.125)
    add(6, J[6], J[9], 0.13)
    add(9, J[9], J[12] + [0, -0.06, 0], 0.125)
    add(12, J[12] + [0, -0.04, 0], J[15], 0.05)
    add(15, J[15] + [0, 0.06, 0.0], J[15] + [0, 0.10, 0.01], 0.095)
    for s, (hip, knee, ankle, foot) in zip((1, -1), ((1, 4, 7, 10), (2, 5, 8, 11))):
        add(hip, J[hip], J[knee], 0.072)
        add(knee, J[knee], J[ankle], 0.052)
        add(ankle, J[ankle], J[foot], 0.042)
        add(foot, J[foot], J[foot] + [0, 0, 0.07], 0.038)
    for coll, sh, el, wr, ha in ((13, 16, 18, 20, 22), (14, 17, 19, 21, 23)):
        add(coll, J[coll], J[sh], 0.06)
        add(sh, J[sh], J[el], 0.048)
        add(el, J[el], J[wr], 0.04)
        add(wr, J[wr], J[ha], 0.035)
        add(ha, J[ha], J[ha] + (J[ha] - J[wr]) * 0.9, 0.033)
    return caps

CAPSULES = _capsules()


def capsule_dists(p: np.ndarray) -> np.ndarray:
    """Signed distance of points p [N,3] to every capsule -> [N, n_caps] (float64)."""
    out = np.empty((p.shape[0], len(CAPSULES)), np.float64)
    for i, (_, a, b, r) in enumerate(CAPSULES):
        ab = b - a
        t = np.clip(((p - a) @ ab) / (ab @ ab), 0.0, 1.0)
        c = a + t[:, None] * ab
        out[:, i] = np.linalg.norm(p - c, axis=1) - r
    return out


def body_sdf(p: np.ndarray, k: float = 0.03) -> np.ndarray:
    """Smooth-union SDF of the capsule body (metres, float64).  k = smooth-min radius."""
    d = capsule_dists(p)
    # log-sum-exp smooth min
    m = d.min(axis=1)
    s = np.exp(-(d - m[:, None]) / k).sum(axis=1)
    return m - k * np.log(s)


def body_weights(p: np.ndarray, sharp: float = 0.04) -> np.ndarray:
    """Soft joint weights [N,24] for canonical points (sum to 1, <=4 non-zero)."""
    d = capsule_dists(p)
    own = np.array([c[0] for c in CAPSULES])
    dj = np.full((p.shape[0], N_JOINTS), 1e3)
    for ci, j in enumerate(own):
        dj[:, j] = np.minimum(dj[:, j], d[:, ci])
    dj = np.maximum(dj, 0.0)
    w = np.exp(-(dj - dj.min(axis=1, keepdims=True)) / sharp)
    # keep the 4 largest
    idx = np.argsort(-w, axis=1)[:, 4:]
    np.put_along_axis(w, idx, 0.0, axis=1)
    w /= w.sum(axis=1, keepdims=True)
    return w


def _rodrigues(rv: np.ndarray) -> np.ndarray:
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.eye(3)
    k = rv / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def sample_body_vertices(rng: np.random.Generator, n: int = N_VERTS) -> np.ndarray:
    """n points on the union surface of the capsules (canonical pose)."""
    pts = []
    total = 0
    areas = np.array([2 * np.pi * r * (np.linalg.norm(b - a) + 2 * r) for _, a, b, r in CAPSULES])
    probs = areas / areas.sum()
    while total < n:
        ci = rng.choice(len(CAPSULES), size=4 * n, p=probs)
        u = rng.normal(size=(4 * n, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        t = rng.uniform(-0.15, 1.15, size=4 * n)
        cand = np.empty((4 * n, 3))
        for k in range(len(CAPSULES)):
            m = ci == k
            if not m.any():
                continue
            _, a, b, r = CAPSULES[k]
            ab = b - a
            L = np.linalg.norm(ab)
            ax = ab / L
            tt = t[m]
            uu = u[m]
            # cylinder part: radial direction orthogonal to the axis; caps: hemisphere
            radial = uu - (uu @ ax)[:, None] * ax
            radial /= np.maximum(np.linalg.norm(radial, axis=1, keepdims=True), 1e-9)
            on_cyl = (tt >= 0) & (tt <= 1)
            p_cyl = a + tt[:, None] * ab + r * radial
            hemi = np.where((uu @ ax)[:, None] * np.where(tt < 0, -1, 1)[:, None] < 0, -uu, uu)
            p_cap = np.where((tt < 0)[:, None], a, b) + r * hemi
            cand[m] = np.where(on_cyl[:, None], p_cyl, p_cap)
        d = capsule_dists(cand).min(axis=1)
        keep = cand[np.abs(d) < 1e-6]
        pts.append(keep)
        total += len(keep)
    return np.concatenate(pts, 0)[:n]


def pose_bone_transforms(rng: np.random.Generator, max_angle: float) -> np.ndarray:
    """Seeded joint rotations (|theta| <= max_angle) composed down the tree -> [24,4,4] canonical->posed."""
    G = np.zeros((N_JOINTS, 4, 4))
    for j in range(N_JOINTS):
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        ang = rng.uniform(-max_angle, max_angle) * (0.35 if j in (0, 3, 6, 9) else 1.0)
        R = _rodrigues(axis * ang)
        p = KTREE_PARENTS[j]
        L = np.eye(4)
        L[:3, :3] = R
        L[:3, 3] = JOINTS_CANO[j] - (JOINTS_CANO[p] if p >= 0 else 0.0)
        G[j] = L if p < 0 else G[p] @ L
    out = np.zeros_like(G)
    for j in range(N_JOINTS):
        G0inv = np.eye(4)
        G0inv[:3, 3] = -JOINTS_CANO[j]
        out[j] = G[j] @ G0inv
    return out


def _weight_norm_init(rng, out_dim, in_dim):
    """torch nn.Linear default init + weight_norm parametrisation (g = row norms of v)."""
    bound = 1.0 / np.sqrt(in_dim)
    v = rng.uniform(-bound, bound, size=(out_dim, in_dim)).astype(F32)
    b = rng.uniform(-bound, bound, size=(out_dim,)).astype(F32)
    g = np.linalg.norm(v.astype(np.float64), axis=1, keepdims=True).astype(F32)
    return v, g, b


def color_net_init(rng: np.random.Generator, d_in_total: int = 417, d_hidden: int = 256, skip_layer: int = 3,
                   gain: float = 2.5):
    """Colour MLP of configs/arah-zju/ZJUMOCAP-377_4gpus.yaml:39 (n_layers 5, skips [3]):
    417->256->256->128 ; [417+128]->256->256->3.  ``gain`` scales g so the random net has visible contrast."""
    dims_in = [d_in_total, d_hidden, d_hidden, d_hidden // 2 + d_in_total, d_hidden, d_hidden]
    dims_out = [d_hidden, d_hidden, d_hidden // 2, d_hidden, d_hidden, 3]
    layers = []
    for i, (di, do) in enumerate(zip(dims_in, dims_out)):
        v, g, b = _weight_norm_init(rng, do, di)
        layers.append({'v': v, 'g': (g * gain).astype(F32), 'b': b})
    return layers


_ASSET = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'synthetic_nets_v1.npz')


def load_fitted_nets(path: str = _ASSET):
    """SDF (7 layers + FiLM) and skinning (5 weight-normed layers) nets fitted to the capsule body."""
    z = np.load(path)
    sdf = {
        'W': [z[f'sdf_W{i}'].astype(F32) for i in range(7)],
        'b': [z[f'sdf_b{i}'].astype(F32) for i in range(7)],
        'freq': z['sdf_freq'].astype(F32),     # [6,256]
        'phase': z['sdf_phase'].astype(F32),   # [6,256]
    }
    skin = [{'v': z[f'skin_v{i}'].astype(F32), 'g': z[f'skin_g{i}'].astype(F32), 'b': z[f'skin_b{i}'].astype(F32)}
            for i in range(5)]
    meta = {'coord_min': F32(z['coord_min']), 'coord_max': F32(z['coord_max']), 'center': z['center'].astype(F32)}
    return sdf, skin, meta


@dataclass
class Frame:
    """One frame's worth of hot-path inputs, all float32 numpy, reference layouts (batch dim dropped)."""
    H: int
    W: int
    pix: np.ndarray              # [P] flat pixel index of each ray
    ray_dirs: np.ndarray         # [P,3] unit, world space
    near_far: np.ndarray         # [P,2]
    cam_loc: np.ndarray          # [3]
    pose: np.ndarray             # [4,4] world->camera [R|T]
    K: np.ndarray                # [3,3]
    bone_transforms: np.ndarray  # [24,4,4]
    smpl_verts: np.ndarray       # [6890,3] posed + trans
    smpl_weights: np.ndarray     # [6890,24]
    minimal_shape: np.ndarray    # [6890,3] canonical verts
    trans: np.ndarray            # [3]
    coord_min: np.float32
    coord_max: np.float32
    center: np.ndarray           # [3]
    sdf: dict = field(repr=False, default=None)
    skin: list = field(repr=False, default=None)
    color: list = field(repr=False, default=None)
    latent: np.ndarray = field(repr=False, default=None)  # [128]
    beta: np.float32 = F32(1e-3)
    n_steps: int = 64
    near_samples: int = 16
    far_samples: int = 16
    cano_view_dirs: bool = False
    color_mode: str = 'idr'       # RenderingNetwork mode: 'idr' | 'no_view_dir' (mono configs) | 'no_normal'

    @property
    def P(self):
        return int(self.ray_dirs.shape[0])


def make_frame(H: int = 64, W: int = 64, seed: int = 0, *, max_angle: float = 0.6, fill: float = 0.95,
               beta: float = 5e-3, cano_view_dirs: bool = False, n_steps: int = 64, near_samples: int = 16,
               far_samples: int = 16, frame_idx: int = 0, all_pixels: bool = False, color_mode: str = 'idr') -> Frame:
    """Build one synthetic frame.

    ``fill``: fraction of the image height the posed body's bbox spans (focal length is solved for it).
    ``frame_idx``: frames of one sequence share seed-derived networks/camera and differ in pose (smooth path).
    ``all_pixels``: keep every pixel whose ray hits the bbox (default), no sub-sampling either way.
    """
    rng = np.random.default_rng(seed)
    sdf, skin, meta = load_fitted_nets()
    verts_rng = np.random.default_rng(1234)          # the body itself does not depend on the frame seed
    verts = sample_body_vertices(verts_rng)
    weights = body_weights(verts)

    # pose: seed picks two key poses, frame_idx interpolates (sequence = smooth seeded trajectory, SURVEY §8d cfg 4)
    prng = np.random.default_rng(seed * 7919 + 17)
    s0 = prng.integers(1 << 30)
    bt_a = pose_bone_transforms(np.random.default_rng(s0), max_angle)
    if frame_idx == 0:
        bone_T = bt_a
    else:
        # re-draw with angles modulated by a smooth phase: cheap but deterministic per (seed, frame)
        ph = 0.5 + 0.5 * np.sin(0.07 * frame_idx)
        bone_T = pose_bone_transforms(np.random.default_rng(s0), max_angle * (0.4 + 0.6 * ph))
    trans = np.array([0.05, 0.1, 3.0]) + rng.normal(scale=0.02, size=3)

    T_v = np.einsum('vj,jab->vab', weights, bone_T)
    posed = np.einsum('vab,vb->va', T_v[:, :3, :3], verts) + T_v[:, :3, 3] + trans

    # camera: looks down +z from the origin with a small seeded rotation, pinhole
    R = _rodrigues(rng.normal(scale=0.05, size=3))
    Tc = rng.normal(scale=0.02, size=3)
    cam_loc = -R.T @ Tc
    pc = posed @ R.T + Tc
    bmin, bmax = posed.min(0) - 0.05, posed.max(0) + 0.05            # box_margin 0.05 (configs/default.yaml)
    ext = (pc[:, 1] / pc[:, 2]).max() - (pc[:, 1] / pc[:, 2]).min()
    focal = fill * H / ext
    cx = W / 2 - focal * 0.5 * ((pc[:, 0] / pc[:, 2]).max() + (pc[:, 0] / pc[:, 2]).min())
    cy = H / 2 - focal * 0.5 * ((pc[:, 1] / pc[:, 2]).max() + (pc[:, 1] / pc[:, 2]).min())
    Kmat = np.array([[focal, 0, cx], [0, focal, cy], [0, 0, 1.0]])

    jj, ii = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
    pixc = np.stack([(ii + 0.5 - cx) / focal, (jj + 0.5 - cy) / focal, np.ones_like(ii, dtype=np.float64)], -1)
    dirs = pixc.reshape(-1, 3) @ R                      # camera -> world (R^T applied to row vectors)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    # slab test against the posed bbox (role of get_near_far, im2mesh/utils/utils.py:56-73)
    with np.errstate(divide='ignore', invalid='ignore'):
        t0 = (bmin - cam_loc) / dirs
        t1 = (bmax - cam_loc) / dirs
    tn = np.minimum(t0, t1).max(axis=1)
    tf = np.maximum(t0, t1).min(axis=1)
    hit = (tf > tn) & (tf > 0)
    pix = np.nonzero(hit)[0]
    near_far = np.stack([np.maximum(tn[pix], 0.0), tf[pix]], -1)

    crng = np.random.default_rng(seed + 4242)
    color = color_net_init(crng, d_in_total={'idr': 417, 'no_view_dir': 390, 'no_normal': 414}[color_mode])
    latent = (crng.normal(size=128) * 0.01).astype(F32)

    pose = np.eye(4)
    pose[:3, :3] = R
    pose[:3, 3] = Tc
    return Frame(H=H, W=W, pix=pix.astype(np.int64), ray_dirs=dirs[pix].astype(F32), near_far=near_far.astype(F32),
                 cam_loc=cam_loc.astype(F32), pose=pose.astype(F32), K=Kmat.astype(F32),
                 bone_transforms=bone_T.astype(F32), smpl_verts=posed.astype(F32), smpl_weights=weights.astype(F32),
                 minimal_shape=verts.astype(F32), trans=trans.astype(F32), coord_min=meta['coord_min'],
                 coord_max=meta['coord_max'], center=meta['center'], sdf=sdf, skin=skin, color=color, latent=latent,
                 beta=F32(beta), n_steps=n_steps, near_samples=near_samples, far_samples=far_samples,
                 cano_view_dirs=cano_view_dirs, color_mode=color_mode)


def canonical_normalisation():
    """coord_min / coord_max / center of the canonical body, as the datasets compute them
    (/root/reference/im2mesh/data/zju_mocap_odp.py:326-331): scalars over all axes after centring."""
    verts = sample_body_vertices(np.random.default_rng(1234))
    center = verts.mean(0)
    vc = verts - center
    return F32(vc.min()), F32(vc.max()), center.astype(F32)


def fold_weight_norm(layer: dict) -> tuple[np.ndarray, np.ndarray]:
    """w = g * v / ||v||_row  (torch.nn.utils.weight_norm, dim=0), float32 like torch._weight_norm."""
    v = layer['v']
    n = np.sqrt((v * v).sum(axis=1, keepdims=True, dtype=F32)).astype(F32)
    return (v * (layer['g'] / n)).astype(F32), layer['b']


def expand_color_weight(W, mode):
    """Bring a RenderingNetwork lin0 / lin3 weight [out, d_in(+128)] of mode 'no_view_dir' / 'no_normal'
    (metaavatar_render/models/decoder.py:101-106) to the 'idr' column layout [points 3 | PE(view) 27 | normals 3 | ...] by
    inserting zero columns: the missing inputs then contribute exact zeros, the result is that of the narrower network."""
    if mode == 'idr':
        return W
    z = np.zeros((W.shape[0], 27 if mode == 'no_view_dir' else 3), W.dtype)
    at = 3 if mode == 'no_view_dir' else 30
    return np.concatenate([W[:, :at], z, W[:, at:]], axis=1)


def train_aux_points(frame: Frame, seed: int = 0, n_uniform: int = 1024, n_inside: int = 256) -> dict:
    """Seeded stand-ins for the extra training inputs the reference's dataset emits
    (/root/reference/im2mesh/data/zju_mocap.py:464-575): ``points_uniform`` (normalised [-1,1]^3, off-surface),
    ``points_skinning`` (24 canonical points, metres) with their target weights, ``points_inside`` (normalised, inside
    the body), a body mask per ray and a pseudo ground-truth colour per ray."""
    rng = np.random.default_rng(seed * 1013 + 7)
    pu = (rng.random((n_uniform, 3)) * 2.0 - 1.0).astype(F32)
    idx = rng.choice(frame.minimal_shape.shape[0], size=N_JOINTS, replace=False)
    ps = frame.minimal_shape[idx].astype(F32)
    pw = frame.smpl_weights[idx].astype(F32)
    d = float(frame.coord_max - frame.coord_min)
    jn = ((JOINTS_CANO - frame.center.astype(np.float64) - float(frame.coord_min) + 0.05 * d) / d / 1.1 - 0.5) * 2.0
    pin = (jn[rng.integers(0, N_JOINTS, size=n_inside)] + rng.normal(scale=0.01, size=(n_inside, 3))).astype(F32)
    body_mask = (rng.random(frame.P) < 0.6)
    rgb_gt = rng.random((frame.P, 3)).astype(F32)
    return {'points_uniform': pu, 'points_skinning': ps, 'sampled_weights': pw, 'points_inside': pin,
            'body_mask': body_mask, 'rgb_gt': rgb_gt}


# ---------------------------------------------------------------------------------------------------- hypernetwork (row f4)
def make_hypernet_state_dict(seed=0, out_scale=0.02, init_scale=0.3):
    """Seeded synthetic parameters of the MetaAvatar hypernetwork, keyed and shaped like
    HyperBVPNet(in_features=3, num_hidden_layers=5, hierarchical_pose=True, hyper_in_ch=144, use_FiLM=True).state_dict()
    (metaavatar/models/siren_modules.py:247-279; 86.6 M parameters).  No checkpoint is reachable offline, and the reference's
    default init zeroes every output layer (hyperlayers.py:418-424) which would make the big GEMVs vanish from the result, so
    every tensor gets small random values (numpy PCG64: identical on every machine).  Returns {key: np.float32 array}."""
    rng = np.random.default_rng(seed)
    sd = {}

    def lin(key, out, inp, scale=None):
        sc = (1.0 / np.sqrt(inp)) if scale is None else scale
        sd[key + '.weight'] = (rng.standard_normal((out, inp), dtype=np.float32) * np.float32(sc))
        sd[key + '.bias'] = (rng.standard_normal(out, dtype=np.float32) * np.float32(0.1))

    lin('pose_encoder.layer_0', 6, 288)
    for j in range(24):
        lin(f'pose_encoder.layers.{j}.0', 19, 19)
        lin(f'pose_encoder.layers.{j}.2', 6, 19)
    in_ch = [3, 256, 256, 256, 256, 256, 256]
    out_ch = [256, 256, 256, 256, 256, 256, 1]
    for l in range(7):
        pre = f'net.layers.{l}.hyper_linear.' if l < 6 else f'net.layers.{l}.'
        n_l = in_ch[l] * out_ch[l] + out_ch[l]
        sd[pre + 'hypo_params_init'] = (rng.standard_normal((1, n_l), dtype=np.float32) * np.float32(init_scale / np.sqrt(in_ch[l])))
        fc = pre + 'hypo_params.net.'
        lin(fc + '0.net.0', 256, 144)
        sd[fc + '0.net.1.weight'] = (1.0 + 0.1 * rng.standard_normal(256, dtype=np.float32)).astype(np.float32)
        sd[fc + '0.net.1.bias'] = (0.1 * rng.standard_normal(256, dtype=np.float32)).astype(np.float32)
        lin(fc + '1.net.0', 256, 256)
        sd[fc + '1.net.1.weight'] = (1.0 + 0.1 * rng.standard_normal(256, dtype=np.float32)).astype(np.float32)
        sd[fc + '1.net.1.bias'] = (0.1 * rng.standard_normal(256, dtype=np.float32)).astype(np.float32)
        lin(fc + '2', n_l, 256, scale=out_scale / 16.0)
    lin('net.mapping_network.network.0', 256, 128)
    lin('net.mapping_network.network.2', 256, 256)
    lin('net.mapping_network.network.4', 256, 256)
    lin('net.mapping_network.network.6', 3072, 256, scale=0.01)
    sd['net.mapping_network.network.6.bias'][:1536] += np.float32(1.0)           # pretrained-siren init: freq ~ 1, phase ~ 0
    return sd


def make_hypernet_inputs(seed=0):
    """rots [1,24,9] (rotation matrices, root = identity as data/zju_mocap_odp.py:262), Jtrs [1,24,3] in [-1,1], latent [1,128]."""
    rng = np.random.default_rng(1000 + seed)
    rots = np.zeros((24, 3, 3), np.float32)
    for j in range(24):
        a = rng.standard_normal(3) * (0.0 if j == 0 else 0.5)
        th = np.linalg.norm(a)
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) if th < 1e-12 else np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
        rots[j] = R.astype(np.float32)
    Jtrs = rng.uniform(-0.8, 0.8, size=(24, 3)).astype(np.float32)
    latent = (0.1 * rng.standard_normal(128)).astype(np.float32)
    return rots.reshape(1, 24, 9), Jtrs.reshape(1, 24, 3), latent.reshape(1, 128)


# ---------------------------------------------------------------------------------------------------- ray set-up (row f3)
def make_smpl_pose_inputs(seed=0, n_verts=6890):
    """Seeded stand-ins for what data/zju_mocap_odp.py:233-276 loads for a frame (no SMPL files offline): minimally clothed
    shape, pose blend-shape basis, pose feature (rotation matrices minus identity, float64 as scipy returns them), skinning
    weights (<= 4 non-zero, rows sum to 1), rigid bone transforms, translation."""
    rng = np.random.default_rng(2000 + seed)
    shape = rng.uniform(-1, 1, size=(n_verts, 3)).astype(np.float32) * np.array([0.35, 0.9, 0.18], np.float32)
    posedirs = (rng.standard_normal((n_verts * 3, 207)) * 0.004).astype(np.float32)
    pose_feature = rng.standard_normal(207) * 0.3
    w = np.zeros((n_verts, 24), np.float32)
    for i in range(n_verts):
        j = rng.choice(24, size=4, replace=False)
        v = rng.uniform(0.05, 1.0, 4); w[i, j] = (v / v.sum()).astype(np.float32)
    B = np.zeros((24, 4, 4), np.float32)
    for j in range(24):
        a = rng.standard_normal(3) * 0.4
        th = np.linalg.norm(a); Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx
        B[j, :3, :3] = R; B[j, :3, 3] = rng.uniform(-0.1, 0.1, 3); B[j, 3, 3] = 1
    return {'minimal_shape': shape, 'posedirs': posedirs, 'pose_feature': pose_feature, 'skinning_weights': w,
            'bone_transforms': B, 'trans': rng.uniform(-0.5, 0.5, 3).astype(np.float32)}


def make_camera(seed, H, W, zoom=1.0):
    """Seeded pinhole camera looking at a body-sized box: (K [3,3], R [3,3], T [3], bounds [2,3]) float32, as the dataset hands
    them to the ray set-up (data/zju_mocap_odp.py:213-231,286-289).  zoom > 1 moves the camera in so that the box leaves the image."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(-0.2, 0.2, 3)
    ext = np.array([rng.uniform(0.5, 0.9), rng.uniform(1.4, 1.9), rng.uniform(0.3, 0.6)])
    bounds = np.stack([c - ext / 2, c + ext / 2]).astype(np.float32)
    ang, el, dist = rng.uniform(0, 2 * np.pi), rng.uniform(-0.3, 0.3), rng.uniform(2.6, 3.6) / zoom
    cam = np.array([dist * np.cos(el) * np.sin(ang), dist * np.sin(el), dist * np.cos(el) * np.cos(ang)])
    z = -cam / np.linalg.norm(cam); x = np.cross([0.0, 1.0, 0.0], z); x /= np.linalg.norm(x); y = np.cross(z, x)
    R = np.stack([x, y, z]).astype(np.float32)
    T = (-R.astype(np.float64) @ cam).astype(np.float32)
    f = 1.05 * W
    K = np.array([[f, 0, W / 2 + rng.uniform(-15, 15)], [0, f, H / 2 + rng.uniform(-15, 15)], [0, 0, 1]], np.float32)
    return K, R, T, bounds
