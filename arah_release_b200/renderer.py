"""Drop-in host side of the B200-native ARAH renderer.

`IDHRNetwork` and `BodyRayTracing` keep the reference's constructor signatures, attribute names (so state_dict keys and
the aliasing `color_decoder.* == idhr_network.rendering_network.*` survive strict checkpoint loads) and return
signatures:

    IDHRNetwork      /root/reference/im2mesh/metaavatar_render/renderer/implicit_differentiable_renderer.py:15-259
    BodyRayTracing   /root/reference/im2mesh/metaavatar_render/renderer/ray_tracing.py:13-172

Their eval `forward` packs the modules' weights + the per-frame buffers into an ArahFrame and calls the C ABI
(include/arah_b200.h) on torch's current CUDA stream.  There is no PyTorch / CPU implementation of the path here: if the
CUDA library is missing or the tensors are not on a CUDA device, the call raises.

Training (`self.training == True`, implicit_differentiable_renderer.py:73-78,117-178,235-249): the tracer runs in training
mode inside the CUDA library (all rays enter the joint search, jittered z samples); the differentiable part is three
`torch.autograd.Function`s whose forward AND backward are hand-written kernels behind the C ABI (arah_train_*): shading
(implicit-gradient LBS correction, SDF, SDF input gradient as normal, colour MLP, compositing), auxiliary SDF evaluations
(eikonal / off-surface / inside) and skinning-weight prediction.  torch only applies weight-norm and ||variance|| (tiny
parameter-space ops) and carries the graph to the hypernetwork / optimiser.
"""
import ctypes as C
import os
import weakref

from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib
from ._lib import ArahConfig, ArahFrame, ArahStats, ArahTrainGrads, check

N_VERTS_DEFAULT = 6890
# Error bound assumed for the fp16 pass of the banded lattice (raw network output units): ~10x the largest deviation measured on
# the 256^3 fixtures (tests/test_gpu_mesh.py prints it); the library checks the bound on every refined point at run time.
BAND_EPS = 0.02


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _f32c(t, device=None):
    t = t.detach()
    if device is not None and t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _effective_weight(lin):
    """`lin.weight` of a (possibly weight-normed) nn.Linear: w = g * v / ||v||  (torch.nn.utils.weight_norm, dim 0)."""
    if hasattr(lin, 'weight_g') and hasattr(lin, 'weight_v'):
        return torch._weight_norm(lin.weight_v.detach(), lin.weight_g.detach(), 0)
    if hasattr(lin, 'parametrizations') and hasattr(lin.parametrizations, 'weight'):
        return lin.weight.detach()
    return lin.weight.detach()


def _effective_weight_grad(lin):
    """Same as _effective_weight but keeping the autograd graph to weight_g / weight_v (training)."""
    if hasattr(lin, 'weight_g') and hasattr(lin, 'weight_v'):
        return torch._weight_norm(lin.weight_v, lin.weight_g, 0)
    return lin.weight


# RenderingNetwork modes (metaavatar_render/models/decoder.py:101-106).  The kernels evaluate the 'idr' input layout
# [points 3 | PE4(view) 27 | normals 3 | feature 256 | latent]; the two narrower modes are brought to it by inserting ZERO weight
# columns for the inputs they do not have (lin0 and the skip layer lin3): the extra products are exact zeros, so the result
# is that of the narrower network.  value = width of the non-latent part of the input.
_COLOR_BASE = {'idr': 289, 'no_view_dir': 262, 'no_normal': 286}


def _color_mode(rn):
    mode = getattr(rn, 'mode', 'idr')
    if mode not in _COLOR_BASE or list(getattr(rn, 'skips', [3])) != [3]:
        raise _lib.ArahError("renderer modes 'idr' / 'no_view_dir' / 'no_normal' with skips=[3] are implemented")
    has_view = getattr(rn, 'embedview_fn', 1) is not None
    if (mode != 'no_view_dir' and not has_view) or getattr(rn, 'embed_fn', None) is not None:
        raise _lib.ArahError('the colour net must use multires_view=4 (when it takes view directions) and multires=0')
    return mode


def _expand_color_weight(W, mode):
    if mode == 'idr':
        return W
    at, n = (3, 27) if mode == 'no_view_dir' else (30, 3)
    return torch.cat([W[:, :at], W.new_zeros(W.shape[0], n), W[:, at:]], dim=1)


class ArahRenderer:
    """Thin RAII wrapper around an ArahHandle (one per device/stream; not thread-safe)."""

    def __init__(self, device, n_steps=64, near_samples=16, far_samples=16, cano_view_dirs=True, latent_dim=128,
                 n_verts=N_VERTS_DEFAULT, max_rays=65536, shade_mode=None, root_mode=None, shade_cull=None, render_last_pt=False):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.ArahError('the ARAH hot path only exists as CUDA kernels; got device %s' % device)
        if shade_mode is None:
            shade_mode = os.environ.get('ARAH_SHADE_MODE', 'tf32')
        mode = {'tf32': 0, 'fp32': 1}[shade_mode] if isinstance(shade_mode, str) else int(shade_mode)
        self.shade_mode = 'fp32' if mode == 1 else 'tf32'
        if root_mode is None:
            root_mode = os.environ.get('ARAH_ROOT_MODE', '3xtf32')
        rmode = {'3xtf32': 0, 'fp32': 1}[root_mode] if isinstance(root_mode, str) else int(root_mode)
        self.root_mode = 'fp32' if rmode == 1 else '3xtf32'
        self.cfg = ArahConfig(device=self.device.index or 0, n_steps=n_steps, near_samples=near_samples,
                              far_samples=far_samples, cano_view_dirs=int(bool(cano_view_dirs)), latent_dim=latent_dim,
                              n_verts=n_verts, max_rays=max_rays, shade_mode=mode, root_mode=rmode,
                              shade_cull=0 if (shade_cull is None or shade_cull) else 1, render_last_pt=int(bool(render_last_pt)))
        self.shade_cull = bool(self.cfg.shade_cull == 0) and os.environ.get('ARAH_SHADE_CULL', '1') != '0'
        self._h = C.c_void_p()
        check(_lib.lib().arah_create(C.byref(self.cfg), C.byref(self._h)))
        self._keep = []
        self.n_steps = n_steps

    def close(self):
        if getattr(self, '_h', None) and self._h.value:
            _lib.lib().arah_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ frame
    def set_frame(self, *, sdf_W, sdf_b, sdf_freq, sdf_phase, skin_W, skin_b, col_W, col_b, latent, beta,
                  bone_transforms, smpl_verts, smpl_weights, trans, coord_min, coord_max, center, cam_loc, pose,
                  pose_on_host=False):
        """All weight arguments: lists of fp32 CUDA tensors in the reference layout [out, in]."""
        dev = self.device
        keep = []
        fr = ArahFrame()

        def d(t):
            t = _f32c(t, dev)
            keep.append(t)
            return _ptr(t)

        def hd(t):
            if pose_on_host:
                t = t.detach().float().contiguous()
                if t.device.type != 'cpu':
                    t = t.cpu()
                keep.append(t)
                return _ptr(t)
            return d(t)
        for i in range(7):
            fr.sdf_W[i] = d(sdf_W[i]); fr.sdf_b[i] = d(sdf_b[i])
        fr.sdf_freq = d(sdf_freq); fr.sdf_phase = d(sdf_phase)
        for i in range(5):
            fr.skin_W[i] = d(skin_W[i]); fr.skin_b[i] = d(skin_b[i])
        for i in range(6):
            fr.col_W[i] = d(col_W[i]); fr.col_b[i] = d(col_b[i])
        fr.latent = d(latent) if latent is not None else None
        fr.beta = float(beta)
        fr.bone_transforms = hd(bone_transforms)
        fr.smpl_verts = hd(smpl_verts)
        fr.smpl_weights = hd(smpl_weights) if smpl_weights is not None else None
        fr.pose_on_host = int(bool(pose_on_host))
        fr.trans[:] = [float(v) for v in trans]
        fr.coord_min = float(coord_min); fr.coord_max = float(coord_max)
        fr.center[:] = [float(v) for v in center]
        fr.cam_loc[:] = [float(v) for v in cam_loc]
        fr.pose[:] = [float(v) for v in pose]
        check(_lib.lib().arah_set_frame(self._h, C.byref(fr), self.stream))
        self._keep = keep        # packed copies live in the handle; inputs only needed until the call returned

    def set_frame_from_modules(self, sdf_network, skinning_model, rendering_network, deviation_network, inputs,
                               pose_on_host=False):
        """Read weights out of modules laid out like the reference's (tools/ref_layout.py builds such modules for tests and bench.py) + the input dict."""
        sdf_W, sdf_b = [], []
        freq, phase = [], []
        if len(sdf_network) != 7:
            raise _lib.ArahError('expected the 7-layer FiLM-SIREN of configs/arah-*/ (hyper_bvp, num_hidden_layers 5)')
        for l in range(6):
            film = sdf_network[l][0]
            if not hasattr(film, 'freq'):
                raise _lib.ArahError('sdf_network layers must be BatchLinearFiLM (use_FiLM: true)')
            sdf_W.append(film.weights.reshape(256, -1)); sdf_b.append(film.biases.reshape(-1))
            freq.append(film.freq.reshape(-1)); phase.append(film.phase_shift.reshape(-1))
        sdf_W.append(sdf_network[6].weights.reshape(1, 256)); sdf_b.append(sdf_network[6].biases.reshape(-1))
        dec = skinning_model.skinning_decoder_fwd
        skin_W = [_effective_weight(getattr(dec, f'lin{i}')) for i in range(5)]
        skin_b = [getattr(dec, f'lin{i}').bias for i in range(5)]
        if skin_W[4].shape[0] != 25:
            raise _lib.ArahError('skinning decoder must have 25 outputs (hierarchical softmax)')
        rn = rendering_network
        cmode = _color_mode(rn)
        pe = getattr(rn, 'pose_encoder_type', None)
        col_W = [_effective_weight(getattr(rn, f'lin{i}')) for i in range(6)]
        col_b = [getattr(rn, f'lin{i}').bias for i in range(6)]
        latent_dim = col_W[0].shape[1] - _COLOR_BASE[cmode]
        col_W[0], col_W[3] = _expand_color_weight(col_W[0], cmode), _expand_color_weight(col_W[3], cmode)
        if latent_dim != self.cfg.latent_dim:
            raise _lib.ArahError(f'colour net expects a {latent_dim}-d pose feature, handle was built for {self.cfg.latent_dim}')
        latent = None
        if latent_dim > 0:
            if pe != 'latent':
                raise _lib.ArahError("only color_pose_encoder 'latent' (or None) is implemented")
            latent = inputs['pose_cond']['latent_code'].reshape(-1)
        # ---- per-frame scalars without stalling the stream more than once.  beta = |variance| changes only with an optimiser step:
        # cached against the parameter's version counter.  The geometry scalars come from `inputs['host_scalars']` when the caller
        # still has the host values the dataset produced (no device read-back at all); otherwise ONE batched device -> host copy.
        var = deviation_network.variance
        key = (var.data_ptr(), var._version)
        if getattr(self, '_beta_key', None) != key:
            self._beta = float(torch.linalg.norm(var.detach()).item())
            self._beta_key = key
        beta = self._beta
        hs = inputs.get('host_scalars')
        if hs is not None:
            trans, cmin, cmax = [float(v) for v in hs['trans']], float(hs['coord_min']), float(hs['coord_max'])
            center, cam_loc, pose = [float(v) for v in hs['center']], [float(v) for v in hs['cam_loc']], [float(v) for v in hs['pose']]
        else:
            flat = torch.cat([inputs['trans'].reshape(-1)[:3].float(), inputs['coord_min'].reshape(-1)[:1].float(), inputs['coord_max'].reshape(-1)[:1].float(),
                              inputs['center'].reshape(-1)[:3].float(), inputs['cam_loc'].reshape(-1)[:3].float(), inputs['pose'].reshape(-1)[:16].float()]).tolist()
            trans, cmin, cmax, center, cam_loc, pose = flat[0:3], flat[3], flat[4], flat[5:8], flat[8:11], flat[11:27]
        self.set_frame(sdf_W=sdf_W, sdf_b=sdf_b, sdf_freq=torch.stack(freq), sdf_phase=torch.stack(phase), skin_W=skin_W,
                       skin_b=skin_b, col_W=col_W, col_b=col_b, latent=latent, beta=beta,
                       bone_transforms=inputs['bone_transforms'][0], smpl_verts=inputs['smpl_verts'][0],
                       smpl_weights=inputs['skinning_weights'][0], trans=trans, coord_min=cmin, coord_max=cmax,
                       center=center, cam_loc=cam_loc, pose=pose, pose_on_host=pose_on_host)

    # ------------------------------------------------------------------ render
    def render(self, ray_dirs, near_far, want_weights=False):
        """ray_dirs [P,3], near_far [P,2] fp32 CUDA -> rgb [P,3], mask [P] bool, points_cam [P,3] (+ weights_sum [P])."""
        rd = _f32c(ray_dirs, self.device).view(-1, 3)
        nf = _f32c(near_far, self.device).view(-1, 2)
        P = rd.shape[0]
        rgb = torch.empty(P, 3, device=self.device, dtype=torch.float32)
        mask = torch.empty(P, device=self.device, dtype=torch.uint8)
        pc = torch.empty(P, 3, device=self.device, dtype=torch.float32)
        ws = torch.empty(P, device=self.device, dtype=torch.float32) if want_weights else None
        check(_lib.lib().arah_render(self._h, _ptr(rd), _ptr(nf), P, _ptr(rgb), _ptr(mask), _ptr(pc),
                                     _ptr(ws) if ws is not None else None, self.stream))
        self._io_keep = (rd, nf)
        out = (rgb, mask.bool(), pc)
        return out + (ws,) if want_weights else out

    def render_host(self, ray_dirs, near_far, rgb=None, mask=None, points_cam=None):
        """Host (ideally pinned) fp32 tensors in, host tensors out; H2D/D2H happen inside the C call."""
        rd = ray_dirs.contiguous().view(-1, 3)
        nf = near_far.contiguous().view(-1, 2)
        assert rd.device.type == 'cpu' and rd.dtype == torch.float32 and nf.dtype == torch.float32
        P = rd.shape[0]
        if rgb is None or mask is None or points_cam is None:
            # pinned result buffers are owned by the renderer and reused (cudaHostAlloc per call costs milliseconds); the returned
            # tensors are views that stay valid until the next render_host call
            cache = getattr(self, '_host_out', None)
            if cache is None or cache[0].shape[0] < P:
                cache = (torch.empty(P, 3, dtype=torch.float32).pin_memory(), torch.empty(P, dtype=torch.uint8).pin_memory(),
                         torch.empty(P, 3, dtype=torch.float32).pin_memory())
                self._host_out = cache
            rgb = cache[0][:P] if rgb is None else rgb
            mask = cache[1][:P] if mask is None else mask
            pc = cache[2][:P] if points_cam is None else points_cam
        else:
            pc = points_cam
        check(_lib.lib().arah_render_host(self._h, _ptr(rd), _ptr(nf), P, _ptr(rgb), _ptr(mask), _ptr(pc), self.stream))
        return rgb, mask, pc

    def trace_outputs(self, P, transforms=True, points=True):
        """BodyRayTracing.forward's 7-tuple for the last render (ray_tracing.py:166-172), batch dim added."""
        S, dev = self.n_steps, self.device
        pts_hat = torch.empty(P, 3, device=dev); m = torch.empty(P, device=dev, dtype=torch.uint8); dists = torch.empty(P, device=dev)
        sp = torch.empty(P, S, 3, device=dev) if points else None; sd = torch.empty(P, S, device=dev) if points else None
        sT = torch.empty(P, S, 4, 4, device=dev) if transforms else None
        sc = torch.empty(P, S, device=dev, dtype=torch.uint8)
        check(_lib.lib().arah_get_trace(self._h, _ptr(pts_hat), _ptr(m), _ptr(dists), _ptr(sp) if sp is not None else None,
                                        _ptr(sd) if sd is not None else None, _ptr(sT) if sT is not None else None, _ptr(sc), self.stream))
        return (pts_hat.unsqueeze(0), m.bool().unsqueeze(0), dists.unsqueeze(0), sp.unsqueeze(0) if sp is not None else None,
                sd.unsqueeze(0) if sd is not None else None, sT.unsqueeze(0) if sT is not None else None, sc.bool().unsqueeze(0))

    # ------------------------------------------------------------------ training (arah_train_*)
    def set_training(self, mode='3xtf32'):
        """mode: False/None = off, '3xtf32' (tcgen05 tensor cores, split precision; default), 'fp32' (SIMT FFMA GEMMs) or
        'tf32' (single-pass TF32 operands)."""
        code = 0 if not mode else {'3xtf32': 1, 'fp32': 2, 'tf32': 3, True: 1}[mode]
        check(_lib.lib().arah_set_training(self._h, code))

    def train_trace(self, ray_dirs, near_far, u_all, u_near, u_far):
        """BodyRayTracing.forward(eval_mode=False); results via trace_outputs()."""
        rd = _f32c(ray_dirs, self.device).view(-1, 3)
        nf = _f32c(near_far, self.device).view(-1, 2)
        P = rd.shape[0]
        ua, un = _f32c(u_all, self.device).view(P, -1), _f32c(u_near, self.device).view(P, -1)
        uf = _f32c(u_far, self.device).view(P, -1) if u_far is not None and u_far.numel() else None
        check(_lib.lib().arah_train_trace(self._h, _ptr(rd), _ptr(nf), P, _ptr(ua), _ptr(un), _ptr(uf) if uf is not None else None, self.stream))
        self._io_keep = (rd, nf, ua, un, uf)
        self._train_P = P
        return P

    def _grad_table(self, shapes):
        """Zero-filled gradient buffers + the ArahTrainGrads pointing at them.  shapes: dict name -> shape (None = skip)."""
        g = ArahTrainGrads()
        bufs = {}
        def z(name):
            shp = shapes.get(name)
            if shp is None:
                return None
            t = torch.zeros(shp, device=self.device, dtype=torch.float32)
            bufs[name] = t
            return _ptr(t)
        for i in range(7):
            g.sdf_W[i] = z(f'sdf_W{i}'); g.sdf_b[i] = z(f'sdf_b{i}')
        g.sdf_freq = z('sdf_freq'); g.sdf_phase = z('sdf_phase')
        for i in range(5):
            g.skin_W[i] = z(f'skin_W{i}'); g.skin_b[i] = z(f'skin_b{i}')
        for i in range(6):
            g.col_W[i] = z(f'col_W{i}'); g.col_b[i] = z(f'col_b{i}')
        g.latent = z('latent'); g.beta = z('beta')
        return g, bufs

    def set_profiling(self, enable=True):
        check(_lib.lib().arah_set_profiling(self._h, int(bool(enable))))

    def stats(self):
        s = ArahStats()
        check(_lib.lib().arah_get_stats(self._h, C.byref(s), self.stream))
        return s.as_dict()

    def phase_clocks(self):
        out = (C.c_uint64 * 32)()
        check(_lib.lib().arah_debug_phase_clocks(self._h, out, self.stream))
        return [int(v) for v in out]

    def eval_sdf(self, xn, grad=True, feat=False):
        xn = _f32c(xn, self.device).view(-1, 3)
        n = xn.shape[0]
        s = torch.empty(n, device=self.device)
        g = torch.empty(n, 3, device=self.device) if grad else None
        f = torch.empty(n, 256, device=self.device) if feat else None
        check(_lib.lib().arah_eval_sdf(self._h, _ptr(xn), n, _ptr(s), _ptr(g) if grad else None, _ptr(f) if feat else None, self.stream))
        return s, g, f

    def eval_skin(self, x_hat):
        x = _f32c(x_hat, self.device).view(-1, 3)
        n = x.shape[0]
        w = torch.empty(n, 24, device=self.device)
        xb = torch.empty(n, 3, device=self.device)
        check(_lib.lib().arah_eval_skin(self._h, _ptr(x), n, _ptr(w), _ptr(xb), self.stream))
        return w, xb

    def knn(self, pts):
        """pytorch3d.ops.knn_points(K=1) of ray_tracing.py:386,407: nearest posed SMPL vertex per point -> int32 [n]."""
        x = _f32c(pts, self.device).view(-1, 3)
        idx = torch.empty(x.shape[0], dtype=torch.int32, device=self.device)
        check(_lib.lib().arah_debug_knn(self._h, _ptr(x), x.shape[0], _ptr(idx), self.stream))
        return idx

    # ------------------------------------------------------------------ canonical mesh
    def sdf_grid(self, N=256):
        """utils/sdf_meshing.py:13-58: the frame's SDF network on the N^3 lattice over [-1,1]^3 -> [N, N, N] (device)."""
        if not 2 <= int(N) <= 1024:
            raise _lib.ArahError('lattice side must be in [2, 1024]')
        out = torch.empty(N, N, N, device=self.device)
        check(_lib.lib().arah_sdf_grid(self._h, int(N), _ptr(out), self.stream))
        return out

    def sdf_grid_banded(self, N=256, level=0.0, eps=BAND_EPS):
        """The lattice for a caller that only extracts the `level` iso-surface from it (include/arah_b200.h, arah_sdf_grid_banded):
        one fp16 pass over all N^3 points, split precision for the corners of every cell within eps of straddling the level.
        Returns (vol [N, N, N], stats int32[2] = [points refined, refined points whose coarse value was off by more than eps]),
        both on the device; marching_cubes(vol, level) equals marching_cubes(sdf_grid(N), level) bit for bit while stats[1] == 0."""
        if not 2 <= int(N) <= 1024:
            raise _lib.ArahError('lattice side must be in [2, 1024]')
        out = torch.empty(N, N, N, device=self.device)
        stats = torch.zeros(2, dtype=torch.int32, device=self.device)
        check(_lib.lib().arah_sdf_grid_banded(self._h, int(N), float(level), float(eps), _ptr(out), _ptr(stats), self.stream))
        return out, stats

    def canonical_mesh(self, N=256, level=0.0, voxel_size=None, origin=(-1.0, -1.0, -1.0)):
        """SDF lattice + iso-surface (utils/sdf_meshing.py:13-114) -> device (verts, faces).  Tensor-core modes use the banded
        lattice and fall back to the full-precision one if its run-time error check fires (never observed)."""
        if self.root_mode != '3xtf32' or os.environ.get('ARAH_GRID_BANDED', '1') == '0':
            return self.marching_cubes(self.sdf_grid(N), level=level, voxel_size=voxel_size, origin=origin)
        vol, stats = self.sdf_grid_banded(N, level=level)
        verts, faces = self.marching_cubes(vol, level=level, voxel_size=voxel_size, origin=origin)     # (synchronises: counts)
        self.last_band_stats = [int(v) for v in stats.tolist()]
        if self.last_band_stats[1] != 0:
            verts, faces = self.marching_cubes(self.sdf_grid(N), level=level, voxel_size=voxel_size, origin=origin)
        return verts, faces

    def marching_cubes(self, vol, level=0.0, voxel_size=None, origin=(-1.0, -1.0, -1.0), max_verts=None, max_faces=None):
        """utils/sdf_meshing.py:69-114 on the GPU: (verts [nv, 3] float32, faces [nf, 3] int32), both device tensors.
        One host synchronisation (the vertex / face counts size the result)."""
        vol = _f32c(vol, self.device)
        N = vol.shape[0]
        if tuple(vol.shape) != (N, N, N):
            raise _lib.ArahError('marching_cubes needs a cubic [N, N, N] lattice')
        voxel_size = 2.0 / (N - 1) if voxel_size is None else float(voxel_size)
        org = (C.c_float * 3)(*[float(v) for v in origin])
        counts = torch.zeros(2, dtype=torch.int32, device=self.device)
        mv = int(max_verts) if max_verts else 6 * N * N
        mf = int(max_faces) if max_faces else 12 * N * N
        ws_bytes = int(_lib.lib().arah_marching_cubes_workspace(N))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)       # torch's caching allocator: no cudaMalloc per call
        while True:
            verts = torch.empty(mv, 3, device=self.device)
            faces = torch.empty(mf, 3, dtype=torch.int32, device=self.device)
            check(_lib.lib().arah_marching_cubes(_ptr(vol), N, float(level), voxel_size, org, _ptr(verts), mv, _ptr(faces), mf,
                                                 _ptr(counts), _ptr(ws), ws_bytes, self.stream))
            nv, nf = (int(v) for v in counts.tolist())
            if nv <= mv and nf <= mf:
                return verts[:nv], faces[:nf]
            mv, mf = max(mv, nv), max(mf, nf)


def create_mesh_vertices_and_faces(renderer, N=256, max_batch=64 ** 3, offset=None, scale=None, **kwargs):
    """Drop-in for im2mesh.utils.sdf_meshing.create_mesh_vertices_and_faces (utils/sdf_meshing.py:13-66) with the frame's
    ArahRenderer in place of the `decoder` module: SDF lattice on the tensor cores, iso-surface on the GPU; returns numpy
    (mesh_points [nv, 3], faces [nf, 3]) like the reference.  `max_batch` is accepted and ignored (no chunking needed)."""
    verts, faces = renderer.canonical_mesh(N, level=0.0, voxel_size=2.0 / (N - 1), origin=(-1.0, -1.0, -1.0))
    if scale is not None:
        verts = verts / scale
    if offset is not None:
        verts = verts - offset
    return verts.cpu().numpy(), faces.cpu().numpy()


# =====================================================================================================================
# training: torch.autograd.Function wrappers around the hand-written forward/backward pairs of the C ABI
# =====================================================================================================================
SDF_NAMES = [f'sdf_W{i}' for i in range(7)] + [f'sdf_b{i}' for i in range(7)] + ['sdf_freq', 'sdf_phase']
SKIN_NAMES = [f'skin_W{i}' for i in range(5)] + [f'skin_b{i}' for i in range(5)]
COL_NAMES = [f'col_W{i}' for i in range(6)] + [f'col_b{i}' for i in range(6)] + ['latent', 'beta']


def _shapes(names, tensors):
    return {n: tuple(t.shape) for n, t in zip(names, tensors)}


class _ShadeFn(torch.autograd.Function):
    """get_rbg_value_vol_sdf in training mode for all rays (implicit_differentiable_renderer.py:261-396)."""

    @staticmethod
    def forward(ctx, r, view, view_orig, ray_augm, train_skin, *params):
        P = r._train_P
        rgb = torch.empty(P, 3, device=r.device, dtype=torch.float32)
        ws = torch.empty(P, device=r.device, dtype=torch.float32)
        v = _f32c(view, r.device).view(-1, 3)
        vo = _f32c(view_orig, r.device).view(-1, 3) if view_orig is not None else None
        check(_lib.lib().arah_train_shade_forward(r._h, _ptr(v), _ptr(vo) if vo is not None else None, int(bool(ray_augm)),
                                                  int(bool(train_skin)), _ptr(rgb), _ptr(ws), r.stream))
        ctx.r = r
        ctx.names = SDF_NAMES + SKIN_NAMES + COL_NAMES
        ctx.shapes = _shapes(ctx.names, params)
        ctx.train_skin = bool(train_skin)
        ctx.keep = (v, vo)
        return rgb, ws

    @staticmethod
    def backward(ctx, g_rgb, g_ws):
        r = ctx.r
        shapes = dict(ctx.shapes)
        if not ctx.train_skin:
            for n in SKIN_NAMES:
                shapes[n] = None
        g, bufs = r._grad_table(shapes)
        g_rgb = _f32c(g_rgb, r.device)
        g_ws = _f32c(g_ws, r.device) if g_ws is not None else None
        check(_lib.lib().arah_train_shade_backward(r._h, _ptr(g_rgb), _ptr(g_ws) if g_ws is not None else None, C.byref(g), r.stream))
        return (None, None, None, None, None) + tuple(bufs.get(n) for n in ctx.names)


class _SdfFn(torch.autograd.Function):
    """sdf_network(points) [+ gradient(sdf, points)] for the regularisers (implicit_differentiable_renderer.py:117-140)."""

    @staticmethod
    def forward(ctx, r, slot, points, with_grad, *params):
        pts = _f32c(points, r.device).view(-1, 3)
        n = pts.shape[0]
        sdf = torch.empty(n, 1, device=r.device, dtype=torch.float32)
        grad = torch.empty(n, 3, device=r.device, dtype=torch.float32) if with_grad else None
        check(_lib.lib().arah_train_sdf_forward(r._h, int(slot), _ptr(pts), n, int(bool(with_grad)), _ptr(sdf),
                                                _ptr(grad) if grad is not None else None, r.stream))
        ctx.r, ctx.slot, ctx.with_grad = r, int(slot), bool(with_grad)
        ctx.shapes = _shapes(SDF_NAMES, params)
        ctx.keep = pts
        if with_grad:
            return sdf, grad
        return sdf

    @staticmethod
    def backward(ctx, g_sdf, g_grad=None):
        r = ctx.r
        g, bufs = r._grad_table(ctx.shapes)
        gs = _f32c(g_sdf, r.device) if g_sdf is not None else None
        gg = _f32c(g_grad, r.device) if (g_grad is not None and ctx.with_grad) else None
        check(_lib.lib().arah_train_sdf_backward(r._h, ctx.slot, _ptr(gs) if gs is not None else None, _ptr(gg) if gg is not None else None,
                                                 C.byref(g), r.stream))
        return (None, None, None, None) + tuple(bufs.get(n) for n in SDF_NAMES)


class _SkinFn(torch.autograd.Function):
    """query_weights(points_skinning) (utils/root_finding_utils.py:54-113)."""

    @staticmethod
    def forward(ctx, r, points, *params):
        pts = _f32c(points, r.device).view(-1, 3)
        n = pts.shape[0]
        w = torch.empty(n, 24, device=r.device, dtype=torch.float32)
        check(_lib.lib().arah_train_skin_forward(r._h, _ptr(pts), n, _ptr(w), r.stream))
        ctx.r = r
        ctx.shapes = _shapes(SKIN_NAMES, params)
        ctx.keep = pts
        return w

    @staticmethod
    def backward(ctx, g_w):
        r = ctx.r
        g, bufs = r._grad_table(ctx.shapes)
        gw = _f32c(g_w, r.device)
        check(_lib.lib().arah_train_skin_backward(r._h, _ptr(gw), C.byref(g), r.stream))
        return (None, None) + tuple(bufs.get(n) for n in SKIN_NAMES)


# =====================================================================================================================
class BodyRayTracing(nn.Module):
    """Ray-tracer for the articulated body SDF — same constructor as the reference (ray_tracing.py:16-49).

    `forward` returns the reference's 7-tuple.  The tracing itself runs inside the fused CUDA path; when called
    stand-alone this module renders the frame through its own ArahRenderer and hands back the tracer outputs.
    """

    def __init__(self, root_finding_threshold=1.0e-5, sphere_tracing_iters=50, n_steps=64, near_surface_vol_samples=16,
                 far_surface_vol_samples=16, surface_vol_range=0.05, sample_bg_pts=0, low_vram=False):
        super().__init__()
        if abs(root_finding_threshold - 1e-5) > 1e-12 or sphere_tracing_iters != 50 or abs(surface_vol_range - 0.05) > 1e-12:
            raise _lib.ArahError('kernels are built for root_finding_threshold=1e-5, 50 sphere-tracing iterations, '
                                 'surface_vol_range=0.05 (the values MetaAvatarRender uses, models/__init__.py:75)')
        self.root_finding_threshold = root_finding_threshold
        self.sphere_tracing_iters = sphere_tracing_iters
        self.n_steps = n_steps
        self.near_surface_vol_samples = near_surface_vol_samples
        self.surface_vol_range = surface_vol_range
        self.far_surface_vol_samples = far_surface_vol_samples
        self.sample_bg_pts = sample_bg_pts
        self.low_vram = low_vram          # accepted and ignored: the kernels never materialise P x 64 x 16 intermediates
        object.__setattr__(self, '_owner', None)   # weakref to the owning IDHRNetwork (not a registered submodule)

    def forward(self, sdf_network, skinning_model, cam_loc, ray_directions, body_bounds_intersections, loc, sc_factor,
                smpl_verts, smpl_verts_cano, skinning_weights, vol_feat, bone_transforms, trans, coord_min, coord_max, center,
                eval_mode=False):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise _lib.ArahError('BodyRayTracing must be owned by an IDHRNetwork (it shares its renderer handle)')
        if isinstance(vol_feat, OrderedDict):
            # query_weights / forward_skinning normalise with (x - loc) * sc_factor for this kind of skinning decoder
            # (root_finding_utils.py:66-72); the kernels implement the plain coord_min / coord_max normalisation only
            raise _lib.ArahError('vol_feat as an OrderedDict (the (x - loc) * sc_factor normalisation branch) is not supported')
        out = owner._trace_only(sdf_network, cam_loc, ray_directions, body_bounds_intersections, smpl_verts,
                                skinning_weights, bone_transforms, trans, coord_min, coord_max, center, train=not eval_mode)
        if self.sample_bg_pts > 0 and not eval_mode:
            # background depths for a NeRF background model (ray_tracing.py:375-378; IDHRNetwork drops them again, :108-110):
            # far / flip(linspace(1e-3, 1 - 1 / (n + 1), n)) per ray
            n = self.sample_bg_pts
            zo = torch.linspace(1e-3, 1.0 - 1.0 / (n + 1.0), n, device=ray_directions.device, dtype=torch.float32).view(1, 1, -1)
            zo = body_bounds_intersections[..., 1:] / torch.flip(zo, dims=[-1])
            out = out[:4] + ((out[4], zo),) + out[5:]
        return out

    def draw_jitter(self, P):
        """The three uniform draws of ray_sampler in training mode, in the reference's order and shapes and — like the
        reference, which calls torch.rand(shape).to(device) (ray_tracing.py:305) — from the CPU generator."""
        u_all = torch.rand(1, P, self.n_steps)
        u_near = torch.rand(1, P, self.near_surface_vol_samples + 1)
        u_far = torch.rand(1, P, self.far_surface_vol_samples) if self.far_surface_vol_samples > 0 else None
        return u_all, u_near, u_far


class IDHRNetwork(nn.Module):
    """Implicit Differentiable Human Renderer — same constructor and attributes as the reference
    (implicit_differentiable_renderer.py:18-40); eval forward runs entirely in libarah_b200.so."""

    def __init__(self, deviation_network, rendering_network, skinning_model, ray_tracer, cano_view_dirs=True,
                 train_skinning_net=False, render_last_pt=False, low_vram=False, shade_mode=None, root_mode=None, train_mode=None,
                 shade_cull=None):
        super().__init__()
        self.shade_cull = shade_cull      # extra, optional: False disables the exact alpha cull (bit-identical results either way)
        self.train_mode = train_mode      # extra, optional: '3xtf32' (tensor-core GEMMs in the training engine, default) | 'fp32' | 'tf32'
        self.root_mode = root_mode        # extra, optional: '3xtf32' (tensor cores, default) | 'fp32'
        self.shade_mode = shade_mode      # extra, optional: 'tf32' (tensor cores, default) | 'fp32' (FFMA tiles)
        self.deviation_network = deviation_network
        self.rendering_network = rendering_network
        self.skinning_model = skinning_model
        self.ray_tracer = ray_tracer
        self.cano_view_dirs = cano_view_dirs
        self.train_skinning_net = train_skinning_net
        self.render_last_pt = render_last_pt
        self.low_vram = low_vram
        if isinstance(ray_tracer, BodyRayTracing):
            object.__setattr__(ray_tracer, '_owner', weakref.ref(self))
        self._renderers = {}
        self.last_stats = None

    def _renderer(self, device, latent_dim, n_verts):
        key = (str(device), latent_dim, n_verts)
        r = self._renderers.get(key)
        if r is None:
            rt = self.ray_tracer
            r = ArahRenderer(device, n_steps=rt.n_steps, near_samples=rt.near_surface_vol_samples,
                             far_samples=rt.far_surface_vol_samples, cano_view_dirs=self.cano_view_dirs,
                             latent_dim=latent_dim, n_verts=n_verts, shade_mode=self.shade_mode, root_mode=self.root_mode,
                             shade_cull=self.shade_cull, render_last_pt=self.render_last_pt)
            self._renderers[key] = r
        return r

    def _prepare(self, input, training=False):
        ray_dirs = input['ray_dirs']
        if ray_dirs.device.type != 'cuda':
            raise _lib.ArahError('IDHRNetwork (B200) needs CUDA tensors; there is no CPU fallback')
        if ray_dirs.shape[0] != 1:
            raise _lib.ArahError('one frame per call (the reference assumes the same, ray_tracing.py:129-132)')
        latent_dim = _effective_weight(self.rendering_network.lin0).shape[1] - _COLOR_BASE[_color_mode(self.rendering_network)]
        r = self._renderer(ray_dirs.device, latent_dim, input['smpl_verts'].shape[1])
        r.set_training((self.train_mode or os.environ.get('ARAH_TRAIN_MODE', '3xtf32')) if training else False)
        r.set_frame_from_modules(input['sdf_network'], self.skinning_model, self.rendering_network, self.deviation_network, input)
        return r

    # ------------------------------------------------------------------ training
    def _param_tensors(self, input):
        """Parameter tensors WITH their autograd graph, in the order / layout of ArahTrainGrads."""
        net = input['sdf_network']
        sdf = [net[l][0].weights.reshape(256, -1) for l in range(6)] + [net[6].weights.reshape(1, 256)]
        sdf += [net[l][0].biases.reshape(-1) for l in range(6)] + [net[6].biases.reshape(-1)]
        sdf += [torch.stack([net[l][0].freq.reshape(-1) for l in range(6)]), torch.stack([net[l][0].phase_shift.reshape(-1) for l in range(6)])]
        dec = self.skinning_model.skinning_decoder_fwd
        skin = [_effective_weight_grad(getattr(dec, f'lin{i}')) for i in range(5)] + [getattr(dec, f'lin{i}').bias for i in range(5)]
        rn = self.rendering_network
        col = [_effective_weight_grad(getattr(rn, f'lin{i}')) for i in range(6)] + [getattr(rn, f'lin{i}').bias for i in range(6)]
        cmode = _color_mode(rn)            # autograd routes the gradient of the zero-padded matrices back to the real columns
        col[0], col[3] = _expand_color_weight(col[0], cmode), _expand_color_weight(col[3], cmode)
        lat = input['pose_cond'].get('latent_code') if isinstance(input.get('pose_cond'), dict) else None
        if lat is None:
            lat = torch.zeros(0, device=sdf[0].device)
        col += [lat.reshape(-1), torch.linalg.norm(self.deviation_network.variance).reshape(1)]
        return sdf, skin, col

    def _rand_device(self, shape, device):
        """torch.rand on the compute device (the eikonal points, implicit_differentiable_renderer.py:126); tests override it."""
        return torch.rand(*shape, device=device, dtype=torch.float32)

    def _forward_train(self, input):
        ray_dirs = input['ray_dirs']
        device = ray_dirs.device
        batch_size, P, _ = ray_dirs.shape
        if self.render_last_pt:
            raise _lib.ArahError('render_last_pt=True is implemented for the eval render only (the training engine composites with 1 / n_steps)')
        r = self._prepare(input, training=True)
        sdf_p, skin_p, col_p = self._param_tensors(input)
        pose_cond = input.get('pose_cond', {})
        rt = self.ray_tracer
        # tracer, no gradient (:84-108)
        with torch.no_grad():
            u_all, u_near, u_far = rt.draw_jitter(P) if isinstance(rt, BodyRayTracing) else BodyRayTracing.draw_jitter(rt, P)
            r.train_trace(ray_dirs[0], input['body_bounds_intersections'][0], u_all, u_near, u_far)
            _, _, _, _, _, _, conv = r.trace_outputs(P, transforms=False, points=False)
        vol_mask = conv.any(-1)                                         # :148
        # view augmentation (:150-162)
        ray_augm = False
        view_orig = ray_dirs
        view = ray_dirs
        if 'view_noise' in pose_cond.keys():
            vn = pose_cond['view_noise']
            if vn is not None:
                if vn.size(-1) == 3 and vn.size(-2) == 3:
                    view = torch.matmul(vn, ray_dirs.transpose(1, 2)).transpose(1, 2)
                    ray_augm = True
                else:
                    view = ray_dirs + vn
            else:
                view = torch.zeros_like(ray_dirs)
        if ray_dirs.requires_grad or input['cam_loc'].requires_grad:
            # model.train_cameras = True: the reference lets d rgb / d view_dir flow into the camera extrinsics; no such gradient here
            raise _lib.ArahError('gradients with respect to ray_dirs / cam_loc (train_cameras) are not implemented')
        rgb, ws = _ShadeFn.apply(r, view[0].detach(), view_orig[0].detach(), ray_augm, self.train_skinning_net, *(sdf_p + skin_p + col_p))
        # regularisers (:73-78,117-140)
        out_extra = {}
        if 'points_skinning' in input.keys():
            out_extra['pred_weights'] = _SkinFn.apply(r, input['points_skinning'].reshape(-1, 3), *skin_p).view(
                input['points_skinning'].shape[0], -1, 24)
        if 'points_inside' in input.keys():
            out_extra['inside_sdf'] = _SdfFn.apply(r, 1, input['points_inside'].reshape(-1, 3), False, *sdf_p)
        n_eik = 1024
        eik = (self._rand_device((batch_size, n_eik, 3), device) - 0.5) * 2
        points_uniform = input['points_uniform'].reshape(-1, 3)
        points_all = torch.cat([eik.reshape(-1, 3), points_uniform], dim=0)
        sdf_all, grad_all = _SdfFn.apply(r, 0, points_all, True, *sdf_p)
        n_e, n_u = batch_size * n_eik, batch_size * 1024
        uniform_sdf = sdf_all[n_e:n_e + n_u, :].reshape(batch_size, 1024, 1)
        grad_eik = grad_all[:n_e, :]
        self._last = (r, P)
        output = {'rgb_values': rgb.unsqueeze(0), 'sdf_output': ws.unsqueeze(0), 'network_body_mask': vol_mask,
                  'body_mask': input['body_mask'], 'off_surface_mask': vol_mask, 'off_surface_sdf': uniform_sdf,
                  'grad_theta': grad_eik, 'surface_normals': None}
        output.update(out_extra)
        return output

    def forward(self, input):
        if self.training:
            return self._forward_train(input)
        r = self._prepare(input)
        P = input['ray_dirs'].shape[1]
        rgb, mask, pc = r.render(input['ray_dirs'][0], input['body_bounds_intersections'][0])
        self._last = (r, P)
        return {'points_cam': pc.unsqueeze(0), 'network_body_mask': mask.unsqueeze(0), 'rgb_values': rgb.unsqueeze(0)}

    def _trace_only(self, sdf_network, cam_loc, ray_directions, body_bounds_intersections, smpl_verts, skinning_weights,
                    bone_transforms, trans, coord_min, coord_max, center, train=False):
        inp = {'ray_dirs': ray_directions, 'cam_loc': cam_loc, 'pose': torch.eye(4, device=ray_directions.device).view(1, 4, 4),
               'body_bounds_intersections': body_bounds_intersections, 'smpl_verts': smpl_verts,
               'skinning_weights': skinning_weights, 'bone_transforms': bone_transforms, 'trans': trans,
               'coord_min': coord_min, 'coord_max': coord_max, 'center': center, 'sdf_network': sdf_network,
               'pose_cond': {'latent_code': torch.zeros(1, max(_effective_weight(self.rendering_network.lin0).shape[1] - _COLOR_BASE[_color_mode(self.rendering_network)], 0),
                                                        device=ray_directions.device)}}
        r = self._prepare(inp)
        P = ray_directions.shape[1]
        if train:
            u_all, u_near, u_far = self.ray_tracer.draw_jitter(P)
            r.train_trace(ray_directions[0], body_bounds_intersections[0], u_all, u_near, u_far)
        else:
            r.render(ray_directions[0], body_bounds_intersections[0])
        return r.trace_outputs(P)

    def extract_canonical_mesh(self, input, N=256):
        """MetaAvatarRender.forward(gen_cano_mesh=True), metaavatar_render/models/__init__.py:203-224: canonical mesh of the
        frame's SDF (normalised coordinates) and its posed vertices `forward_skinning(unnormalize(verts)) + trans`.
        Returns device tensors (verts [nv,3], faces [nf,3] int32, points_bar [nv,3])."""
        r = self._prepare(input)
        verts, faces = r.canonical_mesh(N)
        cmin, cmax = input['coord_min'].reshape(-1)[0], input['coord_max'].reshape(-1)[0]
        pts_hat = (verts / 2.0 + 0.5) * 1.1 * (cmax - cmin) + cmin - 0.05 * (cmax - cmin) + input['center'].reshape(1, 3)
        _, x_bar = r.eval_skin(pts_hat)
        return verts, faces, x_bar + input['trans'].reshape(1, 3)

    def render_normal_maps(self, input, N=256, image_size=(512, 512), mesh=None):
        """The rest of the `gen_cano_mesh` branch (metaavatar_render/models/__init__.py:226-309): rasterise the posed mesh with
        the frame's camera (`input['cam_rot'] [1,3,3]`, `input['cam_trans'] [1,3]`, `input['intrinsics'] [1,3,3]`, the keys the
        reference reads at :246-248) and the canonical mesh from the front / back -> {'output_normal', 'normal_cano_front',
        'normal_cano_back'}: [1, H, W, 3] device tensors in [0, 1], ready for `model_outputs.update(...)`.
        `mesh` = a previous `extract_canonical_mesh` result to reuse."""
        from .images import FrameImages
        verts, faces, points_bar = mesh if mesh is not None else self.extract_canonical_mesh(input, N)
        if getattr(self, '_frame_images', None) is None or self._frame_images.device != verts.device:
            self._frame_images = FrameImages(verts.device)
        H, W = image_size
        return self._frame_images.normal_maps(verts, faces, points_bar, input['cam_rot'].reshape(3, 3), input['cam_trans'].reshape(3),
                                              input['intrinsics'].reshape(3, 3), H, W)

    def tracer_outputs(self):
        """7-tuple of the tracer for the frame rendered by the last forward()."""
        r, P = self._last
        return r.trace_outputs(P)

    def stats(self):
        r, _ = self._last
        return r.stats()
