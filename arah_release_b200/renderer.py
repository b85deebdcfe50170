"""Drop-in host side of the B200-native ARAH renderer.

`IDHRNetwork` and `BodyRayTracing` keep the reference's constructor signatures, attribute names (so state_dict keys and
the aliasing `color_decoder.* == idhr_network.rendering_network.*` survive strict checkpoint loads) and return
signatures:

    IDHRNetwork      /root/reference/im2mesh/metaavatar_render/renderer/implicit_differentiable_renderer.py:15-259
    BodyRayTracing   /root/reference/im2mesh/metaavatar_render/renderer/ray_tracing.py:13-172

Their eval `forward` packs the modules' weights + the per-frame buffers into an ArahFrame and calls the C ABI
(include/arah_b200.h) on torch's current CUDA stream.  There is no PyTorch / CPU implementation of the path here: if the
CUDA library is missing or the tensors are not on a CUDA device, the call raises.

Training (`self.training == True`) raises NotImplementedError: backward through root finding is SURVEY.md §8(f) row f2.
"""
import ctypes as C
import os
import weakref

import torch
import torch.nn as nn

from . import _lib
from ._lib import ArahConfig, ArahFrame, ArahStats, check

N_VERTS_DEFAULT = 6890


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


class ArahRenderer:
    """Thin RAII wrapper around an ArahHandle (one per device/stream; not thread-safe)."""

    def __init__(self, device, n_steps=64, near_samples=16, far_samples=16, cano_view_dirs=True, latent_dim=128,
                 n_verts=N_VERTS_DEFAULT, max_rays=65536, shade_mode=None, root_mode=None):
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
                              n_verts=n_verts, max_rays=max_rays, shade_mode=mode, root_mode=rmode)
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
        """Read weights out of modules laid out like the reference's (see ref_layout.py) + the input dict."""
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
        if getattr(rn, 'mode', 'idr') != 'idr' or list(getattr(rn, 'skips', [3])) != [3] or getattr(rn, 'embedview_fn', 1) is None:
            raise _lib.ArahError("only renderer mode 'idr' with multires_view=4, skips=[3] is implemented")
        pe = getattr(rn, 'pose_encoder_type', None)
        col_W = [_effective_weight(getattr(rn, f'lin{i}')) for i in range(6)]
        col_b = [getattr(rn, f'lin{i}').bias for i in range(6)]
        d_in = col_W[0].shape[1]
        latent_dim = d_in - 289
        if latent_dim != self.cfg.latent_dim:
            raise _lib.ArahError(f'colour net expects a {latent_dim}-d pose feature, handle was built for {self.cfg.latent_dim}')
        latent = None
        if latent_dim > 0:
            if pe != 'latent':
                raise _lib.ArahError("only color_pose_encoder 'latent' (or None) is implemented")
            latent = inputs['pose_cond']['latent_code'].reshape(-1)
        beta = float(torch.linalg.norm(deviation_network.variance.detach()).item())
        cam_loc = inputs['cam_loc'].reshape(-1, 3)[0].tolist()
        self.set_frame(sdf_W=sdf_W, sdf_b=sdf_b, sdf_freq=torch.stack(freq), sdf_phase=torch.stack(phase), skin_W=skin_W,
                       skin_b=skin_b, col_W=col_W, col_b=col_b, latent=latent, beta=beta,
                       bone_transforms=inputs['bone_transforms'][0], smpl_verts=inputs['smpl_verts'][0],
                       smpl_weights=inputs['skinning_weights'][0], trans=inputs['trans'].reshape(-1)[:3].tolist(),
                       coord_min=float(inputs['coord_min'].reshape(-1)[0]), coord_max=float(inputs['coord_max'].reshape(-1)[0]),
                       center=inputs['center'].reshape(-1)[:3].tolist(), cam_loc=cam_loc,
                       pose=inputs['pose'].reshape(-1, 16)[0].tolist(), pose_on_host=pose_on_host)

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
        rgb = torch.empty(P, 3, dtype=torch.float32).pin_memory() if rgb is None else rgb
        mask = torch.empty(P, dtype=torch.uint8).pin_memory() if mask is None else mask
        pc = torch.empty(P, 3, dtype=torch.float32).pin_memory() if points_cam is None else points_cam
        check(_lib.lib().arah_render_host(self._h, _ptr(rd), _ptr(nf), P, _ptr(rgb), _ptr(mask), _ptr(pc), self.stream))
        return rgb, mask, pc

    def trace_outputs(self, P, transforms=True):
        """BodyRayTracing.forward's 7-tuple for the last render (ray_tracing.py:166-172), batch dim added."""
        S, dev = self.n_steps, self.device
        pts_hat = torch.empty(P, 3, device=dev); m = torch.empty(P, device=dev, dtype=torch.uint8); dists = torch.empty(P, device=dev)
        sp = torch.empty(P, S, 3, device=dev); sd = torch.empty(P, S, device=dev)
        sT = torch.empty(P, S, 4, 4, device=dev) if transforms else None
        sc = torch.empty(P, S, device=dev, dtype=torch.uint8)
        check(_lib.lib().arah_get_trace(self._h, _ptr(pts_hat), _ptr(m), _ptr(dists), _ptr(sp), _ptr(sd),
                                        _ptr(sT) if sT is not None else None, _ptr(sc), self.stream))
        return (pts_hat.unsqueeze(0), m.bool().unsqueeze(0), dists.unsqueeze(0), sp.unsqueeze(0), sd.unsqueeze(0),
                sT.unsqueeze(0) if sT is not None else None, sc.bool().unsqueeze(0))

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
        if not eval_mode:
            raise NotImplementedError('training-mode ray tracing (stochastic z perturbation) is SURVEY.md §8 row f2')
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise _lib.ArahError('BodyRayTracing must be owned by an IDHRNetwork (it shares its renderer handle)')
        return owner._trace_only(sdf_network, cam_loc, ray_directions, body_bounds_intersections, smpl_verts,
                                       skinning_weights, bone_transforms, trans, coord_min, coord_max, center)


class IDHRNetwork(nn.Module):
    """Implicit Differentiable Human Renderer — same constructor and attributes as the reference
    (implicit_differentiable_renderer.py:18-40); eval forward runs entirely in libarah_b200.so."""

    def __init__(self, deviation_network, rendering_network, skinning_model, ray_tracer, cano_view_dirs=True,
                 train_skinning_net=False, render_last_pt=False, low_vram=False, shade_mode=None, root_mode=None):
        super().__init__()
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
        if render_last_pt:
            raise _lib.ArahError('render_last_pt=True is not implemented (no shipped config sets it, configs/default.yaml:52)')
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
                             latent_dim=latent_dim, n_verts=n_verts, shade_mode=self.shade_mode, root_mode=self.root_mode)
            self._renderers[key] = r
        return r

    def _prepare(self, input):
        ray_dirs = input['ray_dirs']
        if ray_dirs.device.type != 'cuda':
            raise _lib.ArahError('IDHRNetwork (B200) needs CUDA tensors; there is no CPU fallback')
        if ray_dirs.shape[0] != 1:
            raise _lib.ArahError('one frame per call (the reference assumes the same, ray_tracing.py:129-132)')
        latent_dim = _effective_weight(self.rendering_network.lin0).shape[1] - 289
        r = self._renderer(ray_dirs.device, latent_dim, input['smpl_verts'].shape[1])
        r.set_frame_from_modules(input['sdf_network'], self.skinning_model, self.rendering_network, self.deviation_network, input)
        return r

    def forward(self, input):
        if self.training:
            raise NotImplementedError('training forward/backward through root finding is SURVEY.md §8 row f2; '
                                      'call .eval() for rendering')
        r = self._prepare(input)
        P = input['ray_dirs'].shape[1]
        rgb, mask, pc = r.render(input['ray_dirs'][0], input['body_bounds_intersections'][0])
        self._last = (r, P)
        return {'points_cam': pc.unsqueeze(0), 'network_body_mask': mask.unsqueeze(0), 'rgb_values': rgb.unsqueeze(0)}

    def _trace_only(self, sdf_network, cam_loc, ray_directions, body_bounds_intersections, smpl_verts, skinning_weights,
                    bone_transforms, trans, coord_min, coord_max, center):
        inp = {'ray_dirs': ray_directions, 'cam_loc': cam_loc, 'pose': torch.eye(4, device=ray_directions.device).view(1, 4, 4),
               'body_bounds_intersections': body_bounds_intersections, 'smpl_verts': smpl_verts,
               'skinning_weights': skinning_weights, 'bone_transforms': bone_transforms, 'trans': trans,
               'coord_min': coord_min, 'coord_max': coord_max, 'center': center, 'sdf_network': sdf_network,
               'pose_cond': {'latent_code': torch.zeros(1, max(_effective_weight(self.rendering_network.lin0).shape[1] - 289, 0),
                                                        device=ray_directions.device)}}
        r = self._prepare(inp)
        P = ray_directions.shape[1]
        r.render(ray_directions[0], body_bounds_intersections[0])
        return r.trace_outputs(P)

    def tracer_outputs(self):
        """7-tuple of the tracer for the frame rendered by the last forward()."""
        r, P = self._last
        return r.trace_outputs(P)

    def stats(self):
        r, _ = self._last
        return r.stats()
