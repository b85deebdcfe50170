"""Drop-in for the reference's training criterion ("fused loss reductions").

    from arah_release_b200.loss import IDHRLoss            # was: from im2mesh.metaavatar_render.renderer.loss import IDHRLoss
    criteria = IDHRLoss(rgb_weight=..., perceptual_weight=0.0, eikonal_weight=..., mask_weight=..., off_surface_weight=...,
                        inside_weight=..., params_weight=..., skinning_weight=..., rgb_loss_type='l1')     # lightning_model.py:127-136
    loss_dict = criteria(model_outputs, ground_truth)      # same keys as renderer/loss.py:190-200; loss_dict['loss'].backward()

Same constructor, same `forward(model_outputs, ground_truth)`, same nine result keys and — like the reference — a `[1]`-shaped
`loss` whenever a disabled term contributes its `torch.zeros(1)`.  The nine terms and d loss / d input of every differentiable
input come from ONE fused C-ABI call (`arah_idhr_loss`, csrc/arah_loss.cu: four launches, no host synchronisation) instead of
~60 small torch kernels and three `.sum() == 0` / `.max() > 1` host round trips; `backward` only scales the stored gradients.
CUDA only (no CPU path).  Not supported: the perceptual (LPIPS) term — a VGG network outside this path; `perceptual_weight > 0`
raises.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib
from ._lib import check

TERMS = ('loss', 'rgb_loss', 'perceptual_loss', 'eikonal_loss', 'mask_loss', 'off_surface_loss', 'inside_loss', 'sdf_params_loss', 'skinning_loss')
_RGB_TYPES = {'l1': 0, 'mse': 1, 'smoothed_l1': 2}
_MAX_RAYS = 2048                                   # renderer/loss.py:124-127,132: every per-ray tensor is cut to the first 2048 rays


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else None


class _LossFn(torch.autograd.Function):
    """inputs: (owner, aux dict, rgb_values, sdf_output, grad_theta, off_surface_sdf, inside_sdf, pred_weights, *sdf_params) -> terms [9]."""

    @staticmethod
    def forward(ctx, owner, aux, *tensors):
        dev = aux['device']
        f32 = lambda t: None if t is None else t.detach().to(dev, torch.float32).contiguous()
        rgb, sdf_out, gth, off_sdf, ins_sdf, pw = (f32(t) for t in tensors[:6])
        params = [f32(p).reshape(-1) for p in tensors[6:]]
        if len(params) > 8:
            raise _lib.ArahError('at most 8 sdf_params tensors')
        N = aux['n_rays']
        cfg, inp, g = _lib.ArahLossConfig(), _lib.ArahLossInputs(), _lib.ArahLossGrads()
        for k, v in owner.weights().items():
            setattr(cfg, k, float(v))
        cfg.rgb_loss_type = _RGB_TYPES[owner.rgb_loss_type]
        inp.rgb_values, inp.rgb_gt = _ptr(rgb), _ptr(aux['rgb_gt'])
        inp.network_body_mask, inp.body_mask, inp.off_surface_mask = _ptr(aux['network_body_mask']), _ptr(aux['body_mask']), _ptr(aux['off_surface_mask'])
        inp.sdf_output, inp.grad_theta, inp.off_surface_sdf, inp.inside_sdf = _ptr(sdf_out), _ptr(gth), _ptr(off_sdf), _ptr(ins_sdf)
        inp.pred_weights, inp.sampled_weights = _ptr(pw), _ptr(aux['sampled_weights'])
        inp.n_rays = N
        inp.n_eikonal = 0 if gth is None else gth.numel() // 3
        inp.n_off = 0 if off_sdf is None else off_sdf.numel()
        inp.n_inside = 0 if ins_sdf is None else ins_sdf.numel()
        inp.n_joints = 0 if pw is None else pw.shape[-1]
        inp.n_skin = 0 if pw is None else pw.numel() // max(pw.shape[-1], 1)
        inp.n_param_tensors = len(params)
        grads = [None if t is None else torch.empty_like(t) for t in (rgb, sdf_out, gth, off_sdf, ins_sdf, pw)]
        pgrads = [torch.empty_like(p) for p in params]
        for name, t in zip(('rgb_values', 'sdf_output', 'grad_theta', 'off_surface_sdf', 'inside_sdf', 'pred_weights'), grads):
            setattr(g, name, _ptr(t))
        for i, (p, pg) in enumerate(zip(params, pgrads)):
            inp.sdf_params[i], inp.sdf_params_count[i], g.sdf_params[i] = _ptr(p), p.numel(), _ptr(pg)
        terms = torch.empty(9, device=dev)
        ws = owner._workspace(dev)
        check(_lib.lib().arah_idhr_loss(C.byref(cfg), C.byref(inp), _ptr(terms), C.byref(g), _ptr(ws), ws.numel(), owner._stream(dev)))
        ctx.grads, ctx.pgrads, ctx.shapes = grads, pgrads, [None if t is None else t.shape for t in tensors]
        ctx.keep = (rgb, sdf_out, gth, off_sdf, ins_sdf, pw, params, aux)
        return terms

    @staticmethod
    def backward(ctx, g_terms):
        s = g_terms[0]                               # only `loss` (terms[0]) carries the graph; the stored gradients are d loss / d input
        out = [None, None]
        for t, shp in zip(list(ctx.grads) + list(ctx.pgrads), ctx.shapes):
            out.append(None if t is None or shp is None else (t * s).reshape(shp))
        return tuple(out)


class IDHRLoss(nn.Module):
    """Mirror of im2mesh/metaavatar_render/renderer/loss.py::IDHRLoss (constructor :9-44, forward :122-200)."""

    def __init__(self, rgb_weight, perceptual_weight, eikonal_weight, mask_weight, off_surface_weight, inside_weight, params_weight, skinning_weight,
                 rgb_loss_type='l1', perceptual_loss_fn=None):
        super().__init__()
        self.rgb_weight, self.perceptual_weight, self.eikonal_weight, self.mask_weight = rgb_weight, perceptual_weight, eikonal_weight, mask_weight
        self.off_surface_weight, self.params_weight, self.skinning_weight, self.inside_weight = off_surface_weight, params_weight, skinning_weight, inside_weight
        if rgb_loss_type not in _RGB_TYPES:
            raise ValueError('Unsupported RGB loss type: {}. Only l1, smoothed_l1 and mse are supported'.format(rgb_loss_type))
        self.rgb_loss_type = rgb_loss_type
        self.p_loss = perceptual_loss_fn
        self._ws = {}

    def weights(self):
        return {k: getattr(self, k) for k in ('rgb_weight', 'perceptual_weight', 'eikonal_weight', 'mask_weight', 'off_surface_weight', 'inside_weight',
                                              'params_weight', 'skinning_weight')}

    def _workspace(self, dev):
        ws = self._ws.get(dev)
        if ws is None:
            ws = self._ws[dev] = torch.empty(int(_lib.lib().arah_idhr_loss_workspace()), dtype=torch.uint8, device=dev)
        return ws

    def _stream(self, dev):
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    @staticmethod
    def _require_cuda(dev):
        if dev.type != 'cuda':
            raise _lib.ArahError('IDHRLoss only exists as CUDA kernels; got device %s' % dev)

    def forward(self, model_outputs, ground_truth):
        rgb_all = model_outputs['rgb_values']
        dev = rgb_all.device
        self._require_cuda(dev)
        if self.perceptual_weight > 0:
            raise _lib.ArahError('the perceptual (LPIPS) term is not part of this path: perceptual_weight must be 0')
        if rgb_all.shape[0] != 1:
            raise _lib.ArahError('one frame per batch (the renderer asserts the same, ray_tracing.py:129-132)')
        u8 = lambda t: t[0, :_MAX_RAYS].to(dev).reshape(-1).to(torch.uint8).contiguous()
        body = u8(model_outputs['body_mask'])
        N = int(body.numel())
        aux = {'device': dev, 'n_rays': N, 'body_mask': body, 'network_body_mask': u8(model_outputs['network_body_mask']),
               'off_surface_mask': u8(model_outputs['off_surface_mask']), 'rgb_gt': None, 'sampled_weights': None}
        on = lambda w: w > 0
        # the kernels index every per-ray array up to n_rays: a row-count mismatch (an indexing error in the reference) must not
        # become an out-of-bounds device access here
        for name in ('network_body_mask', 'off_surface_mask'):
            if int(aux[name].numel()) != N:
                raise _lib.ArahError('%s has %d rays but body_mask %d' % (name, int(aux[name].numel()), N))
        rgb = sdf_out = gth = off_sdf = ins_sdf = pw = None
        if on(self.rgb_weight):
            rgb = rgb_all[0, :_MAX_RAYS]
            aux['rgb_gt'] = ground_truth['rgb'][0, :_MAX_RAYS].detach().to(dev, torch.float32).contiguous()
            if rgb.shape[0] != N or rgb.shape[-1] != 3 or tuple(aux['rgb_gt'].shape) != tuple(rgb.shape):
                raise _lib.ArahError('rgb term: rgb_values %s / rgb ground truth %s do not match the %d rays of body_mask'
                                     % (tuple(rgb.shape), tuple(aux['rgb_gt'].shape), N))
        if on(self.mask_weight):
            sdf_out = model_outputs['sdf_output'][0].reshape(-1)                # NOT cut to 2048 by the reference (:143)
            if sdf_out.numel() != N:
                raise _lib.ArahError('mask term: sdf_output has %d rays but the masks %d (the reference fails the same way, loss.py:100)' % (sdf_out.numel(), N))
        if on(self.eikonal_weight):
            gth = model_outputs['grad_theta'].reshape(-1, 3)
        if on(self.off_surface_weight):
            off_sdf = model_outputs['off_surface_sdf'].reshape(-1)
        if on(self.inside_weight):
            ins_sdf = model_outputs['inside_sdf'].reshape(-1)
        params = []
        if on(self.params_weight):
            params = [p.reshape(-1) for p in model_outputs['sdf_params']]
            if any(p.shape[0] != 1 for p in model_outputs['sdf_params']):
                raise _lib.ArahError('sdf_params: batch size 1 expected')
        if on(self.skinning_weight):
            pw = model_outputs['pred_weights']
            pw = pw.reshape(-1, pw.shape[-1])
            aux['sampled_weights'] = ground_truth['sampled_weights'].detach().to(dev, torch.float32).reshape(-1, pw.shape[-1]).contiguous()
            if aux['sampled_weights'].shape != pw.shape:
                raise _lib.ArahError('pred_weights / sampled_weights shape mismatch')
        terms = _LossFn.apply(self, aux, rgb, sdf_out, gth, off_sdf, ins_sdf, pw, *params)
        out = {}
        any_off = False
        for i, k in enumerate(TERMS):
            w = 1.0 if k == 'loss' else getattr(self, k.replace('_loss', '_weight').replace('sdf_params', 'params'))
            if k != 'loss' and not on(w):
                out[k] = torch.zeros(1, device=dev)                              # loss.py:135-170
                any_off = True
            else:
                out[k] = terms[i] if k == 'loss' else terms[i].detach()
        if any_off:
            out['loss'] = out['loss'].reshape(1)                                 # scalar + zeros(1) broadcasts to [1] in the reference
        return out
