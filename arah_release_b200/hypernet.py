"""Host mirror of the MetaAvatar hypernetwork forward.

`HyperSDFDecoder` stands where `MetaAvatarRender.sdf_decoder` (a `HyperBVPNet`, metaavatar/models/siren_modules.py:247-312)
stands in the reference: it is called with the same `decoder_input` dict (`rots`, `Jtrs`, optional `latent`, `rots_noise`;
metaavatar_render/models/__init__.py:152-183) and returns a dict with the same `decoder` / `params` entries, the decoder being
the nn.Sequential(Sequential(BatchLinearFiLM, Sine) x 6, BatchLinear) the renderer reads its per-frame SDF weights from.
The arithmetic runs in libarah_b200.so (arah_hyper_forward: two launches, one HBM pass over the 341 MB of output matrices);
the parameters stay where the caller keeps them (a state_dict with the reference's key names) and are read in place.

Inference only: training differentiates through the hypernetwork and keeps using the reference module for that.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import ArahHyperWeights, ArahSdfParams, check
from . import containers as rl

IN_CH = [3, 256, 256, 256, 256, 256, 256]
OUT_CH = [256, 256, 256, 256, 256, 256, 1]


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class HyperSDFDecoder(nn.Module):
    """Drop-in for HyperBVPNet(in_features=3, num_hidden_layers=5, hierarchical_pose=True, hyper_in_ch=144, use_FiLM=True)."""

    def __init__(self, state_dict, device, rel_joints=False):
        super().__init__()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.ArahError('the hypernetwork forward only exists as CUDA kernels; got device %s' % device)
        sd = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in state_dict.items()}
        g = lambda k: sd[k]
        # the 24 per-joint encoders are stacked so that the kernel walks one array (siren_modules.py:208-214)
        self._pe = [torch.stack([g(f'pose_encoder.layers.{j}.0.weight') for j in range(24)]).contiguous(),
                    torch.stack([g(f'pose_encoder.layers.{j}.0.bias') for j in range(24)]).contiguous(),
                    torch.stack([g(f'pose_encoder.layers.{j}.2.weight') for j in range(24)]).contiguous(),
                    torch.stack([g(f'pose_encoder.layers.{j}.2.bias') for j in range(24)]).contiguous()]
        self._sd = sd
        w = ArahHyperWeights()
        w.pe_l0_W, w.pe_l0_b = _ptr(g('pose_encoder.layer_0.weight')), _ptr(g('pose_encoder.layer_0.bias'))
        w.pe_W1, w.pe_b1, w.pe_W2, w.pe_b2 = (_ptr(t) for t in self._pe)
        for i, n in enumerate((0, 2, 4, 6)):
            w.map_W[i] = _ptr(g(f'net.mapping_network.network.{n}.weight'))
            w.map_b[i] = _ptr(g(f'net.mapping_network.network.{n}.bias'))
        for l in range(7):
            pre = f'net.layers.{l}.hyper_linear.' if l < 6 else f'net.layers.{l}.'       # HyperLayerFiLM wraps a HyperLinearFiLM
            fc = pre + 'hypo_params.net.'
            n_l = IN_CH[l] * OUT_CH[l] + OUT_CH[l]
            assert tuple(g(fc + '2.weight').shape) == (n_l, 256), (l, g(fc + '2.weight').shape)
            w.fc1_W[l], w.fc1_b[l] = _ptr(g(fc + '0.net.0.weight')), _ptr(g(fc + '0.net.0.bias'))
            w.ln1_g[l], w.ln1_b[l] = _ptr(g(fc + '0.net.1.weight')), _ptr(g(fc + '0.net.1.bias'))
            w.fc2_W[l], w.fc2_b[l] = _ptr(g(fc + '1.net.0.weight')), _ptr(g(fc + '1.net.0.bias'))
            w.ln2_g[l], w.ln2_b[l] = _ptr(g(fc + '1.net.1.weight')), _ptr(g(fc + '1.net.1.bias'))
            w.out_W[l], w.out_b[l] = _ptr(g(fc + '2.weight')), _ptr(g(fc + '2.bias'))
            w.init[l] = _ptr(g(pre + 'hypo_params_init'))
        w.rel_joints = int(bool(rel_joints))
        self._w = w
        self._ws = torch.empty(int(_lib.lib().arah_hyper_workspace()), dtype=torch.uint8, device=self.device)
        self.weight_bytes = sum(v.numel() * 4 for v in sd.values())

    def launch_raw(self, rots, Jtrs, latent, out_struct):
        """The bare C-ABI call with prepared device buffers (bench.py: device time without the tensor bookkeeping of forward)."""
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        check(_lib.lib().arah_hyper_forward(C.byref(self._w), _ptr(rots), _ptr(Jtrs), _ptr(latent), C.byref(out_struct), _ptr(self._ws), stream))

    def forward(self, model_input):
        rots = model_input['rots']
        if 'rots_noise' in model_input:                                   # siren_modules.py:289-290
            rots = rots + model_input['rots_noise']
        rots = rots.to(self.device, torch.float32).reshape(-1).contiguous()
        Jtrs = model_input['Jtrs'].to(self.device, torch.float32).reshape(-1).contiguous()
        if rots.numel() != 216 or Jtrs.numel() != 72:
            raise _lib.ArahError('one frame per call: rots [1,24,9], Jtrs [1,24,3]')
        latent = model_input.get('latent')
        if latent is not None:
            latent = latent.to(self.device, torch.float32).reshape(-1).contiguous()
        dev = self.device
        # one allocation per call, carved into the reference's tensors (each slice starts on a 16-byte boundary)
        sizes = [OUT_CH[l] * IN_CH[l] for l in range(7)] + [(OUT_CH[l] + 3) // 4 * 4 for l in range(7)] + [1536, 1536]
        flat = torch.empty(sum(sizes), device=dev)
        parts = torch.split(flat, sizes)
        W = [parts[l].view(1, OUT_CH[l], IN_CH[l]) for l in range(7)]
        b = [parts[7 + l][:OUT_CH[l]].view(1, 1, OUT_CH[l]) for l in range(7)]
        freq, phase = parts[14].view(6, 256), parts[15].view(6, 256)
        out = ArahSdfParams()
        for l in range(7):
            out.sdf_W[l], out.sdf_b[l] = _ptr(W[l]), _ptr(b[l])
        out.sdf_freq, out.sdf_phase = _ptr(freq), _ptr(phase)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(_lib.lib().arah_hyper_forward(C.byref(self._w), _ptr(rots), _ptr(Jtrs), _ptr(latent) if latent is not None else None,
                                            C.byref(out), _ptr(self._ws), stream))
        self._keep = (rots, Jtrs, latent)
        layers = [nn.Sequential(rl.BatchLinearFiLM(W[l], b[l], freq[l].view(1, -1), phase[l].view(1, -1)), rl.Sine()) for l in range(6)]
        layers.append(rl.BatchLinear(W[6], b[6]))
        decoder = nn.Sequential(*layers)
        params = [W[l].view(1, -1) for l in range(7)]                     # siren_modules.py:306-310
        return {'model_in': model_input.get('coords'), 'model_out': None, 'params': params, 'decoder': decoder}
