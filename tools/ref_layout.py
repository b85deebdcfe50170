"""Test / bench scaffolding: parameter containers with the reference's module / state_dict layout, WITHOUT any forward computation.

On the GPU box /root/reference does not exist, so tests and bench.py need something that *looks* like the modules the
reference hands to the renderer (attribute names, tensor shapes, weight-norm parametrisation) to exercise the drop-in
classes in renderer.py.  They deliberately cannot compute: the only compute path is the CUDA library.

    sdf_network            nn.Sequential(Sequential(BatchLinearFiLM, Sine) x 6, BatchLinear)   hyperlayers.py:270-285
    skinning_model         SkinningModel(skinning_decoder_fwd=Deformer)                        metaavatar_render/models/skinning_model.py:3
    rendering_network      RenderingNetwork (lin0..lin5, weight_g / weight_v / bias)            metaavatar_render/models/decoder.py:10
    deviation_network      SingleVarianceNetwork (.variance)                                   metaavatar_render/models/decoder.py:127
"""
import numpy as np
import torch
import torch.nn as nn

from arah_release_b200.containers import BatchLinear, BatchLinearFiLM, Sine, _NoForward  # noqa: F401


class WNLinear(_NoForward):
    """nn.utils.weight_norm(nn.Linear) parameter layout: weight_g [out,1], weight_v [out,in], bias [out]."""
    def __init__(self, v, g, b):
        super().__init__()
        self.weight_g = nn.Parameter(torch.as_tensor(g).clone().float().view(-1, 1))
        self.weight_v = nn.Parameter(torch.as_tensor(v).clone().float())
        self.bias = nn.Parameter(torch.as_tensor(b).clone().float())


class Deformer(_NoForward):
    def __init__(self, layers):
        super().__init__()
        for i, L in enumerate(layers):
            setattr(self, f'lin{i}', WNLinear(L['v'], L['g'], L['b']))
        self.num_layers = len(layers) + 1


class SkinningModel(_NoForward):
    def __init__(self, skinning_decoder_fwd):
        super().__init__()
        self.skinning_decoder_fwd = skinning_decoder_fwd


class RenderingNetwork(_NoForward):
    def __init__(self, layers, mode='idr', skips=(3,), pose_encoder='latent', multires_view=4):
        super().__init__()
        for i, L in enumerate(layers):
            setattr(self, f'lin{i}', WNLinear(L['v'], L['g'], L['b']))
        self.num_layers = len(layers) + 1
        self.mode = mode
        self.skips = list(skips)
        self.pose_encoder_type = pose_encoder
        self.embedview_fn = object() if multires_view > 0 else None
        self.embed_fn = None
        self.squeeze_out = True
        self.multires_view = multires_view


class SingleVarianceNetwork(_NoForward):
    def __init__(self, init_val):
        super().__init__()
        self.register_parameter('variance', nn.Parameter(torch.tensor(float(init_val))))


def sdf_network_from_frame(frame, device):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(device)
    layers = []
    for i in range(6):
        layers.append(nn.Sequential(BatchLinearFiLM(t(frame.sdf['W'][i]).unsqueeze(0), t(frame.sdf['b'][i]).view(1, 1, -1),
                                                    t(frame.sdf['freq'][i]).view(1, -1), t(frame.sdf['phase'][i]).view(1, -1)), Sine()))
    layers.append(BatchLinear(t(frame.sdf['W'][6]).unsqueeze(0), t(frame.sdf['b'][6]).view(1, 1, -1)))
    return nn.Sequential(*layers)


def modules_from_frame(frame, device):
    """(deviation_network, rendering_network, skinning_model, sdf_network) in the reference layout."""
    dev = SingleVarianceNetwork(float(frame.beta)).to(device)
    mode = getattr(frame, 'color_mode', 'idr')
    rend = RenderingNetwork(frame.color, mode=mode, multires_view=0 if mode == 'no_view_dir' else 4).to(device)
    skin = SkinningModel(Deformer(frame.skin)).to(device)
    return dev, rend, skin, sdf_network_from_frame(frame, device)


def inputs_from_frame(frame, sdf_network, device):
    """The `input` dict IDHRNetwork.forward reads (implicit_differentiable_renderer.py:52-71), batch size 1."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(device)
    P = frame.P
    return {
        'ray_dirs': t(frame.ray_dirs).view(1, P, 3), 'cam_loc': t(frame.cam_loc).view(1, 3), 'pose': t(frame.pose).view(1, 4, 4),
        'body_mask': torch.ones(1, P, dtype=torch.bool, device=device),
        'body_bounds_intersections': t(frame.near_far).view(1, P, 2),
        'loc': torch.zeros(1, 1, 3, device=device), 'sc_factor': torch.ones(1, 1, 1, device=device),
        'smpl_verts': t(frame.smpl_verts).view(1, -1, 3), 'skinning_weights': t(frame.smpl_weights).view(1, -1, 24),
        'vol_feat': torch.empty(1, 0, device=device), 'bone_transforms': t(frame.bone_transforms).view(1, 24, 4, 4),
        'trans': t(frame.trans).view(1, 1, 3), 'coord_min': t(np.array([frame.coord_min])).view(1, 1, 1),
        'coord_max': t(np.array([frame.coord_max])).view(1, 1, 1), 'center': t(frame.center).view(1, 1, 3),
        'minimal_shape': t(frame.minimal_shape).view(1, -1, 3), 'sdf_network': sdf_network,
        'pose_cond': {'latent_code': t(frame.latent).view(1, -1)},
        # host copies of the per-frame scalars (the dataset has them before .cuda()): lets the renderer skip the device read-back
        'host_scalars': {'trans': [float(v) for v in np.asarray(frame.trans).reshape(-1)[:3]], 'coord_min': float(frame.coord_min),
                         'coord_max': float(frame.coord_max), 'center': [float(v) for v in np.asarray(frame.center).reshape(-1)[:3]],
                         'cam_loc': [float(v) for v in np.asarray(frame.cam_loc).reshape(-1)[:3]],
                         'pose': [float(v) for v in np.asarray(frame.pose).reshape(-1)[:16]]},
    }
