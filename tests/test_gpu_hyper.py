"""GPU parity for the hypernetwork row: arah_hyper_forward through HyperSDFDecoder against the reference's own
outputs (tests/golden/hyper_s*.npz) and the numpy oracle; then the full chain hypernetwork -> renderer on the device.
Floating point (fp32 sums of <= 288 products in a different order than MKL): tolerance 2e-5 absolute on O(1) outputs."""
import numpy as np
import pytest
import torch

import helpers_hyper as hh

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _decoder(seed, rel):
    from arah_release_b200 import synthetic as syn
    from arah_release_b200.hypernet import HyperSDFDecoder
    sd = syn.make_hypernet_state_dict(seed)
    return sd, HyperSDFDecoder({k: torch.from_numpy(v) for k, v in sd.items()}, DEV, rel_joints=rel)


def _as_out(res):
    dec = res['decoder']
    return {'W': [(dec[l][0] if l < 6 else dec[l]).weights[0].cpu().numpy() for l in range(7)],
            'b': [(dec[l][0] if l < 6 else dec[l]).biases.reshape(-1).cpu().numpy() for l in range(7)],
            'freq': torch.stack([dec[l][0].freq[0] for l in range(6)]).cpu().numpy(),
            'phase': torch.stack([dec[l][0].phase_shift[0] for l in range(6)]).cpu().numpy()}


@pytest.mark.parametrize('seed', [0, 1])
def test_hypernetwork_matches_reference_and_oracle(seed):
    from arah_release_b200 import synthetic as syn
    from oracle import hyper_oracle as ho
    gold = hh.load(seed)
    rel = bool(int(gold['rel_joints']))
    sd, dec = _decoder(seed, rel)
    rots, Jtrs, latent = syn.make_hypernet_inputs(seed)
    t = lambda a: torch.from_numpy(a).to(DEV)
    res = dec({'coords': torch.zeros(1, 1, 3, device=DEV), 'rots': t(rots), 'Jtrs': t(Jtrs), 'latent': t(latent)})
    torch.cuda.synchronize()
    out = _as_out(res)
    worst = hh.check(out, gold, seed, label=f'cuda s{seed}')
    o = ho.forward(sd, rots, Jtrs, latent, rel_joints=rel)
    full = max(max(np.abs(out['W'][l] - o['W'][l]).max() for l in range(7)), max(np.abs(out['b'][l] - o['b'][l]).max() for l in range(7)))
    print(f'hypernet s{seed}: worst |cuda - reference| {worst:.2e}, worst |cuda - oracle| over all 1.3 M outputs {full:.2e}')
    assert full <= hh.ATOL
    assert [p.numel() for p in res['params']] == [int(n) for n in gold['params_numel']]
    # rots_noise is added to rots (siren_modules.py:289-290); no latent == zero latent
    noise = (0.01 * np.random.default_rng(5).standard_normal(rots.shape)).astype(np.float32)
    res2 = dec({'rots': t(rots), 'Jtrs': t(Jtrs), 'latent': t(latent), 'rots_noise': t(noise)})
    o2 = ho.forward(sd, rots + noise, Jtrs, latent, rel_joints=rel)
    assert np.abs(_as_out(res2)['W'][2] - o2['W'][2]).max() <= hh.ATOL
    res3 = dec({'rots': t(rots), 'Jtrs': t(Jtrs)})
    o3 = ho.forward(sd, rots, Jtrs, None, rel_joints=rel)
    assert np.abs(_as_out(res3)['freq'] - o3['freq']).max() <= hh.ATOL


def test_hypernetwork_feeds_the_renderer():
    """decoder from the hypernetwork kernel -> IDHRNetwork on the device == the same SDF parameters uploaded from the host."""
    from arah_release_b200 import synthetic as syn
    from tools import ref_layout as rl
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    fr = syn.make_frame(16, 16, seed=3)
    sd, dec = _decoder(0, False)
    rots, Jtrs, latent = syn.make_hypernet_inputs(0)
    t = lambda a: torch.from_numpy(a).to(DEV)
    res = dec({'rots': t(rots), 'Jtrs': t(Jtrs), 'latent': t(latent)})
    dvn, rend, skin, _ = rl.modules_from_frame(fr, DEV)
    net = IDHRNetwork(dvn, rend, skin, BodyRayTracing(n_steps=fr.n_steps), cano_view_dirs=fr.cano_view_dirs).eval()
    out_a = net(rl.inputs_from_frame(fr, res['decoder'], DEV))['rgb_values'].clone()
    o = _as_out(res)
    fr.sdf = {'W': o['W'], 'b': o['b'], 'freq': o['freq'], 'phase': o['phase']}
    out_b = net(rl.inputs_from_frame(fr, rl.sdf_network_from_frame(fr, DEV), DEV))['rgb_values']
    torch.cuda.synchronize()
    assert torch.equal(out_a, out_b)
