"""TEST INFRASTRUCTURE — golden vectors for the hypernetwork row (f4) from the UNMODIFIED reference.

    python oracle/gen_golden_hyper.py          # build container only (/root/reference must exist)

Loads arah_release_b200.synthetic.make_hypernet_state_dict(seed) into the reference's own HyperBVPNet with strict=True (so
key names and shapes are exactly the reference's), runs its forward on CPU and stores, per case: the small outputs in full
(layer 0 / layer 6 weights, all biases, freq, phase) and, for the five 256x256 weight matrices, a seeded 4096-entry sample plus
sum and Frobenius norm (float64) -> tests/golden/hyper_s{seed}.npz.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402
from oracle import hyper_oracle as ho         # noqa: E402
from arah_release_b200 import synthetic as syn  # noqa: E402


def main():
    rh.install()
    import im2mesh.metaavatar_render  # noqa: F401
    from im2mesh.metaavatar.models import siren_modules
    for seed, rel in ((0, False), (1, True)):
        sd = syn.make_hypernet_state_dict(seed)
        net = siren_modules.HyperBVPNet(in_features=3, num_hidden_layers=5, hierarchical_pose=True, hyper_in_ch=144, use_FiLM=True,
                                        rel_joints=rel)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        net.eval()
        rots, Jtrs, latent = syn.make_hypernet_inputs(seed)
        torch.set_num_threads(8)
        with torch.no_grad():
            out = net({'coords': torch.zeros(1, 1, 3), 'rots': torch.from_numpy(rots), 'Jtrs': torch.from_numpy(Jtrs),
                       'latent': torch.from_numpy(latent)})
        dec = out['decoder']
        g = {'rel_joints': np.array(int(rel))}
        for l in range(7):
            lay = dec[l][0] if l < 6 else dec[l]
            W = lay.weights[0].numpy()
            g[f'b{l}'] = lay.biases.reshape(-1).numpy()
            if l in (0, 6):
                g[f'W{l}'] = W
            else:
                idx = ho.sample_index(100 * seed + l, W.size)
                g[f'W{l}_sample'] = W.reshape(-1)[idx]
                g[f'W{l}_sum'] = np.array(W.astype(np.float64).sum())
                g[f'W{l}_norm'] = np.array(np.sqrt((W.astype(np.float64) ** 2).sum()))
        g['freq'] = torch.stack([dec[l][0].freq[0] for l in range(6)]).numpy()
        g['phase'] = torch.stack([dec[l][0].phase_shift[0] for l in range(6)]).numpy()
        g['params_numel'] = np.array([p.numel() for p in out['params']])
        path = os.path.join(ROOT, 'tests', 'golden', f'hyper_s{seed}.npz')
        np.savez_compressed(path, **g)
        print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
