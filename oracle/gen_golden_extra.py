#!/usr/bin/env python
"""TEST INFRASTRUCTURE — extra fixtures from the UNMODIFIED reference (build container only, needs /root/reference):

  * eval `weights_sum` (the second return of IDHRNetwork.get_rbg_value_vol_sdf, implicit_differentiable_renderer.py:392,
    scattered over the rays as :224-226 does) for two of the existing cases -> tests/golden/wsum_*.npz;
  * one frame rendered with IDHRNetwork(render_last_pt=True) (:380-381) -> tests/golden/lastpt_*.npz.

Usage: python oracle/gen_golden_extra.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from arah_release_b200 import synthetic as syn  # noqa: E402
from oracle import ref_harness as rh            # noqa: E402

CASES = {
    'wsum_zju377_24x24_s0': (dict(H=24, W=24, seed=0), False),
    'wsum_cano_20x20_s1': (dict(H=20, W=20, seed=1, cano_view_dirs=True, max_angle=0.8, beta=2e-3), False),
    'lastpt_16x16_s3': (dict(H=16, W=16, seed=3, beta=4e-3), True),
}


def run(kw, last_pt):
    fr = syn.make_frame(**kw)
    idhr, sdf_network = rh.build_reference_modules(fr)
    idhr.render_last_pt = bool(last_pt)
    inputs = rh.reference_inputs(fr, sdf_network)
    rec = []
    orig = idhr.get_rbg_value_vol_sdf

    def wrapped(*a, **k):
        rgb, ws = orig(*a, **k)
        rec.append(ws.detach().clone())
        return rgb, ws
    idhr.get_rbg_value_vol_sdf = wrapped
    torch.set_num_threads(os.cpu_count())
    res = idhr(inputs)
    mask = res['network_body_mask'][0].numpy().copy()
    ws = np.zeros(fr.P, np.float32)
    ws[mask.astype(bool)] = torch.cat(rec, 0).reshape(-1).numpy()
    return fr, {'rgb_values': res['rgb_values'][0].detach().numpy().copy(), 'network_body_mask': mask, 'weights_sum': ws}


def main():
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for name, (kw, last_pt) in CASES.items():
        t = time.time()
        fr, arrays = run(kw, last_pt)
        meta = {'make_frame': kw, 'render_last_pt': bool(last_pt), 'P': fr.P, 'reference_seconds': time.time() - t,
                'generator': 'oracle/gen_golden_extra.py', 'reference_commit': '1040cf7'}
        path = os.path.join(out_dir, name + '.npz')
        np.savez_compressed(path, meta=json.dumps(meta), **arrays)
        print(name, fr.P, f'{meta["reference_seconds"]:.1f}s', os.path.getsize(path), 'bytes',
              'wsum range', float(arrays['weights_sum'].min()), float(arrays['weights_sum'].max()))


if __name__ == '__main__':
    main()
