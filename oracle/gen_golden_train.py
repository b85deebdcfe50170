#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generate tests/golden/train_*.npz: one TRAINING step (forward + backward) of the UNMODIFIED
reference IDHRNetwork on CPU (oracle/ref_harness.run_reference_train).  Build-container only (needs /root/reference).

Stored per case: the recipe (make_frame kwargs, seeds, flags, loss weights), the training-mode tracer outputs, the model
outputs, the loss terms and the gradient of every parameter tensor.  Gradients of large matrices are stored as a seeded
sample of entries plus their sum and Frobenius norm (grad_digest below) to keep fixtures small.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from arah_release_b200 import synthetic as syn  # noqa: E402
from oracle import ref_harness as rh            # noqa: E402

CASES = {
    # BASELINE configs[2] shape: ZJU-313 (cano_view_dirs false, train_skinning_net true)
    'train_zju313_16x16_s5': dict(frame=dict(H=16, W=16, seed=5), seed=11, train_skinning_net=True, loss_weights={}),
    # implicit-gradient path in isolation: the skinning net only gets gradient through pi' = pi - J^-1 (lbs - lbs.detach())
    'train_implicit_12x12_s7': dict(frame=dict(H=12, W=12, seed=7, max_angle=0.8), seed=3, train_skinning_net=True,
                                    loss_weights=dict(skinning_weight=0.0, eikonal_weight=0.0, off_surface_weight=0.0, inside_weight=0.0)),
    # canonical view directions (H36M-style), fewer samples, skinning net frozen
    'train_cano_12x12_s6': dict(frame=dict(H=12, W=12, seed=6, cano_view_dirs=True, n_steps=32, near_samples=8, far_samples=4, beta=2e-3),
                                seed=5, train_skinning_net=False, loss_weights={}),
}
SAMPLE = 2048


def grad_digest(name, g):
    g = np.asarray(g, np.float32)
    if g.size <= 4096:
        return {'full': g}
    rng = np.random.default_rng(sum(map(ord, name)))
    idx = rng.choice(g.size, size=SAMPLE, replace=False).astype(np.int64)
    return {'idx': idx, 'val': g.reshape(-1)[idx], 'sum': np.float64(g.astype(np.float64).sum()),
            'fro': np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))}


def main(only=None):
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for name, c in CASES.items():
        if only and name not in only:
            continue
        fr = syn.make_frame(**c['frame'])
        aux = syn.train_aux_points(fr, seed=c['frame']['seed'])
        t = time.time()
        ref = rh.run_reference_train(fr, aux, seed=c['seed'], train_skinning_net=c['train_skinning_net'], threads=os.cpu_count(),
                                     loss_weights=c['loss_weights'])
        dt = time.time() - t
        meta = dict(make_frame=c['frame'], seed=c['seed'], train_skinning_net=c['train_skinning_net'], loss_weights=c['loss_weights'],
                    P=fr.P, reference_seconds=dt, generator='oracle/gen_golden_train.py')
        arrays = {}
        for k, v in ref.items():
            if k.startswith('grad.'):
                for kk, vv in grad_digest(k, v).items():
                    arrays[k.replace('.', '__') + '___' + kk] = vv
            elif k == 'trace.sampled_transforms':
                arrays['trace__sampled_transforms'] = v[..., :3, :].astype(np.float32)      # rows 0..2; row 3 is [0 0 0 1] / 0
            else:
                arrays[k.replace('.', '__')] = v
        path = os.path.join(out_dir, name + '.npz')
        np.savez_compressed(path, meta=json.dumps(meta), **arrays)
        print(name, 'P', fr.P, f'{dt:.1f}s', f'{os.path.getsize(path) / 1e6:.2f} MB', {k: float(v) for k, v in ref.items() if k.startswith('loss.')})


if __name__ == '__main__':
    main(sys.argv[1:] or None)
