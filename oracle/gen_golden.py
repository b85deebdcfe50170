#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Build-container only (needs /root/reference).  Usage: ``python oracle/gen_golden.py``.
Each fixture stores the frame recipe (arguments of arah_release_b200.synthetic.make_frame, plus optional
degenerate rays) and the reference's outputs; tests rebuild the inputs from the recipe, so only outputs ship.
``sampled_transforms`` (4 KB/ray) is kept for the first 48 rays only to bound fixture size.
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
    # name: (make_frame kwargs, n degenerate rays appended with near == far)
    'zju377_24x24_s0': (dict(H=24, W=24, seed=0), 0),
    'cano_20x20_s1': (dict(H=20, W=20, seed=1, cano_view_dirs=True, max_angle=0.8, beta=2e-3), 0),
    'n32_16x16_s2': (dict(H=16, W=16, seed=2, n_steps=32, near_samples=8, far_samples=4, beta=1e-2), 3),
    # BASELINE configs[4] shape (H36M-style): 128 near-surface samples need n_steps >= 145 (ray_tracing.py:336,346), canonical view dirs
    'h36m_n160_12x12_s4': (dict(H=12, W=12, seed=4, n_steps=160, near_samples=128, far_samples=16, cano_view_dirs=True, beta=3e-3), 0),
    # monocular configs (configs/arah-zju/ZJUMOCAP-39x-mono_4gpus.yaml:36): colour net without view directions (390 inputs)
    'mono_noview_16x16_s8': (dict(H=16, W=16, seed=8, color_mode='no_view_dir'), 0),
    # third RenderingNetwork mode (decoder.py:104-106; no shipped config uses it): colour net without the normal (414 inputs)
    'nonormal_16x16_s9': (dict(H=16, W=16, seed=9, color_mode='no_normal', cano_view_dirs=True), 0),
    # BASELINE configs[0]: the reference's own CPU-runnable plumbing case, a 64x64 frame.  Final outputs + per-ray tracer outputs only
    # (the per-sample stage tensors would be 4.5 MB): see FINAL_ONLY
    'zju377_64x64_s0': (dict(H=64, W=64, seed=0), 0),
}
T_RAYS = 48
FINAL_ONLY = {'zju377_64x64_s0'}
FINAL_KEYS = ('rgb_values', 'network_body_mask', 'points_cam', 'trace.network_body_mask', 'trace.dists', 'trace.points_hat_norm')


def build_case(kw, n_degenerate):
    fr = syn.make_frame(**kw)
    if n_degenerate:
        # empty intervals: near == far (the reference asserts near <= far, ray_tracing.py:182)
        rd = np.concatenate([fr.ray_dirs, fr.ray_dirs[:n_degenerate]], 0)
        nf = np.concatenate([fr.near_far, np.repeat(fr.near_far[:n_degenerate, 1:2], 2, axis=1)], 0)
        fr.ray_dirs, fr.near_far = rd, nf
        fr.pix = np.concatenate([fr.pix, fr.pix[:n_degenerate]])
    return fr


def main(only=None):
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    for name, (kw, ndeg) in CASES.items():
        if only and name not in only:
            continue
        fr = build_case(kw, ndeg)
        c = rh.Counters()
        t = time.time()
        ref = rh.run_reference(fr, counters=c, threads=os.cpu_count())
        dt = time.time() - t
        iso = [x for x in c.calls if x['dim'] == 4][:1]
        corr = [x for x in c.calls if x['dim'] == 3][:1]
        meta = {'make_frame': kw, 'n_degenerate': ndeg, 'P': fr.P, 'reference_seconds': dt,
                'reference_threads': os.cpu_count(), 'iso_calls': iso, 'corr_calls': corr,
                'generator': 'oracle/gen_golden.py', 'reference_commit': '1040cf7'}
        if name in FINAL_ONLY:
            arrays = {k.replace('.', '__'): ref[k] for k in FINAL_KEYS}
        else:
            arrays = {k.replace('.', '__'): v for k, v in ref.items() if k != 'trace.sampled_transforms'}
            arrays['trace__sampled_transforms_head'] = ref['trace.sampled_transforms'][:T_RAYS]
        path = os.path.join(out_dir, name + '.npz')
        np.savez_compressed(path, meta=json.dumps(meta), **arrays)
        print(name, 'P', fr.P, f'{dt:.1f}s', os.path.getsize(path) / 1e6, 'MB', meta['iso_calls'], meta['corr_calls'])


if __name__ == '__main__':
    main(sys.argv[1:] or None)
