"""Comparison of a hypernetwork output (dict W/b/freq/phase) against tests/golden/hyper_s*.npz (reference outputs)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
# fp32, outputs are O(1): the reference is MKL sgemv, ours sequential / butterfly sums of 256 products
ATOL = 2e-5


def load(seed):
    return dict(np.load(os.path.join(GOLDEN_DIR, f'hyper_s{seed}.npz')))


def check(out, gold, seed, atol=ATOL, label=''):
    from oracle import hyper_oracle as ho
    worst = 0.0
    for l in range(7):
        W, b = np.asarray(out['W'][l], np.float32), np.asarray(out['b'][l], np.float32).reshape(-1)
        d = np.abs(b - gold[f'b{l}']).max(); worst = max(worst, d)
        assert d <= atol, (label, 'bias', l, d)
        if l in (0, 6):
            d = np.abs(W - gold[f'W{l}']).max(); worst = max(worst, d)
            assert d <= atol, (label, 'W', l, d)
        else:
            idx = ho.sample_index(100 * seed + l, W.size)
            d = np.abs(W.reshape(-1)[idx] - gold[f'W{l}_sample']).max(); worst = max(worst, d)
            assert d <= atol, (label, 'W sample', l, d)
            assert abs(W.astype(np.float64).sum() - float(gold[f'W{l}_sum'])) <= atol * np.sqrt(W.size) * 4, (label, 'W sum', l)
            nrm = np.sqrt((W.astype(np.float64) ** 2).sum())
            assert abs(nrm - float(gold[f'W{l}_norm'])) <= 1e-5 * float(gold[f'W{l}_norm']), (label, 'W norm', l)
    for k in ('freq', 'phase'):
        d = np.abs(np.asarray(out[k], np.float32).reshape(6, 256) - gold[k]).max(); worst = max(worst, d)
        assert d <= atol, (label, k, d)
    return worst
