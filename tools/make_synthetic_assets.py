#!/usr/bin/env python
"""Fit the synthetic SDF (FiLM-SIREN) and skinning nets to the analytic capsule body and write
arah_release_b200/data/synthetic_nets_v1.npz.

Run once on CPU (a few minutes): ``python tools/make_synthetic_assets.py``.  The output is committed; parity
fixtures under tests/golden/ are generated from it, so re-running this script invalidates them.

Why: a random-init hypernetwork yields SDF == 0 (/root/reference/im2mesh/hyperlayers.py:440-441 zero-init),
and no pretrained checkpoint is reachable, so a synthetic network with a body-like zero level
set is the only way to exercise sphere tracing / root finding meaningfully.  Shapes follow
configs/arah-zju/ZJUMOCAP-377_4gpus.yaml:34-43.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as Fnn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from arah_release_b200 import synthetic as syn  # noqa: E402


def tree_softmax(x):
    """Hierarchical softmax over the SMPL kinematic tree; same function as
    /root/reference/im2mesh/utils/utils.py:138-181 written as a table of (child, parent, gate) edges."""
    sig = torch.sigmoid
    p = [None] * 24
    root = torch.ones_like(x[:, 0])
    sm = torch.softmax(x[:, 1:4], -1)
    for i in range(3):
        p[1 + i] = root * sig(x[:, 0]) * sm[:, i]
    p[0] = root * (1 - sig(x[:, 0]))
    def split(children, parents, gates):
        for c, q, g in zip(children, parents, gates):
            p[c] = p[q] * sig(x[:, g])
        for q, g in zip(parents, gates):
            p[q] = p[q] * (1 - sig(x[:, g]))
    split([4, 5, 6], [1, 2, 3], [4, 5, 6])
    split([7, 8, 9], [4, 5, 6], [7, 8, 9])
    split([10, 11], [7, 8], [10, 11])
    sm2 = torch.softmax(x[:, 12:15], -1)
    p9 = p[9]
    for i in range(3):
        p[12 + i] = p9 * sig(x[:, 24]) * sm2[:, i]
    p[9] = p9 * (1 - sig(x[:, 24]))
    split([15], [12], [15])
    split([16, 17], [13, 14], [16, 17])
    split([18, 19], [16, 17], [18, 19])
    split([20, 21], [18, 19], [20, 21])
    split([22, 23], [20, 21], [22, 23])
    return torch.stack(p, -1)


class FilmSiren(nn.Module):
    def __init__(self, freq, phase):
        super().__init__()
        dims = [3] + [256] * 6 + [1]
        self.lins = nn.ModuleList()
        for i in range(7):
            lin = nn.Linear(dims[i], dims[i + 1])
            with torch.no_grad():
                if i == 0:
                    lin.weight.uniform_(-1 / dims[i], 1 / dims[i])
                else:
                    lin.weight.uniform_(-np.sqrt(6 / dims[i]) / 30, np.sqrt(6 / dims[i]) / 30)
            self.lins.append(lin)
        self.register_buffer('freq', freq)
        self.register_buffer('phase', phase)

    def forward(self, x):
        h = x
        for i in range(6):
            h = torch.sin(30 * (self.freq[i] * self.lins[i](h) + self.phase[i]))
        return self.lins[6](h)


class SkinMLP(nn.Module):
    def __init__(self):
        super().__init__()
        dims = [3, 128, 128, 128, 128, 25]
        self.lins = nn.ModuleList([nn.utils.weight_norm(nn.Linear(dims[i], dims[i + 1])) for i in range(5)])

    def forward(self, x):
        h = x
        for i in range(5):
            h = self.lins[i](h)
            if i < 4:
                h = Fnn.softplus(h, beta=100)
        return h


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    cmin, cmax, center = syn.canonical_normalisation()
    scale = 1.1 * (float(cmax) - float(cmin)) / 2.0   # normalised unit -> metres (root_finding_utils.py:37-51)
    print('coord_min', cmin, 'coord_max', cmax, 'center', center, 'scale', scale)

    def unnorm(pn):
        pad = (cmax - cmin) * 0.05
        return (pn / 2.0 + 0.5) * 1.1 * (cmax - cmin) + cmin - pad + center

    verts = syn.sample_body_vertices(np.random.default_rng(1234))

    def sample_points(n):
        n1 = n // 2
        a = rng.uniform(-1, 1, size=(n - n1, 3))
        v = verts[rng.integers(len(verts), size=n1)] + rng.normal(scale=0.04, size=(n1, 3))
        pad = (cmax - cmin) * 0.05
        vn = ((v - center - cmin + pad) / (cmax - cmin) / 1.1 - 0.5) * 2
        return np.concatenate([a, vn], 0)

    # ---------------- SDF ----------------
    frng = np.random.default_rng(77)
    freq = torch.tensor(1.0 + 0.1 * frng.normal(size=(6, 256)), dtype=torch.float32)
    phase = torch.tensor(0.02 * frng.normal(size=(6, 256)), dtype=torch.float32)
    net = FilmSiren(freq, phase)
    steps = int(os.environ.get('SDF_STEPS', 2500))
    opt = torch.optim.Adam(net.parameters(), lr=2e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, steps, eta_min=1e-6)
    t0 = time.time()
    for it in range(steps):
        pn = sample_points(8192)
        tgt = syn.body_sdf(unnorm(pn)) / scale
        x = torch.tensor(pn, dtype=torch.float32)
        y = torch.tensor(tgt, dtype=torch.float32).unsqueeze(-1)
        pred = net(x)
        w = 1.0 + 4.0 * (y.abs() < 0.1).float()
        loss = (w * (pred - y).abs()).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        if it % 100 == 0 or it == steps - 1:
            print(f'sdf {it} loss {loss.item():.5f} ({time.time() - t0:.0f}s)', flush=True)

    # ---------------- skinning ----------------
    skin = SkinMLP()
    steps = int(os.environ.get('SKIN_STEPS', 3000))
    opt = torch.optim.Adam(skin.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, steps, eta_min=1e-5)
    for it in range(steps):
        pn = sample_points(4096)
        tgt = syn.body_weights(unnorm(pn), sharp=0.05)
        x = torch.tensor(pn, dtype=torch.float32)
        y = torch.tensor(tgt, dtype=torch.float32)
        w = tree_softmax(skin(x) * 20)
        loss = (w - y).abs().sum(-1).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        if it % 200 == 0 or it == steps - 1:
            print(f'skin {it} loss {loss.item():.5f}', flush=True)

    out = {'coord_min': np.float32(cmin), 'coord_max': np.float32(cmax), 'center': center.astype(np.float32),
           'sdf_freq': freq.numpy(), 'sdf_phase': phase.numpy()}
    for i in range(7):
        out[f'sdf_W{i}'] = net.lins[i].weight.detach().numpy().astype(np.float32)
        out[f'sdf_b{i}'] = net.lins[i].bias.detach().numpy().astype(np.float32)
    for i in range(5):
        out[f'skin_v{i}'] = skin.lins[i].weight_v.detach().numpy().astype(np.float32)
        out[f'skin_g{i}'] = skin.lins[i].weight_g.detach().numpy().astype(np.float32)
        out[f'skin_b{i}'] = skin.lins[i].bias.detach().numpy().astype(np.float32)
    path = os.path.join(os.path.dirname(os.path.abspath(syn.__file__)), 'data', 'synthetic_nets_v1.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path) / 1e6, 'MB')


if __name__ == '__main__':
    main()
