"""TEST INFRASTRUCTURE — numpy fp32 restatement of the MetaAvatar hypernetwork forward.

Follows /root/reference/im2mesh:
  metaavatar/models/siren_modules.py:196-244   HierarchicalPoseEncoder.forward
  hyperlayers.py:107-139                       CustomMappingNetwork (Linear / LeakyReLU(0.2) x3, Linear; split in halves)
  hyperlayers.py:270-285                       HyperFCFiLM.forward (freq / phase slices per layer)
  hyperlayers.py:497-510, 453-466              HyperLinearFiLM / HyperLinear.forward (hypo_params + hypo_params_init, split)
  /root/reference/pytorch_prototyping/pytorch_prototyping.py:12-81   FCBlock = FCLayer(Linear, LayerNorm, ReLU) x2 + Linear
Parity pinned: tests/golden/hyper_s*.npz are outputs of the UNMODIFIED reference module (oracle/gen_golden_hyper.py) for
the seeded parameters of arah_release_b200.synthetic.make_hypernet_state_dict; tests/test_hyper_oracle.py checks this file
against them.  Only tests/ and bench.py's CPU leg may import this module.
"""
import numpy as np

KTREE_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21], dtype=np.int32)
IN_CH = [3, 256, 256, 256, 256, 256, 256]
OUT_CH = [256, 256, 256, 256, 256, 256, 1]
F32 = np.float32


def _linear(sd, key, x):
    return (sd[key + '.weight'] @ x + sd[key + '.bias']).astype(F32)


def _layernorm(x, g, b, eps=1e-5):
    mean = x.mean(dtype=F32)
    var = ((x - mean) ** 2).mean(dtype=F32)
    return ((x - mean) / np.sqrt(var + F32(eps)) * g + b).astype(F32)


def pose_encoder(sd, rots, Jtrs, rel_joints=False):
    rots = np.asarray(rots, F32).reshape(24, 9)
    Jtrs = np.asarray(Jtrs, F32).reshape(24, 3).copy()
    if rel_joints:                                                     # siren_modules.py:220-224
        rel = Jtrs.copy()
        rel[1:] = Jtrs[1:] - Jtrs[KTREE_PARENTS[1:]]
        Jtrs = rel
    gfeat = _linear(sd, 'pose_encoder.layer_0', np.concatenate([rots.reshape(-1), Jtrs.reshape(-1)]))
    out = [None] * 24
    for j in range(24):
        p = KTREE_PARENTS[j]
        if p == -1:
            bone = np.linalg.norm(Jtrs[j])
            feat = gfeat
        else:
            bone = np.linalg.norm(Jtrs[j] if rel_joints else Jtrs[j] - Jtrs[p])
            feat = out[p]
        x = np.concatenate([rots[j], Jtrs[j], np.array([bone], F32), feat]).astype(F32)
        h = np.maximum(_linear(sd, f'pose_encoder.layers.{j}.0', x), 0)
        out[j] = _linear(sd, f'pose_encoder.layers.{j}.2', h)
    return np.concatenate(out).astype(F32)


def mapping_network(sd, latent):
    h = np.asarray(latent, F32).reshape(-1)
    for n in (0, 2, 4):
        h = _linear(sd, f'net.mapping_network.network.{n}', h)
        h = np.where(h > 0, h, F32(0.2) * h).astype(F32)
    fo = _linear(sd, 'net.mapping_network.network.6', h)
    return fo[:fo.shape[0] // 2], fo[fo.shape[0] // 2:]


def forward(sd, rots, Jtrs, latent=None, rel_joints=False):
    """-> dict(W=[7 arrays [out,in]], b=[7 arrays [out]], freq [6,256], phase [6,256]) — what HyperFCFiLM.forward assembles."""
    cond = pose_encoder(sd, rots, Jtrs, rel_joints)
    freq, phase = mapping_network(sd, np.zeros(128, F32) if latent is None else latent)
    W, b = [], []
    for l in range(7):
        pre = f'net.layers.{l}.hyper_linear.' if l < 6 else f'net.layers.{l}.'
        fc = pre + 'hypo_params.net.'
        h = np.maximum(_layernorm(_linear(sd, fc + '0.net.0', cond), sd[fc + '0.net.1.weight'], sd[fc + '0.net.1.bias']), 0)
        h = np.maximum(_layernorm(_linear(sd, fc + '1.net.0', h), sd[fc + '1.net.1.weight'], sd[fc + '1.net.1.bias']), 0)
        hp = _linear(sd, fc + '2', h) + sd[pre + 'hypo_params_init'].reshape(-1)
        n_w = IN_CH[l] * OUT_CH[l]
        W.append(hp[:n_w].reshape(OUT_CH[l], IN_CH[l]).astype(F32))
        b.append(hp[n_w:n_w + OUT_CH[l]].astype(F32))
    return {'W': W, 'b': b, 'freq': freq.reshape(6, 256).astype(F32), 'phase': phase.reshape(6, 256).astype(F32)}


def sample_index(seed, n, k=4096):
    return np.random.default_rng(seed).choice(n, size=min(k, n), replace=False)
