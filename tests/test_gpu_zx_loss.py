"""GPU parity for the fused training loss ("fused loss reductions"): `arah_release_b200.loss.IDHRLoss` ->
`arah_idhr_loss` (csrc/arah_loss.cu) against

* the nine terms and the autograd gradients of the UNMODIFIED reference `IDHRLoss` (tests/golden/loss_s*.npz,
  oracle/gen_golden_loss.py): terms 2e-6 relative, gradients 2e-6 of each tensor's largest entry, identical zero patterns
  (masks, sign(0), norm at the origin) — tolerances in tests/helpers_loss.py;
* itself: bit-identical re-runs (integer atomic + fixed reduction trees);
* torch on the GPU inside a real training step: the same `IDHRNetwork` training forward, once with the torch restatement of the
  loss (oracle/train_oracle.py::loss_terms) and once with the fused criterion — parameter gradients must agree to 5e-5 relative
  Frobenius error.
The same assertions already hold on the CPU for the host build of the kernels' arithmetic (tests/test_loss_host.py)."""
import numpy as np
import pytest
import torch

from helpers_loss import LOSS_SEEDS, TERMS, check_loss, load_loss_golden

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _run(cfg, full):
    from arah_release_b200.loss import IDHRLoss
    crit = IDHRLoss(rgb_loss_type=cfg['rgb_loss_type'], **{k: v for k, v in cfg.items() if k.endswith('_weight')})
    t = lambda a, rg=False: torch.from_numpy(np.ascontiguousarray(a)).to(DEV).unsqueeze(0).requires_grad_(rg)
    leaves = {k: t(full[k], True) for k in ('rgb_values', 'sdf_output', 'pred_weights')}
    leaves.update({k: torch.from_numpy(full[k]).to(DEV).requires_grad_(True) for k in ('grad_theta', 'off_surface_sdf', 'inside_sdf')})
    params = [t(p, True) for p in full['sdf_params']]
    mo = {'rgb_values': leaves['rgb_values'], 'sdf_output': leaves['sdf_output'], 'network_body_mask': t(full['network_body_mask']),
          'body_mask': t(full['body_mask']), 'off_surface_mask': t(full['off_surface_mask']), 'surface_normals': None, 'grad_theta': leaves['grad_theta'],
          'off_surface_sdf': leaves['off_surface_sdf'], 'inside_sdf': leaves['inside_sdf'], 'pred_weights': leaves['pred_weights'], 'sdf_params': params}
    out = crit(mo, {'rgb': t(full['rgb_gt']), 'sampled_weights': t(full['sampled_weights'])})
    out['loss'].sum().backward()
    torch.cuda.synchronize()
    terms = {k: float(out[k].detach().reshape(-1)[0]) for k in TERMS}
    grads = {k: (v.grad.cpu().numpy().reshape(np.asarray(full[k]).shape) if v.grad is not None else None) for k, v in leaves.items()}
    grads['sdf_params'] = [p.grad.cpu().numpy().reshape(-1) if p.grad is not None else np.zeros(p.numel(), np.float32) for p in params]
    return out, terms, grads


@pytest.mark.parametrize('seed', LOSS_SEEDS)
def test_fused_loss_matches_reference(seed):
    cfg, cut, full, ref = load_loss_golden(seed)
    out, terms, grads = _run(cfg, full)
    assert tuple(out) == TERMS and tuple(out['loss'].shape) == tuple(ref['loss_shape'])
    check_loss(terms, grads, ref, full)
    if full['rgb_values'].shape[0] > 2048 and grads['rgb_values'] is not None:
        assert not grads['rgb_values'][2048:].any()                   # rows beyond the reference's 2048-ray cut
    _, terms2, grads2 = _run(cfg, full)
    assert terms2 == terms
    for k in ('rgb_values', 'sdf_output', 'grad_theta', 'off_surface_sdf', 'inside_sdf', 'pred_weights'):
        assert (grads[k] is None and grads2[k] is None) or np.array_equal(grads[k], grads2[k])


def test_fused_loss_rejects_what_it_does_not_implement():
    from arah_release_b200 import _lib
    from arah_release_b200.loss import IDHRLoss
    cfg, cut, full, ref = load_loss_golden(3)
    w = {k: v for k, v in cfg.items() if k.endswith('_weight')}
    with pytest.raises(ValueError):
        IDHRLoss(rgb_loss_type='huber', **w)
    crit = IDHRLoss(**dict(w, perceptual_weight=0.1))
    rgb = torch.zeros(1, 8, 3, device=DEV)
    with pytest.raises(_lib.ArahError):
        crit({'rgb_values': rgb}, {})
    with pytest.raises(_lib.ArahError):
        IDHRLoss(**w)({'rgb_values': rgb.cpu()}, {})                   # no CPU path


def test_fused_loss_inside_a_training_step():
    """One training forward of the drop-in renderer, then backward through (a) torch's restatement of the loss and (b) the fused
    criterion: same loss, same parameter gradients."""
    from helpers_train import TRAIN_CASES, load_train_golden
    from test_gpu_train import build_net, loss_of, train_inputs
    from arah_release_b200.loss import IDHRLoss
    from oracle import train_oracle as to
    name = TRAIN_CASES[0]
    fr, aux, ref, _, meta = load_train_golden(name)
    lw = dict(to.LOSS_WEIGHTS); lw.update(meta['loss_weights'])
    results = []
    for fused in (False, True):
        net, sdf, sdf_leaves = build_net(fr, meta['train_skinning_net'], train_mode='fp32')
        inp = train_inputs(fr, sdf, aux)
        torch.manual_seed(meta['seed'])
        out = net(inp)
        if fused:
            crit = IDHRLoss(rgb_loss_type='l1', **lw)
            mo = dict(out)
            mo['sdf_params'] = []
            gt = {'rgb': torch.from_numpy(np.ascontiguousarray(aux['rgb_gt'])).to(DEV).view(1, -1, 3),
                  'sampled_weights': torch.from_numpy(np.ascontiguousarray(aux['sampled_weights'])).to(DEV).view(1, -1, 24)}
            loss = crit(mo, gt)['loss'].sum()
        else:
            loss = loss_of(out, aux, lw)['loss']
        loss.backward()
        torch.cuda.synchronize()
        g = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
        g.update({k: v.grad.detach().clone() for k, v in sdf_leaves.items() if v.grad is not None})
        results.append((float(loss.detach()), g))
    (l0, g0), (l1, g1) = results
    assert abs(l0 - l1) <= 5e-6 * max(1.0, abs(l0)), (l0, l1)
    assert set(g0) == set(g1) and len(g0) > 20
    for k in g0:
        num, den = float((g0[k] - g1[k]).norm()), float(g0[k].norm())
        assert num <= 5e-5 * den + 1e-9, (k, num, den)          # split-K weight gradients accumulate with atomics: not bit-reproducible
