"""CPU: the torch restatement of the training branch (oracle/train_oracle.py) and the C oracle's training-mode tracer
against one training step of the UNMODIFIED reference (tests/golden/train_*.npz, oracle/gen_golden_train.py)."""
import numpy as np
import pytest

from helpers_train import TRAIN_CASES, compare_grad, load_train_golden, ref_to_oracle_name


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_c_oracle_train_trace_matches_reference(name):
    """BodyRayTracing.forward(eval_mode=False): all rays enter the joint search, z samples are jittered (ray_tracing.py:249,298-311)."""
    from oracle import oracle as orc
    fr, aux, ref, grads, meta = load_train_golden(name)
    c = orc.render(fr, train_noise=orc.train_noise(fr, meta['seed']))
    assert (ref['trace.network_body_mask'] != c['trace.network_body_mask']).mean() <= 0.01
    cm = ref['trace.sampler_converge_mask'].astype(bool)
    assert (cm != c['trace.sampler_converge_mask']).mean() <= 1e-3
    both = cm & c['trace.sampler_converge_mask']
    assert np.abs(ref['trace.sampled_dists'] - c['trace.sampled_dists']).max() <= 1e-5          # incl. the jitter
    assert np.abs(ref['trace.sampled_pts'] - c['trace.sampled_pts'])[both].max() <= 3e-4
    # forward values of the training render = eval shading of the jittered samples
    assert np.abs(ref['out.rgb_values'][0] - c['rgb_values']).max() <= 2e-4
    assert np.abs(ref['out.sdf_output'][0] - c['weights_sum']).max() <= 2e-4


@pytest.mark.parametrize('name', TRAIN_CASES)
def test_torch_oracle_matches_reference_step(name):
    from oracle import train_oracle as to
    fr, aux, ref, grads, meta = load_train_golden(name)
    trace = {k[len('trace.'):]: v for k, v in ref.items() if k.startswith('trace.')}
    out, g = to.train_step(fr, aux, trace, meta['seed'], train_skinning_net=meta['train_skinning_net'], loss_weights=meta['loss_weights'])
    assert np.abs(out['rgb_values'] - ref['out.rgb_values'][0]).max() <= 1e-5
    assert np.abs(out['sdf_output'] - ref['out.sdf_output'][0]).max() <= 1e-5
    assert np.abs(out['grad_theta'] - ref['out.grad_theta']).max() <= 1e-4 * max(1.0, np.abs(ref['out.grad_theta']).max())
    assert np.abs(out['pred_weights'] - ref['out.pred_weights'][0]).max() <= 1e-5
    lw = dict(to.LOSS_WEIGHTS); lw.update(meta['loss_weights'])
    for k in ('loss', 'rgb_loss', 'eikonal_loss', 'mask_loss', 'off_surface_loss', 'inside_loss', 'skinning_loss'):
        if k != 'loss' and lw[k.replace('_loss', '_weight')] == 0:
            continue                      # IDHRLoss reports zeros for disabled terms (loss.py:137-175)
        assert abs(out['loss.' + k] - float(ref['loss.' + k])) <= 1e-5 * max(1.0, abs(float(ref['loss.' + k]))), k
    worst = 1.0
    for k, dig in grads.items():
        st = compare_grad(k, g[ref_to_oracle_name(k)], dig)
        worst = min(worst, st['cos'])
    print(name, 'min gradient cosine vs reference', worst)
