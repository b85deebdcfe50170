"""GPU parity at BASELINE configs[0] (the reference's own CPU-runnable case: one 64x64 frame): the CUDA path through the drop-in
modules -> C ABI against the unmodified reference's outputs (tests/golden/zju377_64x64_s0.npz, oracle/gen_golden.py) and against
the CPU oracle.  Tolerances: tests/helpers.py::TOL; the tensor-core mode relaxes only PSNR(ours, reference) to 55 dB (TF32 operand
rounding in the shading MLPs), the 0.05 dB delta-PSNR bar stays."""
import numpy as np
import pytest
import torch

from helpers import TOL, check_render, load_golden
from test_gpu_parity import _build, _render_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('mode', ['fp32', 'tf32'])
def test_config0_64x64_matches_reference(mode):
    fr, ref, meta = load_golden('zju377_64x64_s0')
    net, inputs = _build(fr, mode)
    out = _render_dict(net, inputs)
    st = check_render(out, ref, label='config0:' + mode, tol=dict(TOL, rgb_psnr_min=55.0) if mode == 'tf32' else TOL, stages=False)
    stats = net.stats()
    assert stats['rays'] == fr.P and stats['vol_rays'] == int(out['network_body_mask'].sum())
    print('config0', mode, st)


def test_config0_64x64_matches_oracle():
    from oracle import oracle as orc
    fr, _, _ = load_golden('zju377_64x64_s0')
    net, inputs = _build(fr)
    out = _render_dict(net, inputs)
    o = orc.render(fr)
    st = check_render(out, o, label='config0:oracle')
    print('config0 oracle', st)


@pytest.mark.parametrize('mode', ['fp32', 'tf32'])
def test_no_normal_colour_mode_matches_reference(mode):
    """RenderingNetwork mode 'no_normal' (decoder.py:104-106) against the unmodified reference (tests/golden/nonormal_16x16_s9.npz)."""
    fr, ref, meta = load_golden('nonormal_16x16_s9')
    net, inputs = _build(fr, mode)
    out = _render_dict(net, inputs)
    st = check_render(out, ref, label='no_normal:' + mode, tol=dict(TOL, rgb_psnr_min=55.0) if mode == 'tf32' else TOL)
    print('no_normal', mode, st)
