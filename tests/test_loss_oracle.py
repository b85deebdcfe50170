"""CPU: the numpy oracle of the training loss (oracle/loss_oracle.py) against the terms and autograd gradients of the unmodified
reference `IDHRLoss` (tests/golden/loss_s*.npz, oracle/gen_golden_loss.py).  Tolerances: tests/helpers_loss.py."""
import numpy as np
import pytest

from helpers_loss import LOSS_SEEDS, check_loss, load_loss_golden
from oracle import loss_oracle as lo


@pytest.mark.parametrize('seed', LOSS_SEEDS)
def test_loss_oracle_matches_reference(seed):
    cfg, cut, full, ref = load_loss_golden(seed)
    terms, grads = lo.idhr_loss(cfg, cut)
    check_loss(terms, grads, ref, full)
    # the reference's loss is a [1]-shaped tensor whenever a disabled term contributes its torch.zeros(1) (loss.py:135-170)
    assert tuple(ref['loss_shape']) == ((1,) if min(cfg[k] for k in cfg if k.endswith('_weight')) <= 0 else ())


def test_mask_term_histogram_form_equals_the_broadcast_form():
    """loss.py:100-101 broadcasts [n,1] - [n] to [n,n]; the O(n) form the kernels use — sqrt(sum_v count_v (w_i - v)^2) over the
    histogram of ground-truth values — must give the same numbers."""
    rng = np.random.default_rng(0)
    w = rng.random(500)
    gt = rng.choice([0, 1, 100], size=500, p=[0.5, 0.4, 0.1]).astype(np.float64)
    D = w[:, None] - gt[None, :]
    literal = np.sqrt((D * D).sum(-1))
    vals, cnt = np.unique(gt, return_counts=True)
    hist = np.sqrt(((w[:, None] - vals[None, :]) ** 2 * cnt[None, :]).sum(-1))
    np.testing.assert_allclose(hist, literal, rtol=1e-12)
    np.testing.assert_allclose(((w[:, None] - vals[None, :]) * cnt[None, :]).sum(-1) / hist, D.sum(-1) / literal, rtol=1e-10)
