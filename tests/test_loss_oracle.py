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
