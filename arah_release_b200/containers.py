"""Parameter containers of the per-frame SDF network in the reference's layout, WITHOUT any forward computation.

`HyperSDFDecoder` (hypernet.py) returns the frame's SDF as nn.Sequential(Sequential(BatchLinearFiLM, Sine) x 6, BatchLinear), the
module structure `HyperBVPNet` produces (hyperlayers.py:270-285) and `IDHRNetwork` reads its weights from.  The containers
deliberately cannot compute: the only compute path is the CUDA library.
"""
import torch.nn as nn

class _NoForward(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f'{type(self).__name__} is a parameter container; the ARAH hot path runs in libarah_b200.so only')


class BatchLinearFiLM(_NoForward):
    def __init__(self, weights, biases, freq, phase_shift):
        super().__init__()
        self.weights, self.biases, self.freq, self.phase_shift = weights, biases, freq, phase_shift


class BatchLinear(_NoForward):
    def __init__(self, weights, biases):
        super().__init__()
        self.weights, self.biases = weights, biases


class Sine(_NoForward):
    pass
