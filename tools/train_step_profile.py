#!/usr/bin/env python
"""Run a few training steps (BASELINE configs[2] shape) for profiling: `ncu --metrics gpu__time_duration.sum ... python tools/train_step_profile.py`."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import bench  # noqa: E402

if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=1)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--train-rays', type=int, default=2048)
    ap.add_argument('--size', type=int, default=512)
    a = ap.parse_args()
    import torch
    from arah_release_b200 import synthetic as syn
    fr = syn.make_frame(a.size, a.size, seed=0)
    print(bench.train_step_bench(a, torch.device('cuda:0'), fr, steps=a.steps, warmup=a.warmup))
