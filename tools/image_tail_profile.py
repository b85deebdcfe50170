"""Stand-alone driver of the image-space tail for `ncu`: one 512x512 frame's validation images + PSNR and the three
normal maps of a 256^3-lattice iso-surface, no renderer involved.  Usage: python tools/image_tail_profile.py [--iters 3] [--N 256]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--N', type=int, default=256)
    ap.add_argument('--size', type=int, default=512)
    a = ap.parse_args()
    from arah_release_b200.images import FrameImages
    from arah_release_b200.renderer import ArahRenderer
    dev = torch.device('cuda:0')
    H = W = a.size
    # an analytic two-sphere lattice through the product's own marching cubes
    ax = torch.linspace(-1, 1, a.N, device=dev)
    x, y, z = torch.meshgrid(ax, ax, ax, indexing='ij')
    vol = torch.minimum(((x - 0.35) ** 2 + y ** 2 + z ** 2).sqrt() - 0.3, ((x + 0.35) ** 2 + y ** 2 + z ** 2).sqrt() - 0.33).contiguous()
    verts, faces = ArahRenderer(dev, max_rays=1024).marching_cubes(vol)
    R, T = np.eye(3, dtype=np.float32), np.array([0, 0, 2.6], np.float32)
    K = np.array([[0.9 * W, 0, W / 2], [0, 0.9 * H, H / 2], [0, 0, 1]], np.float32)
    g = torch.Generator(device='cpu').manual_seed(0)
    mask = torch.rand(H, W, generator=g) < 0.25
    pix = mask.view(-1).nonzero().squeeze(1).to(torch.int32).to(dev)
    P = pix.numel()
    rgb, gt = torch.rand(P, 3, generator=g).to(dev), torch.rand(P, 3, generator=g).to(dev)
    pts = torch.rand(P, 3, generator=g).to(dev) + 2.0
    fi = FrameImages(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for i in range(a.iters):
        ev[0].record()
        fi.assemble(rgb, pts, pix, H, W)
        res = fi.psnr_device(rgb, gt)
        maps = fi.normal_maps(verts, faces, verts, R, T, K, H, W)
        ev[1].record()
        torch.cuda.synchronize()
        print(f'iter {i}: {ev[0].elapsed_time(ev[1]):.3f} ms  P={P} verts={verts.shape[0]} faces={faces.shape[0]} psnr={res[1].item():.3f} '
              f'covered={[float((m[0] != m[0, 0, 0]).any(-1).float().mean()) for m in maps.values()]}')


if __name__ == '__main__':
    main()
