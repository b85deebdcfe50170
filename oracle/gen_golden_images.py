"""TEST INFRASTRUCTURE — golden fixtures for the validation-image tail (tests/golden/images_s*.npz).

Runs the UNMODIFIED `LightningModel.validation_step` of the reference
(/root/reference/im2mesh/metaavatar_render/lightning_model.py:158-229) on seeded synthetic renderer outputs: `self` is a
stand-in whose `model(...)` returns the canned `rgb_values` / `points_cam`, and `ssim_metric` / `lpips_metric` (skimage / lpips:
absent here, not part of this row) are replaced by constants.  Everything between — masked_scatter_, the finite-difference
normal map, `psnr_metric` — is the reference's own code executed by torch / numpy on the CPU.

    python -m oracle.gen_golden_images            (build container only; needs /root/reference)
"""
import os
import types

import numpy as np
import torch

from oracle import ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def synth_case(seed, H, W):
    """A blob-shaped image mask, a smooth depth surface seen by a pinhole camera, rays that hit nothing (points_cam == 0),
    and — on purpose — neighbouring pixels with identical x or y (division by zero) to exercise the inf / NaN handling."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    cy, cx = H * (0.45 + 0.1 * rng.random()), W * (0.45 + 0.1 * rng.random())
    r = np.hypot((yy - cy) / (0.42 * H), (xx - cx) / (0.36 * W))
    mask = r < 1.0 + 0.08 * np.sin(5 * np.arctan2(yy - cy, xx - cx))
    hit = r < 0.8
    f = np.float32(1.2 * max(H, W))
    z = (3.0 - 0.6 * np.sqrt(np.maximum(0.0, 1.0 - r ** 2)) + 0.02 * rng.standard_normal((H, W))).astype(np.float32)
    pts = np.stack([(xx - W / 2) / f * z, (yy - H / 2) / f * z, z], -1).astype(np.float32)
    # degenerate neighbours: same x as the right neighbour / same y as the lower neighbour at a few pixels
    for _ in range(6):
        y, x = int(rng.integers(1, H - 2)), int(rng.integers(1, W - 2))
        pts[y, x, 0] = pts[y, x + 1, 0]
        pts[y, x, 1] = pts[y + 1, x, 1]
    y, x = int(cy), int(cx)
    pts[y, x] = pts[y, x + 1]                      # 0 / 0 inside the body
    pts[~hit] = 0.0                                # the renderer zeroes points_cam of rays without a surface (row a11)
    pix = np.flatnonzero(mask.reshape(-1)).astype(np.int32)
    rgb = rng.random((len(pix), 3)).astype(np.float32)
    gt = np.clip(rgb + 0.05 * rng.standard_normal(rgb.shape), 0, 1).astype(np.float32)
    return {'H': H, 'W': W, 'mask': mask, 'pix': pix, 'rgb': rgb, 'gt': gt, 'points_cam': pts.reshape(-1, 3)[pix]}


def run_reference(case):
    rh.install()
    import im2mesh.metaavatar_render  # noqa: F401
    from im2mesh.metaavatar_render import lightning_model as lm
    lm.ssim_metric = lambda *a, **k: 0.0
    lm.lpips_metric = lambda *a, **k: 0.0
    P = len(case['pix'])
    outputs = {'rgb_values': torch.from_numpy(case['rgb']).view(1, P, 3), 'points_cam': torch.from_numpy(case['points_cam']).view(1, P, 3)}
    fake = types.SimpleNamespace(compose_inputs=lambda batch, eval=True: {}, model=lambda inputs, gen_cano_mesh=False, eval=True: outputs,
                                 device=torch.device('cpu'), loss_fn_vgg=None)
    batch = {'inputs.img_height': torch.tensor([case['H']]), 'inputs.img_width': torch.tensor([case['W']]),
             'inputs.image_mask': torch.from_numpy(case['mask']).view(1, -1), 'inputs': torch.from_numpy(case['gt']).view(1, P, 3)}
    ev = lm.LightningModel.validation_step(fake, batch, 0)
    return {'psnr': np.float64(ev['psnr']), 'rgb_pred': ev['rgb_pred'].permute(1, 2, 0).numpy(), 'normal_pred': ev['normal_pred'].permute(1, 2, 0).numpy(),
            'rgb_gt': ev['rgb_gt'].permute(1, 2, 0).numpy()}


def main():
    for seed, (H, W) in enumerate([(24, 20), (48, 64), (96, 96)]):
        case = synth_case(seed, H, W)
        ref = run_reference(case)
        path = os.path.join(OUT, f'images_s{seed}.npz')
        np.savez_compressed(path, H=H, W=W, mask=case['mask'], pix=case['pix'], rgb=case['rgb'], gt=case['gt'], points_cam=case['points_cam'],
                            **{'ref.' + k: v for k, v in ref.items()})
        print(path, 'P =', len(case['pix']), 'psnr =', float(ref['psnr']), 'normals in', ref['normal_pred'].min(), ref['normal_pred'].max())


if __name__ == '__main__':
    main()
