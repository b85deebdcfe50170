"""TEST INFRASTRUCTURE — golden vectors for the ray set-up row (f3) from the UNMODIFIED reference functions.

    python oracle/gen_golden_rays.py          # build container only (/root/reference must exist)

Calls the reference's own `get_bound_2d_mask` / `get_near_far` (im2mesh/utils/utils.py:43-73, the former through the real
cv2.fillPoly) and the dataset class's own methods `init_grid_homo_2d`, `normalize_vectors`, `get_camera_rays`,
`get_camera_location` (im2mesh/data/zju_mocap_odp.py:137-178, called unbound on a bare object: they use no instance state),
glued exactly as `__getitem__` does (:285-315; the dataset itself needs SMPL model files that are absent here).
-> tests/golden/rays_s{seed}.npz: inputs (K, R, T, bounds, H, W) and outputs (bit-packed bound mask, pixel list, rays, near/far).
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh        # noqa: E402


from arah_release_b200.synthetic import make_camera as camera   # noqa: E402  (seeded synthetic camera + body box)


def main():
    rh.install()
    from im2mesh.utils.utils import get_bound_2d_mask, get_near_far
    from im2mesh.data.zju_mocap_odp import ZJUMOCAPODPDataset as DS
    obj = object.__new__(DS)
    for seed, (H, W, zoom) in enumerate([(128, 128, 1.0), (96, 160, 1.0), (128, 128, 2.2), (512, 512, 1.0)]):
        K, R, T, bounds = camera(seed, H, W, zoom)
        cam_loc = DS.get_camera_location(obj, R, T)                                            # :216
        K_inv = np.linalg.inv(K)                                                               # :231
        homo_2d = DS.init_grid_homo_2d(obj, H, W)                                              # :94
        bound_mask = get_bound_2d_mask(bounds, K, np.concatenate([R, T.reshape([3, 1])], axis=-1), H, W)   # :291 (img_size[0], img_size[1])
        y_inds, x_inds = np.where(bound_mask != 0)                                             # :292
        sampled_uv = np.dot(homo_2d.copy()[y_inds, x_inds].reshape([-1, 3]), K_inv.T)          # :298
        sampled_rays = DS.get_camera_rays(obj, R, sampled_uv)                                  # :300
        near, far, mask_at_box = get_near_far(bounds, np.broadcast_to(cam_loc, sampled_rays.shape), sampled_rays)   # :302
        g = {'H': np.array(H), 'W': np.array(W), 'K': K, 'R': R, 'T': T, 'bounds': bounds, 'cam_loc': cam_loc.astype(np.float32),
             'bound_mask_bits': np.packbits(bound_mask.astype(np.uint8)),
             'pix': (y_inds[mask_at_box] * W + x_inds[mask_at_box]).astype(np.int32),
             'ray_dirs': sampled_rays[mask_at_box].astype(np.float32),
             'near_far': np.stack([near[mask_at_box], far[mask_at_box]], axis=-1).astype(np.float32),
             'n_bound': np.array(int((bound_mask != 0).sum()))}
        if H * W > 128 * 128:          # keep the fixture small: a seeded sample of the rays + checksums
            idx = np.random.default_rng(7).choice(g['pix'].shape[0], size=4096, replace=False)
            idx.sort()
            g['sample_idx'] = idx.astype(np.int32)
            g['pix_sum'] = np.array(int(g['pix'].astype(np.int64).sum())); g['n_rays'] = np.array(g['pix'].shape[0])
            for k in ('pix', 'ray_dirs', 'near_far'):
                g[k] = g[k][idx]
        path = os.path.join(ROOT, 'tests', 'golden', f'rays_s{seed}.npz')
        np.savez_compressed(path, **g)
        print('wrote', path, os.path.getsize(path), 'bytes; bound', int(g['n_bound']), 'rays', int(mask_at_box.sum()),
              'corner outside image:', bool(((np.where(bound_mask)[0].min() == 0) or (np.where(bound_mask)[0].max() == H - 1))))


if __name__ == '__main__':
    main()
