#!/usr/bin/env python
"""BASELINE configs[3]: a novel-pose sequence (default 258 synthetic frames, AIST++ demo length) sharded frame-wise over the
GPUs of one node: one process per GPU, one broadcast of the frame-invariant weights, frame i -> rank i mod N,
no collective on the data path, one gather of the uint8 images at the end.

    python tools/render_sequence.py --frames 16 --size 256                                             # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 \
        tools/render_sequence.py --frames 258 --size 512

Prints one JSON line on rank 0: frames, rays, seconds (max over ranks, CUDA events), rays/s, frames/s.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=258)
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--gather', action='store_true', help='gather the uint8 images on rank 0 at the end')
    ap.add_argument('--normal-maps', type=int, default=0, metavar='N', help='also extract the canonical mesh on an N^3 lattice and rasterise '
                    'the three normal maps per frame (test.py: gen_cano_mesh=True, models/__init__.py:203-309); 0 = off')
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from arah_release_b200 import sharding as sh, synthetic as syn
    from tools import ref_layout as rl
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    torch.cuda.set_device(local)
    dev = torch.device(f'cuda:{local}')
    mine = sh.frames_for_rank(a.frames, rank, world)
    f0 = syn.make_frame(a.size, a.size, seed=0, frame_idx=0)
    dvn, rend, skin, _ = rl.modules_from_frame(f0, dev)
    sh.broadcast_module_weights([dvn, rend, skin])
    net = IDHRNetwork(dvn, rend, skin, BodyRayTracing(n_steps=f0.n_steps), cano_view_dirs=f0.cano_view_dirs).eval()
    # frames are prepared on the host up front (the dataset's job); the timed region is the device work of the shard
    frames = [syn.make_frame(a.size, a.size, seed=0, frame_idx=i) for i in mine]
    inputs = [rl.inputs_from_frame(f, rl.sdf_network_from_frame(f, dev), dev) for f in frames]
    if inputs:
        net(inputs[0])                                            # warm-up (workspace allocation)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    images, rays = {}, 0
    for fi, f, inp in zip(mine, frames, inputs):
        out = net(inp)
        rays += f.P
        if a.normal_maps:
            inp = dict(inp, cam_rot=torch.from_numpy(f.pose[:3, :3].copy()).view(1, 3, 3), cam_trans=torch.from_numpy(f.pose[:3, 3].copy()).view(1, 3),
                       intrinsics=torch.from_numpy(f.K).view(1, 3, 3))
            out.update(net.render_normal_maps(inp, N=a.normal_maps, image_size=(f.H, f.W)))
        if a.gather:
            images[fi] = sh.to_image_u8(out['rgb_values'][0], torch.from_numpy(f.pix).to(dev), f.H, f.W)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) / 1e3
    if world > 1:
        t = torch.tensor([secs], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n = torch.tensor([float(rays)], device=dev, dtype=torch.float64); dist.all_reduce(n, op=dist.ReduceOp.SUM)
        secs, rays = float(t[0]), float(n[0])
    gathered = sh.gather_frames(images, a.frames, H=a.size, W=a.size, device=dev) if a.gather else None
    if rank == 0:
        print(json.dumps({'workload': f'{a.frames}-frame synthetic novel-pose sequence at {a.size}x{a.size}, frames sharded over {world} GPU(s)',
                          'frames': a.frames, 'rays': int(rays), 'seconds': secs, 'rays_per_s': rays / secs, 'frames_per_s': a.frames / secs,
                          'normal_maps_lattice': a.normal_maps, 'n_gpus': world, 'gathered_images': None if gathered is None else len(gathered)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
