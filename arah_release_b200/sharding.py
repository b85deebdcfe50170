"""Frame-level sharding of a sequence across ranks.

Frames are independent given the network weights, so the multi-GPU path is: one process per GPU, a single broadcast of
the frame-invariant module weights from rank 0, frame i -> rank i mod world (what DistributedSampler does under the
reference's `strategy='ddp'`, /root/reference/test.py:71), no collective on the data path, and one gather of the
uint8 images at the end (the reference all_gathers fp32 stacks, lightning_model.py:357-360).
Backend-agnostic (nccl on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def frames_for_rank(n_frames: int, rank: int, world: int):
    return list(range(rank, n_frames, world))


def broadcast_module_weights(modules, src: int = 0):
    """The one collective of the render path: every parameter/buffer of the frame-invariant modules from `src`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    n = 0
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src)
            n += t.numel()
    return n


def to_image_u8(rgb, pix, H, W):
    """Scatter per-ray rgb [P,3] (float 0..1) into a uint8 image [H,W,3] (background 0).  Conversion as the reference's
    `(x * 255.0).astype(np.uint8)` (lightning_model.py:388): truncation, after the clamp the scattered images already carry."""
    img = torch.zeros(H * W, 3, dtype=torch.uint8, device=rgb.device)
    img[pix] = (rgb.clamp(0, 1) * 255.0).to(torch.uint8)
    return img.view(H, W, 3)


def gather_frames(local_images: dict, n_frames: int, dst: int = 0, H=None, W=None, device=None):
    """local_images: {frame_idx: uint8 [H,W,3]} on this rank -> on `dst` a list of n_frames images (None elsewhere).
    EVERY rank enters the collective, also one that owns no frame (n_frames < world): pass H, W (and the device) in that case."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local_images[i] for i in range(n_frames)]
    world, rank = dist.get_world_size(), dist.get_rank()
    if local_images:
        any_img = next(iter(local_images.values()))
        shape, device = tuple(any_img.shape), any_img.device
    else:
        if H is None or W is None:
            raise ValueError('gather_frames: a rank without frames must be given H and W')
        shape = (H, W, 3)
        device = device if device is not None else torch.device('cpu')
    per_rank = (n_frames + world - 1) // world
    buf = torch.zeros((per_rank,) + shape, dtype=torch.uint8, device=device)
    for k, fi in enumerate(frames_for_rank(n_frames, rank, world)):
        buf[k] = local_images[fi]
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst)
    if rank != dst:
        return None
    res = [None] * n_frames
    for r in range(world):
        for k, fi in enumerate(frames_for_rank(n_frames, r, world)):
            res[fi] = outs[r][k]
    return res
