#!/usr/bin/env python
"""bench.py — rays/sec of the ARAH hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

A *step* is one frame: one pass of the hot path (set_frame weight packing + sphere tracing + joint root finding +
sample correspondences + SDF/colour shading + compositing) over all bbox rays of a synthetic 512x512 ZJU-377-like
frame (BASELINE.json configs[1]; synthetic data, fitted synthetic SDF/skinning nets, random-init colour net — no dataset
or checkpoint is reachable offline).  N > 1: frames of a sequence shard across ranks (weak scaling: one frame per rank
per step, no data-path collective; a single NCCL broadcast of the frame-invariant weights at start).

JSON keys follow the driver contract; `value` = device-resident inputs (CUDA events), `e2e` = the same frame through the
host-buffer C-ABI entry point (arah_set_frame with host pose buffers + arah_render_host: H2D of rays/pose and D2H of
rgb/mask/points inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAC_SDF, MAC_SKIN, MAC_COL = 328704, 52736, 345344     # per-evaluation MACs (colour: latent folded)
H = W = 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--size', type=int, default=512, help='image side (default = BASELINE configs[1]); smaller is for debugging only')
    ap.add_argument('--cpu-sample-seconds', type=float, default=15.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train-step', action='store_true', help='skip the BASELINE configs[2] training-step measurement')
    ap.add_argument('--no-mesh', action='store_true', help='skip the canonical-mesh (row f1) measurement')
    ap.add_argument('--train-rays', type=int, default=2048, help='rays per training step (configs/default.yaml:13-14: 1024 fg + 1024 bg)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d['bf16_tflops_sustained'], 'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'src': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': float(max(mx)) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_frames(size, n_frames, first):
    from arah_release_b200 import synthetic as syn
    return [syn.make_frame(size, size, seed=0, frame_idx=first + i) for i in range(n_frames)]


def algorithmic_flops(st):
    """MACs defined through the iteration counters; FLOPs = 2 x MACs."""
    mac = (st['trace_sdf_evals'] * MAC_SDF + st['iso_rays'] * (4 * MAC_SKIN + 2 * MAC_SDF) + st['iso_g_evals'] * (MAC_SKIN + MAC_SDF)
           + st['corr_skin_evals'] * MAC_SKIN + st['shaded_samples'] * (2 * MAC_SDF + MAC_COL))
    return 2.0 * mac


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_rate(frame, seconds, threads=0):
    """rays/s of the oracle port (oracle/arah_oracle.c, all host threads) on a bounded random ray sample of the frame."""
    from oracle import oracle as orc
    nthr = orc.num_threads() if threads <= 0 else threads
    rng = np.random.default_rng(0)
    n0 = min(frame.P, 64 * nthr)
    sel = rng.choice(frame.P, size=n0, replace=False)
    t = time.perf_counter()
    orc.render(frame, ray_dirs=frame.ray_dirs[sel], near_far=frame.near_far[sel], stages=False, threads=nthr)
    r0 = n0 / (time.perf_counter() - t)
    n = int(min(frame.P, max(n0, r0 * seconds)))
    sel = rng.choice(frame.P, size=n, replace=False)
    t = time.perf_counter()
    o = orc.render(frame, ray_dirs=frame.ray_dirs[sel], near_far=frame.near_far[sel], stages=False, threads=nthr)
    dt = time.perf_counter() - t
    return n / dt, nthr, n, dt, sel, o


def parity_on_sample(ours_rgb, ours_mask, sel, o):
    """BASELINE metric, second half ('PSNR delta vs reference'): the GPU frame against the CPU port on the timed ray sample.
    delta = PSNR(ours, gt) - PSNR(port, gt) against a fixed pseudo ground truth (port + N(0, 0.03) noise, seed 123), bar 0.05 dB."""
    ref = np.asarray(o['rgb_values'], np.float64)
    mine = np.asarray(ours_rgb[sel], np.float64)
    ps = lambda a, b: float(-10.0 * np.log10(max(float(np.mean((a - b) ** 2)), 1e-30)))
    gt = np.clip(ref + np.random.default_rng(123).normal(scale=0.03, size=ref.shape), 0.0, 1.0)
    return {'rays': int(len(sel)), 'psnr_vs_cpu_port_db': ps(mine, ref), 'delta_psnr_vs_pseudo_gt_db': ps(mine, gt) - ps(ref, gt), 'bar_db': 0.05,
            'mask_agreement': float(np.mean(np.asarray(ours_mask[sel]).astype(bool) == np.asarray(o['network_body_mask']).astype(bool))),
            'note': 'same frame, same rays: CUDA path (device-resident render) vs oracle/arah_oracle.c (pinned to the unmodified reference by tests/golden)'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import oracle as orc
    frame = make_frames(args.size, 1, 0)[0]
    nthr = orc.num_threads()
    rng = np.random.default_rng(0)
    # bounded sample per step, sized so that the whole run stays within a few minutes
    t = time.perf_counter()
    n0 = min(frame.P, 32 * nthr)
    sel = rng.choice(frame.P, size=n0, replace=False)
    orc.render(frame, ray_dirs=frame.ray_dirs[sel], near_far=frame.near_far[sel], stages=False, threads=nthr)
    r0 = n0 / (time.perf_counter() - t)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    n = int(min(frame.P, max(n0, r0 * min(budget, 20.0))))
    times = []
    for i in range(args.warmup + args.steps):
        sel = rng.choice(frame.P, size=n, replace=False)
        t = time.perf_counter()
        orc.render(frame, ray_dirs=frame.ray_dirs[sel], near_far=frame.near_far[sel], stages=False, threads=nthr)
        if i >= args.warmup:
            times.append(time.perf_counter() - t)
    tot = float(sum(times))
    val = n * args.steps / tot
    line = {'impl': 'reference', 'metric': 'rays/sec at 512x512 ZJU-377 render', 'value': val, 'unit': 'rays/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * tot / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'ZJU-377-like {args.size}x{args.size} novel-view render, fp32 (BASELINE configs[1])', 'rays_per_frame': frame.P,
                       'n_steps': frame.n_steps, 'near_far_samples': [frame.near_samples, frame.far_samples]},
            'cpu_baseline': {'value': val, 'unit': 'rays/s', 'cores': nthr, 'kind': 'port',
                             'sample': f'{n} random bbox rays of the frame per step; oracle/arah_oracle.c (C restatement of the reference '
                                       f'algorithm, OpenMP over rays); the Python reference cannot travel to this box'},
            'e2e': {'value': val, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ training step (configs[2])
def train_step_bench(args, dev, frame, steps=5, warmup=3, variant='torch'):
    """One training step = IDHRNetwork.forward (training branch: tracer with jitter, differentiable shading, regulariser
    evaluations) + IDHRLoss-style loss + backward to every parameter tensor, on `--train-rays` random bbox rays of the frame
    (BASELINE configs[2]: 1024 + 1024 rays, train_skinning_net).  Secondary metric; the headline stays the 512x512 render."""
    import torch
    from arah_release_b200 import ref_layout as rl, synthetic as syn
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(frame.P, size=min(args.train_rays, frame.P), replace=False))
    import copy
    fr = copy.copy(frame)
    fr.ray_dirs, fr.near_far, fr.pix = frame.ray_dirs[sel], frame.near_far[sel], frame.pix[sel]
    aux = syn.train_aux_points(fr, seed=0)
    dvn, rend, skin, sdf = rl.modules_from_frame(fr, dev)
    leaves = []
    for l in range(6):
        for n in ('weights', 'biases', 'freq', 'phase_shift'):
            v = getattr(sdf[l][0], n).clone().requires_grad_(True); setattr(sdf[l][0], n, v); leaves.append(v)
    for n in ('weights', 'biases'):
        v = getattr(sdf[6], n).clone().requires_grad_(True); setattr(sdf[6], n, v); leaves.append(v)
    net = IDHRNetwork(dvn, rend, skin, BodyRayTracing(n_steps=fr.n_steps, near_surface_vol_samples=fr.near_samples,
                                                      far_surface_vol_samples=fr.far_samples), cano_view_dirs=fr.cano_view_dirs,
                      train_skinning_net=True).train()
    inp = rl.inputs_from_frame(fr, sdf, dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    inp['pose_cond']['latent_code'] = inp['pose_cond']['latent_code'].clone().requires_grad_(True)
    inp['body_mask'] = t(aux['body_mask']).view(1, -1)
    inp['points_uniform'] = t(aux['points_uniform']).float().view(1, -1, 3)
    inp['points_skinning'] = t(aux['points_skinning']).float().view(1, -1, 3)
    inp['points_inside'] = t(aux['points_inside']).float().view(1, -1, 3)
    gt, wt, body = t(aux['rgb_gt']).float(), t(aux['sampled_weights']).float(), t(aux['body_mask'])
    P = fr.P
    params = leaves + [inp['pose_cond']['latent_code']] + list(net.parameters())

    def loss_fn(out):            # IDHRLoss.forward, rgb 'l1' (/root/reference/im2mesh/metaavatar_render/renderer/loss.py:122-200)
        vm = out['network_body_mask'][0]
        l = (out['rgb_values'][0][vm] - gt[vm]).abs().sum() / P
        l = l + torch.norm(out['sdf_output'][0][vm] - body[vm].float(), dim=-1).sum() / P
        l = l + 0.1 * (out['grad_theta'].norm(2, dim=-1) - 1).abs().sum() / P
        l = l + 0.01 * torch.exp(-1e2 * out['off_surface_sdf']).sum() / P + 0.01 * torch.sigmoid(out['inside_sdf'] * 5e3).sum() / P
        return l + 10.0 * (out['pred_weights'][0] - wt).abs().sum(-1).mean()

    if variant == 'fused':
        # the same step with the fused criterion (arah_idhr_loss: terms + gradients in four launches, no host round trips) in place
        # of the ~60 small torch kernels of loss_fn; run as the LAST secondary entry (first GPU run of these kernels)
        from arah_release_b200.loss import IDHRLoss
        crit = IDHRLoss(rgb_weight=1.0, perceptual_weight=0.0, eikonal_weight=0.1, mask_weight=1.0, off_surface_weight=0.01, inside_weight=0.01,
                        params_weight=0.0, skinning_weight=10.0, rgb_loss_type='l1')
        gtd = {'rgb': gt.view(1, -1, 3), 'sampled_weights': wt.view(1, -1, 24)}
        msf = []
        for i in range(warmup + steps):
            for p_ in params:
                p_.grad = None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lf = crit(net(inp), gtd)['loss'].sum()
            lf.backward()
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                msf.append(e0.elapsed_time(e1))
        return {'ms_per_step_fused_loss': float(np.mean(msf)), 'loss_fused': float(lf.detach())}
    ms = {'forward': [], 'backward': [], 'total': []}
    launches, samples = 0, 0
    for i in range(warmup + steps):
        for p_ in params:
            p_.grad = None
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out = net(inp)
        loss = loss_fn(out)
        e[1].record()
        loss.backward()
        e[2].record()
        torch.cuda.synchronize()
        if i >= warmup:
            ms['forward'].append(e[0].elapsed_time(e[1])); ms['backward'].append(e[1].elapsed_time(e[2])); ms['total'].append(e[0].elapsed_time(e[2]))
        st = net.stats()
        launches, samples = st['kernel_launches'], st['shaded_samples']
    tot = float(np.mean(ms['total']))
    # algorithmic MACs of the differentiable part per sample: SDF fwd + input-gradient sweep + their second- and first-order
    # backward (4 data GEMM sweeps + 2 weight-gradient sweeps), colour net fwd + data + weight gradient, skinning net value +
    # 3 tangents + backward
    mac = samples * (6 * MAC_SDF + 3 * (MAC_COL + 128 * 256) + 6 * MAC_SKIN)
    return {'workload': f'{P} rays of the ZJU-377-like frame, train_skinning_net, fp32 (BASELINE configs[2] shape)', 'rays': int(P),
            'shaded_samples': int(samples), 'ms_per_step': tot, 'ms_forward_incl_tracer': float(np.mean(ms['forward'])),
            'ms_backward': float(np.mean(ms['backward'])), 'rays_per_s': P / tot * 1e3, 'steps': steps, 'warmup': warmup,
            'gpu_launches_per_step': int(launches), 'loss': float(loss.detach()),
            'algorithmic_tflops_differentiable_part': 2.0 * mac / (tot * 1e-3) / 1e12}


# ------------------------------------------------------------------------------------------------ canonical mesh (row f1)
def mesh_extract_bench(net, frame, steps=3, warmup=2, N=256):
    """MetaAvatarRender.forward(gen_cano_mesh=True) front half (models/__init__.py:203-206): 256^3 SDF lattice + iso-surface.
    SDF lattice: tensor roofline (N^3 x 328 704 MAC, 3xTF32 counts as 1x useful); marching cubes: HBM roofline, algorithmic
    bytes = one read of the lattice (4 N^3) + the vertices and faces written."""
    import torch
    from oracle import oracle as orc
    r = net._last[0]
    ms_g, ms_m = [], []
    for i in range(warmup + steps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        vol = r.sdf_grid(N)
        e[1].record()
        v, f = r.marching_cubes(vol, max_verts=1 << 20, max_faces=1 << 21)
        e[2].record()
        torch.cuda.synchronize()
        if i >= warmup:
            ms_g.append(e[0].elapsed_time(e[1])); ms_m.append(e[1].elapsed_time(e[2]))
    pk = peaks()
    g, m = float(np.mean(ms_g)), float(np.mean(ms_m))
    mc_bytes = 4.0 * N ** 3 + 12.0 * v.shape[0] + 12.0 * f.shape[0]
    # CPU: the oracle's marching cubes on the same lattice (1 core) and its SDF on a bounded lattice sample (all cores)
    volh = vol.cpu().numpy()
    t = time.perf_counter(); orc.marching_cubes(volh); t_mc = time.perf_counter() - t
    pts, _ = orc.grid_points(64)
    t = time.perf_counter(); orc.sdf(frame, pts, grad=False); t_sdf = time.perf_counter() - t
    return {'workload': f'{N}^3 canonical SDF lattice + iso-surface (utils/sdf_meshing.py:13-114)', 'ms_sdf_grid': g, 'ms_marching_cubes': m,
            'n_verts': int(v.shape[0]), 'n_faces': int(f.shape[0]),
            'sdf_grid_tflops_useful': 2.0 * MAC_SDF * N ** 3 / (g * 1e-3) / 1e12, 'sdf_grid_frac_of_bf16_peak': 2.0 * MAC_SDF * N ** 3 / (g * 1e-3) / 1e12 / pk['tf_sustained'],
            'marching_cubes_gbs': mc_bytes / (m * 1e-3) / 1e9, 'marching_cubes_frac_of_hbm_peak': mc_bytes / (m * 1e-3) / 1e9 / pk['hbm_gbs'],
            'gpu_launches': 5, 'steps': steps, 'warmup': warmup,
            'cpu_port': {'ms_marching_cubes_1core': 1e3 * t_mc, 'ms_sdf_grid_extrapolated': 1e3 * t_sdf * (N / 64.0) ** 3, 'cores': orc.num_threads(),
                         'sample': '64^3 lattice points timed, scaled by (N/64)^3'}}


# ------------------------------------------------------------------------------------------------ hypernetwork (row f4)
def hypernet_bench(dev, steps=20, warmup=5):
    """HyperBVPNet.forward up to the assembled decoder (siren_modules.py:280-312): batch-1 GEMV over 86.9 M parameters.
    HBM roofline: algorithmic bytes = every parameter read once (the caches are flushed between timed calls)."""
    import torch
    from arah_release_b200 import synthetic as syn
    from arah_release_b200.hypernet import HyperSDFDecoder
    from oracle import hyper_oracle as ho
    sd = syn.make_hypernet_state_dict(0)
    dec = HyperSDFDecoder({k: torch.from_numpy(v) for k, v in sd.items()}, dev)
    rots, Jtrs, latent = syn.make_hypernet_inputs(0)
    inp = {'rots': torch.from_numpy(rots).to(dev), 'Jtrs': torch.from_numpy(Jtrs).to(dev), 'latent': torch.from_numpy(latent).to(dev)}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ms = []
    for i in range(warmup + steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dec(inp); e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ms.append(e0.elapsed_time(e1))
    # the same two launches through the bare C-ABI call with prepared buffers
    from arah_release_b200._lib import ArahSdfParams
    import ctypes as C
    res = dec(inp)
    d = res['decoder']
    outs = ArahSdfParams()
    for l in range(7):
        lay = d[l][0] if l < 6 else d[l]
        outs.sdf_W[l], outs.sdf_b[l] = C.c_void_p(lay.weights.data_ptr()), C.c_void_p(lay.biases.data_ptr())
    fq, ph = torch.empty(6, 256, device=dev), torch.empty(6, 256, device=dev)
    outs.sdf_freq, outs.sdf_phase = C.c_void_p(fq.data_ptr()), C.c_void_p(ph.data_ptr())
    r_, j_, l_ = inp['rots'].reshape(-1).contiguous(), inp['Jtrs'].reshape(-1).contiguous(), inp['latent'].reshape(-1).contiguous()
    ms_raw = []
    for i in range(warmup + steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dec.launch_raw(r_, j_, l_, outs); e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ms_raw.append(e0.elapsed_time(e1))
    t = time.perf_counter(); ho.forward(sd, rots, Jtrs, latent); t_cpu = time.perf_counter() - t
    pk = peaks()
    m_mirror = float(np.median(ms))
    m = float(np.median(ms_raw))
    gbs = dec.weight_bytes / (m * 1e-3) / 1e9
    return {'workload': 'MetaAvatar hypernetwork forward, 86.9 M parameters, batch 1 (configs/arah-zju/ZJUMOCAP-377_4gpus.yaml:34)',
            'ms': m, 'ms_through_python_mirror': m_mirror, 'algorithmic_bytes': int(dec.weight_bytes), 'achieved_gbs': gbs, 'peak_gbs': pk['hbm_gbs'], 'frac_of_hbm_peak': gbs / pk['hbm_gbs'],
            'gpu_launches': 2, 'steps': steps, 'warmup': warmup, 'l2': '256 MB memset before every timed call',
            'cpu_port': {'ms': 1e3 * t_cpu, 'kind': 'oracle/hyper_oracle.py (numpy fp32, BLAS threads)'}}


# ------------------------------------------------------------------------------------------------ ray set-up (row f3)
def ray_setup_bench(dev, size, steps=10, warmup=3):
    """data/zju_mocap_odp.py:250-315 per frame: SMPL posing + bounding box, box mask, rays, near/far, ordered compaction.
    HBM roofline: algorithmic bytes = blend-shape basis + weights + shape read once, vertices written; per pixel one mask byte
    written and read, per ray 24 B written."""
    import torch
    from arah_release_b200 import synthetic as syn
    from arah_release_b200.rays import FrameRays
    from oracle import rays_oracle as ro
    p = syn.make_smpl_pose_inputs(0)
    K, R, T, _ = syn.make_camera(3, size, size, 1.0)
    fr = FrameRays(dev)
    dp = {k: (torch.as_tensor(v, dtype=torch.float32).to(dev) if k not in ('pose_feature', 'trans') else v) for k, v in p.items()}
    ms_pose, ms_rays, P = [], [], 0
    for i in range(warmup + steps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        verts, bounds = fr.pose_smpl(**dp)
        e[1].record()
        out = fr.gen_rays(K, R, T, bounds, size, size)
        e[2].record()
        torch.cuda.synchronize()
        if i >= warmup:
            ms_pose.append(e[0].elapsed_time(e[1])); ms_rays.append(e[1].elapsed_time(e[2]))
        P = out['n_rays']
    t = time.perf_counter(); v_ref, b_ref = ro.pose_smpl(**p); t_pose = time.perf_counter() - t
    t = time.perf_counter(); ref = ro.gen_rays(K, R, T, b_ref, size, size); t_rays = time.perf_counter() - t
    pk = peaks()
    mp, mr = float(np.median(ms_pose)), float(np.median(ms_rays))
    pose_bytes = 6890 * (3 * 207 * 4 + 24 * 4 + 12 + 12) + 207 * 8 + 24 * 64
    ray_bytes = 2.0 * size * size + P * (12 + 8 + 4) + size * size
    return {'workload': f'per-frame ray set-up at {size}x{size} (data/zju_mocap_odp.py:250-315)', 'n_rays': int(P), 'n_rays_cpu_port': int(ref['pix'].shape[0]),
            'ms_pose_smpl': mp, 'pose_smpl_gbs': pose_bytes / (mp * 1e-3) / 1e9, 'ms_frame_rays': mr, 'frame_rays_gbs': ray_bytes / (mr * 1e-3) / 1e9,
            'frac_of_hbm_peak': {'pose_smpl': pose_bytes / (mp * 1e-3) / 1e9 / pk['hbm_gbs'], 'frame_rays': ray_bytes / (mr * 1e-3) / 1e9 / pk['hbm_gbs']},
            'gpu_launches': 3 + 6, 'steps': steps, 'warmup': warmup,
            'note': 'small launch-latency-bound kernels (9 launches, one host read of the ray count); timed through the Python mirror',
            'cpu_port': {'ms_pose_smpl': 1e3 * t_pose, 'ms_frame_rays': 1e3 * t_rays, 'kind': 'oracle/rays_oracle.py (numpy; python loops for the mask)'}}


# ------------------------------------------------------------------------------------------------ image tail (rows f4 / f1)
def image_tail_bench(net, frame, inp, steps=10, warmup=3, N=256):
    """What follows the renderer in validation_step / gen_cano_mesh (lightning_model.py:176-221, models/__init__.py:226-309):
    scatter rgb / points into images + finite-difference normal map, PSNR, and the three rasterised normal maps of the extracted
    mesh.  HBM roofline; algorithmic bytes: every input read once and every output written once (z-buffer keys: one clear, one
    resolve read)."""
    import torch
    from arah_release_b200.images import FrameImages
    from oracle import images_oracle as io
    out = net(inp)
    dev = out['rgb_values'].device
    fi = FrameImages(dev)
    H, W = frame.H, frame.W
    pix = torch.from_numpy(frame.pix.astype(np.int32)).to(dev)
    rgb, pts = out['rgb_values'][0].contiguous(), out['points_cam'][0].contiguous()
    gt = (rgb + 0.02 * torch.randn_like(rgb)).clamp(0, 1)
    mesh = net.extract_canonical_mesh(inp, N=N)
    verts, faces, posed = mesh
    R, T, K = frame.pose[:3, :3], frame.pose[:3, 3], frame.K
    mask_img = torch.zeros(H * W, dtype=torch.uint8, device=dev)
    mask_img[pix.long()] = 1
    gtimg, _ = fi.assemble(gt, None, pix, H, W, normals=False)
    ms = [[], [], [], []]
    for i in range(warmup + steps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record()
        pp, pn = fi.assemble(rgb, pts, pix, H, W)
        e[1].record()
        res = fi.psnr_device(rgb, gt)
        e[2].record()
        maps = fi.normal_maps(verts, faces, posed, R, T, K, H, W)
        e[3].record()
        ss = fi.ssim_device(pp, gtimg, mask_img)
        e[4].record()
        torch.cuda.synchronize()
        if i >= warmup:
            for k in range(4):
                ms[k].append(e[k].elapsed_time(e[k + 1]))
    m_img, m_psnr, m_maps, m_ssim = (float(np.median(x)) for x in ms)
    P, V, Fc, n = int(pix.numel()), int(verts.shape[0]), int(faces.shape[0]), H * W
    b_img = P * (12 + 12 + 4 + 24) + n * (12 + 12) + n * (12 + 12)          # rows in, scattered out, 2 clears, normals read + write
    b_psnr = 2 * 12 * P
    b_maps = 3 * (24 * V + 12 * Fc + 36 * Fc + 8 * n + 12 * n + 4 * n + 12 * n)
    pk = peaks()
    # CPU port (numpy; the rasteriser loops over faces in Python: a bounded face sample of one view, scaled to 3 views x all faces)
    rgb_h, pts_h, gt_h, pix_h = rgb.cpu().numpy(), pts.cpu().numpy(), gt.cpu().numpy(), frame.pix
    t = time.perf_counter(); io.frame_images(rgb_h, pts_h, pix_h, H, W); t_img = time.perf_counter() - t
    t = time.perf_counter(); io.psnr_metric(rgb_h, gt_h); t_psnr = time.perf_counter() - t
    vh, fh = posed.cpu().numpy(), faces.cpu().numpy()
    ns = min(Fc, 4000)
    t = time.perf_counter(); io.rasterize(io.project(vh, io.opencv_camera(R, T, K, H, W)), fh[:ns], H, W); t_r = time.perf_counter() - t
    return {'workload': f'validation images + PSNR at {H}x{W} ({P} rays), 3 normal maps of the {N}^3 mesh ({V} verts, {Fc} faces)',
            'ms_frame_images': m_img, 'ms_psnr': m_psnr, 'ms_normal_maps': m_maps, 'ms_ssim': m_ssim, 'psnr_db': float(res[1].item()),
            'ssim': float(ss[0].item()),
            'frame_images_gbs': b_img / (m_img * 1e-3) / 1e9, 'psnr_gbs': b_psnr / (m_psnr * 1e-3) / 1e9, 'normal_maps_gbs': b_maps / (m_maps * 1e-3) / 1e9,
            'frac_of_hbm_peak': {'frame_images': b_img / (m_img * 1e-3) / 1e9 / pk['hbm_gbs'], 'psnr': b_psnr / (m_psnr * 1e-3) / 1e9 / pk['hbm_gbs'],
                                 'normal_maps': b_maps / (m_maps * 1e-3) / 1e9 / pk['hbm_gbs']},
            'gpu_launches': 2 + 2 + 3 * 4 + 4, 'steps': steps, 'warmup': warmup,
            'note': 'a few MB per call: launch-latency-bound small kernels, timed through the Python mirror',
            'cpu_port': {'ms_frame_images': 1e3 * t_img, 'ms_psnr': 1e3 * t_psnr, 'ms_normal_maps_extrapolated': 1e3 * t_r * 3 * Fc / max(ns, 1),
                         'kind': 'oracle/images_oracle.py (numpy; python loop over faces)', 'sample': f'{ns} faces of one view timed, scaled to 3 views x {Fc} faces'}}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from arah_release_b200 import ref_layout as rl
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    assert world == args.gpus or world == 1, (world, args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device(f'cuda:{local}')
    n_total = args.warmup + args.steps
    # frames of one sequence, round-robin over ranks: rank r renders frames r, r+N, ...
    frames = [make_frames(args.size, 1, rank + i * world)[0] for i in range(min(n_total, 4))]
    f0 = frames[0]
    dvn, rend, skin, _ = rl.modules_from_frame(f0, dev)
    if world > 1:      # the single collective of the path: frame-invariant weights from rank 0
        for m in (dvn, rend, skin):
            for p_ in m.parameters():
                dist.broadcast(p_.data, src=0)
    net = IDHRNetwork(dvn, rend, skin, BodyRayTracing(n_steps=f0.n_steps, near_surface_vol_samples=f0.near_samples,
                                                      far_surface_vol_samples=f0.far_samples), cano_view_dirs=f0.cano_view_dirs).eval()
    sdfs = [rl.sdf_network_from_frame(f, dev) for f in frames]
    inputs = [rl.inputs_from_frame(f, s, dev) for f, s in zip(frames, sdfs)]
    host = [{k: (v.cpu().pin_memory() if torch.is_tensor(v) and k in ('ray_dirs', 'body_bounds_intersections', 'bone_transforms', 'smpl_verts', 'skinning_weights') else v)
             for k, v in inp.items()} for inp in inputs]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        inp = inputs[i % len(inputs)]
        net(inp)

    def step_host(i):
        hi = host[i % len(host)]
        r = net._renderer(dev, 128, hi['smpl_verts'].shape[1])
        r.set_frame_from_modules(hi['sdf_network'], net.skinning_model, net.rendering_network, net.deviation_network, hi, pose_on_host=True)
        return r.render_host(hi['ray_dirs'][0], hi['body_bounds_intersections'][0])

    rays_step = [inputs[i % len(inputs)]['ray_dirs'].shape[1] for i in range(n_total)]

    # ---------------- device-resident timing (value)
    for i in range(args.warmup):
        step_device(i)
    r = net._last[0]
    r.set_profiling(True)
    barrier()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    shade_ms, shade_flops, shade_flops_ref, step_flops, stats_last = [], [], [], [], None
    corr_ms, corr_flops = [], []
    for k in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        step_device(args.warmup + k)
        ev[k][1].record()
        st = r.stats()                                  # syncs; counters + stage events of this step
        shade_ms.append(st['ms_shade'])
        # executed work of the shading stage: with the exact alpha cull every converged sample gets an SDF-only forward pass and
        # only the survivors (alpha != 0) the full SDF fwd + gradient + colour pass; without it all samples get the full pass
        culled, shaded = st.get('culled_samples', 0), st['shaded_samples']
        cull_ran = r.shade_cull and r.shade_mode == 'tf32'
        shade_flops.append(2.0 * ((shaded * MAC_SDF if cull_ran else 0) + (shaded - culled) * (2 * MAC_SDF + MAC_COL)))
        shade_flops_ref.append(2.0 * shaded * (2 * MAC_SDF + MAC_COL))
        step_flops.append(algorithmic_flops(st))
        corr_ms.append(st['ms_sample_corr'])
        corr_flops.append(2.0 * st['corr_skin_evals'] * MAC_SKIN)
        stats_last = st
        phase_clk = r.phase_clocks()
    barrier()
    clocks = clk.stop() if rank == 0 else None
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3
    rays_timed = sum(rays_step[args.warmup:])

    # ---------------- end-to-end timing through the host-buffer entry point
    r.set_profiling(False)
    for i in range(min(args.warmup, 2)):
        step_host(i)
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall = 0.0
    for k in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev2[k][0].record()
        out = step_host(args.warmup + k)
        ev2[k][1].record()
        torch.cuda.synchronize()
        wall += time.perf_counter() - t0
    barrier()
    t_e2e = max(sum(a.elapsed_time(b) for a, b in ev2) / 1e3, wall)     # host-visible completion dominates
    P0 = rays_step[args.warmup]
    h2d = P0 * 20 + 24 * 16 * 4 + f0.smpl_verts.size * 4 + f0.smpl_weights.size * 4
    d2h = P0 * (12 + 1 + 12)

    # ---------------- reduce over ranks: MAX time, SUM rays
    if world > 1:
        t = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n = torch.tensor([float(rays_timed)], device=dev, dtype=torch.float64)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        t_dev, t_e2e, rays_all = float(t[0]), float(t[1]), float(n[0])
    else:
        rays_all = float(rays_timed)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ours0 = None
    if not args.no_cpu_baseline and args.gpus == 1:
        try:                                                            # the frame the CPU leg samples, rendered once more for the parity entry
            o0 = net(inputs[0])
            torch.cuda.synchronize()
            ours0 = (o0['rgb_values'][0].cpu().numpy(), o0['network_body_mask'][0].cpu().numpy())
        except Exception:
            ours0 = None
    ach = (sum(shade_flops) / max(sum(shade_ms), 1e-9)) / 1e9          # FLOP/ms -> TFLOP/s
    traffic = None
    tp = os.path.join(ROOT, 'profiles', 'k_shade_tc3_traffic.json' if r.shade_mode == 'tf32' else 'k_shade_traffic.json')
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get('dram_bytes_per_launch')
        except Exception:
            traffic = None
    line = {
        'metric': 'rays/sec at 512x512 ZJU-377 render', 'value': rays_all / t_dev, 'unit': 'rays/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_dev / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'ZJU-377-like {args.size}x{args.size} novel-view render, fp32 (BASELINE configs[1])',
                   'rays_per_frame': P0, 'n_steps': f0.n_steps, 'near_far_samples': [f0.near_samples, f0.far_samples],
                   'parallelism': f'frames sharded over {args.gpus} GPU(s), 1 frame/rank/step', 'l2': '256 MB memset between timed steps',
                   'precision': f'fp32 storage/accumulate; shading MLP operands {r.shade_mode}; root finding (sphere tracing, joint search, correspondences) {r.root_mode}',
                   'timing': 'CUDA events on the launching stream, per step, summed; max over ranks'},
        'e2e': {'value': rays_all / t_e2e, 'unit': 'rays/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'api': 'arah_set_frame(pose_on_host) + arah_render_host via IDHRNetwork host wrapper'},
        'gpu_launches': int((stats_last['kernel_launches'] + stats_last['pack_launches']) * args.steps),
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'kernel': f'shading stage: k_shade_tc3<sdf-only> + k_alpha_cull + k_shade_tc3 (SDF fwd + reverse-mode grad + colour MLP; tcgen05 {r.shade_mode} operands, fp32 accumulate in TMEM)'
                     if r.shade_mode == 'tf32' else 'k_shade (fp32 FFMA tiles)',
                     'achieved': ach, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s', 'frac': ach / pk['tf_sustained'],
                     'peak_source': pk['src'] + ' bf16 cuBLAS sustained (MEASURED_PEAKS.json)', 'traffic': traffic,
                     'algorithmic_flops_per_launch': float(np.mean(shade_flops)), 'ms_per_launch': float(np.mean(shade_ms)),
                     'kernel_share_of_step': float(sum(shade_ms) / (1e3 * t_dev)),
                     'exact_alpha_cull': {'enabled': bool(r.shade_cull and r.shade_mode == 'tf32'), 'culled_samples': int(stats_last.get('culled_samples', 0)),
                                          'shaded_samples': int(stats_last['shaded_samples']),
                                          'reference_equivalent_tflops': float(sum(shade_flops_ref) / max(sum(shade_ms), 1e-9) / 1e9),
                                          'note': 'achieved counts EXECUTED flops only (SDF-only pass over all converged samples + full pass over '
                                                  'samples with alpha != 0); reference_equivalent counts the full pass for every sample as the reference executes it'},
                     'whole_step_tflops_reference_equivalent': float(sum(step_flops) / t_dev / 1e12)},
        'roofline_corr': {'bound': 'tensor', 'kernel': f'k_knn_samples + 51 x k_corr_tc3 (per-sample correspondence search; skinning MLP {r.root_mode}: 3 TF32 products count as 1 useful)',
                          'achieved': (sum(corr_flops) / max(sum(corr_ms), 1e-9)) / 1e9, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                          'frac': (sum(corr_flops) / max(sum(corr_ms), 1e-9)) / 1e9 / pk['tf_sustained'], 'ms_per_step': float(np.mean(corr_ms)),
                          'share_of_step': float(sum(corr_ms) / (1e3 * t_dev))},
        'stages_ms_last_step': {k: stats_last[k] for k in ('ms_trace', 'ms_iso', 'ms_sample_corr', 'ms_shade', 'ms_composite', 'ms_total')},
        'phase_cycles_last_step': {'corr': phase_clk[:6], 'shade': phase_clk[8:15], 'trace': phase_clk[16:22],
                                   'note': 'SM cycles of one thread per CTA summed over CTAs/launches: corr = [gather, layer0, mma_wait, epilogue, out_layer, per_point]; '
                                           'shade = [setup+layer0, fwd_wait, fwd_epi, rev_wait, rev_epi, colour_inputs, colour_mlp]'},
        'counters_last_step': {k: stats_last[k] for k in ('rays', 'trace_sdf_evals', 'iso_rays', 'iso_g_evals', 'on_samples', 'corr_skin_evals',
                                                           'shaded_samples', 'hit_rays', 'vol_rays')},
    }
    if args.gpus == 1 and not args.no_train_step:
        try:
            line['train_step'] = train_step_bench(args, dev, f0)
        except Exception as ex:          # secondary metric: never lose the headline line
            line['train_step'] = {'error': repr(ex)[:300]}
    if args.gpus == 1 and not args.no_mesh:
        try:
            step_device(0)                                   # make f0 the current frame of the handle again
            line['mesh_extract'] = mesh_extract_bench(net, f0)
        except Exception as ex:
            line['mesh_extract'] = {'error': repr(ex)[:300]}
    if args.gpus == 1 and not args.no_mesh:
        try:
            line['hypernet'] = hypernet_bench(dev)
        except Exception as ex:
            line['hypernet'] = {'error': repr(ex)[:300]}
    if args.gpus == 1 and not args.no_mesh:
        try:
            line['ray_setup'] = ray_setup_bench(dev, args.size)
        except Exception as ex:
            line['ray_setup'] = {'error': repr(ex)[:300]}
    if args.gpus == 1 and not args.no_mesh:
        try:
            line['image_tail'] = image_tail_bench(net, f0, inputs[0])
        except Exception as ex:
            line['image_tail'] = {'error': repr(ex)[:300]}
    if args.gpus == 1 and not args.no_train_step and isinstance(line.get('train_step'), dict) and 'error' not in line['train_step']:
        try:
            line['train_step'].update(train_step_bench(args, dev, f0, variant='fused'))
        except Exception as ex:
            line['train_step']['fused_loss_error'] = repr(ex)[:300]
    if args.gpus == 1 and not args.no_cpu_baseline:
        v, cores, n, dt, sel, o_cpu = cpu_rate(f0, args.cpu_sample_seconds)
        line['cpu_baseline'] = {'value': v, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
                                'sample': f'{n} random bbox rays of the same frame in {dt:.1f} s (oracle/arah_oracle.c, OpenMP)'}
        if ours0 is not None:
            try:
                line['parity'] = parity_on_sample(ours0[0], ours0[1], sel, o_cpu)
            except Exception as ex:
                line['parity'] = {'error': repr(ex)[:300]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
