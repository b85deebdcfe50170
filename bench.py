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
    ap.add_argument('--seq-frames', type=int, default=258, help='frames of the BASELINE configs[3] sequence entry (0 = skip)')
    ap.add_argument('--seq-lattice', type=int, default=256, help='lattice side of the canonical mesh extracted per sequence frame (0 = no normal maps)')
    ap.add_argument('--no-h36m', action='store_true', help='skip the BASELINE configs[4] entry (1024x1024, 128 near samples)')
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


def host_threads():
    """Every host core this process may run on: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which must not
    shrink the CPU arm (the reference's CPU path uses all cores)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def workload_config(size, frame, gpus, ours=True, sample=None):
    """`config` of the JSON line: the SAME keys in both arms."""
    return {'workload': f'ZJU-377-like {size}x{size} novel-view render, fp32 storage (BASELINE configs[1])', 'rays_per_frame': int(frame.P),
            'n_steps': int(frame.n_steps), 'near_far_samples': [int(frame.near_samples), int(frame.far_samples)],
            'parallelism': f'frames sharded over {gpus} GPU(s), 1 frame/rank/step' if ours else 'rank 0 only, all host cores (OpenMP over rays)',
            'l2': '256 MB memset between timed steps' if ours else 'n/a (CPU arm)',
            'timing': 'CUDA events on the launching stream, per step, summed; max over ranks' if ours else 'perf_counter around each step',
            'sample': sample or 'all bbox rays of the frame, every step'}


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
    nthr = host_threads() if threads <= 0 else threads
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
    nthr = host_threads()
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
            'config': workload_config(args.size, frame, args.gpus, ours=False,
                                      sample=f'{n} random bbox rays of the frame per step (value = rays timed / seconds: the reference cost is linear in P)'),
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
    from arah_release_b200 import synthetic as syn
    from tools import ref_layout as rl
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
    tracer_ms = None
    try:                                   # where the forward's tracer part goes (one extra, untimed step with the library's stage events)
        r = net._last[0]
        r.set_profiling(True)
        net(inp)
        stp = r.stats()
        tracer_ms = {k: stp[k] for k in ('ms_trace', 'ms_iso', 'ms_sample_corr')}
        r.set_profiling(False)
    except Exception as ex:
        tracer_ms = {'error': repr(ex)[:200]}
    # algorithmic MACs of the differentiable part per sample: SDF fwd + input-gradient sweep + their second- and first-order
    # backward (4 data GEMM sweeps + 2 weight-gradient sweeps), colour net fwd + data + weight gradient, skinning net value +
    # 3 tangents + backward
    mac = samples * (6 * MAC_SDF + 3 * (MAC_COL + 128 * 256) + 6 * MAC_SKIN)
    return {'workload': f'{P} rays of the ZJU-377-like frame, train_skinning_net, fp32 (BASELINE configs[2] shape)', 'rays': int(P),
            'shaded_samples': int(samples), 'ms_per_step': tot, 'ms_forward_incl_tracer': float(np.mean(ms['forward'])),
            'ms_backward': float(np.mean(ms['backward'])), 'rays_per_s': P / tot * 1e3, 'steps': steps, 'warmup': warmup,
            'gpu_launches_per_step': int(launches), 'loss': float(loss.detach()), 'tracer_stages_ms': tracer_ms,
            'algorithmic_tflops_differentiable_part': 2.0 * mac / (tot * 1e-3) / 1e12,
            'frac_of_bf16_peak': 2.0 * mac / (tot * 1e-3) / 1e12 / peaks()['tf_sustained'],
            'note': 'tracer (persistent kernels) + hand-written forward/backward GEMM chains; latency-bound at 2048 rays'}


# ------------------------------------------------------------------------------------------------ canonical mesh (row f1)
def mesh_extract_bench(net, frame, steps=3, warmup=2, N=256):
    """MetaAvatarRender.forward(gen_cano_mesh=True) front half (models/__init__.py:203-206): 256^3 SDF lattice + iso-surface.
    SDF lattice: tensor roofline (N^3 x 328 704 MAC, 3xTF32 counts as 1x useful); marching cubes: HBM roofline, algorithmic
    bytes = one read of the lattice (4 N^3) + the vertices and faces written."""
    import torch
    from oracle import oracle as orc
    r = net._last[0]
    banded = r.root_mode == '3xtf32'
    ms_g, ms_m, ms_full = [], [], []
    stats = None
    for i in range(warmup + steps):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        # what the mesh branch runs: the banded lattice (fp16 pass + split precision near the surface) where the fp16 images exist
        if banded:
            vol, stats = r.sdf_grid_banded(N)
        else:
            vol = r.sdf_grid(N)
        e[1].record()
        v, f = r.marching_cubes(vol, max_verts=1 << 20, max_faces=1 << 21)
        e[2].record()
        full = r.sdf_grid(N)                 # every point in split precision (round 1's lattice), for comparison
        e[3].record()
        torch.cuda.synchronize()
        if i >= warmup:
            ms_g.append(e[0].elapsed_time(e[1])); ms_m.append(e[1].elapsed_time(e[2])); ms_full.append(e[2].elapsed_time(e[3]))
    del full
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
            'gpu_launches': 8 if banded else 5, 'steps': steps, 'warmup': warmup,
            'lattice': ('banded: fp16 pass over all points, split precision for the corners of cells within eps of the level; marching-cubes '
                        'output bit-identical to the full lattice (tests/test_gpu_mesh.py)') if banded else 'full precision',
            'banded_stats': [int(x) for x in stats.tolist()] if stats is not None else None,
            'ms_sdf_grid_full_split_precision': float(np.mean(ms_full)),
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



# ------------------------------------------------------------------------------------------------ per-stage rooflines
def stage_table(stats, t_dev, pk, r):
    """Algorithmic and executed FLOPs of every stage of the timed steps against the measured bf16 peak."""
    n = len(stats)
    S = lambda k: float(sum(st[k] for st in stats))
    cull_ran = bool(r.shade_cull and r.shade_mode == 'tf32')
    shaded, culled = S('shaded_samples'), (S('culled_samples') if cull_ran else 0.0)
    rows = {
        'trace': ('k_trace_persist (1-NN + inverse NN skinning + SDF per marching step)', 'ms_trace', 2.0 * S('trace_sdf_evals') * MAC_SDF, None),
        'iso': ('k_iso_init_tc3 + k_iso_persist (joint search: skinning MLP + SDF per Broyden step)', 'ms_iso',
                2.0 * (S('iso_rays') * (4 * MAC_SKIN + 2 * MAC_SDF) + S('iso_g_evals') * (MAC_SKIN + MAC_SDF)), None),
        # the reference evaluates g twice at the start of every search (J init + g(x0)); the kernel evaluates it once
        'sample_corr': ('k_knn_samples + k_corr_persist (per-sample correspondence search, skinning MLP in 3xfp16 split precision)', 'ms_sample_corr',
                        2.0 * S('corr_skin_evals') * MAC_SKIN, 2.0 * (S('corr_skin_evals') - S('on_samples')) * MAC_SKIN),
        'shade': ('k_sdf_fwd16 + k_alpha_cull + k_shade16 (SDF value of all samples; gradient + colour MLP of the alpha != 0 samples; fp16 operands)',
                  'ms_shade', 2.0 * shaded * (2 * MAC_SDF + MAC_COL),
                  2.0 * ((shaded * MAC_SDF if cull_ran else 0.0) + (shaded - culled) * (2 * MAC_SDF + MAC_COL))),
    }
    main = {'trace': 'k_trace_persist', 'iso': 'k_iso_persist', 'sample_corr': 'k_corr_persist', 'shade': 'k_shade16'}
    # the co-bounding MUFU pipe: EXECUTED transcendental operations of the MLP activations against 16 / clk / SM at the
    # maximum SM clock — sines of the SDF (1536 per evaluation; + 1536 cosines where the gradient is taken), exp + log of the
    # skinning MLP's softplus (1024 per evaluation)
    mufu = {'trace': 1536.0 * S('trace_sdf_evals'),
            'iso': (1536.0 + 1024.0) * S('iso_g_evals') + S('iso_rays') * 4 * (1536.0 + 2 * 1024.0),
            'sample_corr': 1024.0 * (S('corr_skin_evals') - S('on_samples')),
            'shade': 1536.0 * shaded + 2 * 1536.0 * (shaded - culled)}
    try:
        import torch
        mufu_peak = torch.cuda.get_device_properties(0).multi_processor_count * 16 * 1.965e9
    except Exception:
        mufu_peak = 148 * 16 * 1.965e9
    out = {}
    for k, (name, msk, alg, ex) in rows.items():
        ms = S(msk)
        ex = alg if ex is None else ex
        out[k] = {'kernel': name, 'main_kernel': main[k], 'ms_per_step': ms / n, 'share_of_step': ms / (1e3 * t_dev),
                  'algorithmic_flops_per_step': alg / n, 'executed_flops_per_step': ex / n,
                  'achieved': alg / max(ms, 1e-9) / 1e9, 'frac': alg / max(ms, 1e-9) / 1e9 / pk['tf_sustained'],
                  'executed_tflops': ex / max(ms, 1e-9) / 1e9, 'executed_frac': ex / max(ms, 1e-9) / 1e9 / pk['tf_sustained'],
                  'mufu_ops_per_step': mufu[k] / n, 'mufu_frac_of_16_per_clk_per_sm': mufu[k] / max(ms * 1e-3, 1e-12) / mufu_peak}
    return out


# ------------------------------------------------------------------------------------------------ configs[3]: sharded sequence
def sequence_bench(args, dev, net, world, rank):
    """BASELINE configs[3] (test.py:71-80, lightning_model.py:306-360): a novel-pose sequence, frame i -> rank i mod N, per frame the
    render, the canonical mesh + three normal maps (gen_cano_mesh=True) and the uint8 image; one gather of the images at the end.
    Device time per frame from CUDA events (frame preparation on the host is the dataset's job and stays outside), max over ranks."""
    import torch
    import torch.distributed as dist
    from arah_release_b200 import sharding as sh, synthetic as syn
    from tools import ref_layout as rl
    mine = sh.frames_for_rank(args.seq_frames, rank, world)
    t_render = t_mesh = 0.0
    rays = 0
    images = {}
    t_wall = time.perf_counter()
    H = W = args.size
    for fi in mine:
        f = syn.make_frame(args.size, args.size, seed=0, frame_idx=fi)
        inp = rl.inputs_from_frame(f, rl.sdf_network_from_frame(f, dev), dev)
        pix = torch.from_numpy(f.pix).to(dev)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out = net(inp)
        images[fi] = sh.to_image_u8(out['rgb_values'][0], pix, f.H, f.W)
        e[1].record()
        if args.seq_lattice > 0:
            inp2 = dict(inp, cam_rot=torch.from_numpy(f.pose[:3, :3].copy()).view(1, 3, 3), cam_trans=torch.from_numpy(f.pose[:3, 3].copy()).view(1, 3),
                        intrinsics=torch.from_numpy(f.K).view(1, 3, 3))
            net.render_normal_maps(inp2, N=args.seq_lattice, image_size=(f.H, f.W))
        e[2].record()
        torch.cuda.synchronize()
        t_render += e[0].elapsed_time(e[1]) / 1e3
        t_mesh += e[1].elapsed_time(e[2]) / 1e3
        rays += f.P
    t_wall = time.perf_counter() - t_wall
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    g0 = time.perf_counter()
    gathered = sh.gather_frames(images, args.seq_frames, H=H, W=W, device=dev)
    torch.cuda.synchronize()
    t_gather = time.perf_counter() - g0
    per_rank = [[t_render, t_mesh, t_wall, float(rays), t_gather]]
    if world > 1:
        tl = torch.tensor(per_rank[0], device=dev, dtype=torch.float64)
        allt = [torch.zeros_like(tl) for _ in range(world)]
        dist.all_gather(allt, tl)
        per_rank = [[float(v) for v in x] for x in allt]
    if rank != 0:
        return None
    tr, tm = max(x[0] for x in per_rank), max(x[0] + x[1] for x in per_rank)
    rays_all = sum(x[3] for x in per_rank)
    return {'workload': f'{args.seq_frames}-frame synthetic novel-pose sequence at {args.size}x{args.size} (BASELINE configs[3]), frame i -> rank i mod {world}',
            'frames': args.seq_frames, 'n_gpus': world, 'rays': int(rays_all),
            'seconds_render_only': tr, 'rays_per_s_render_only': rays_all / tr, 'frames_per_s_render_only': args.seq_frames / tr,
            'seconds_with_normal_maps': tm, 'frames_per_s_with_normal_maps': args.seq_frames / tm, 'normal_maps_lattice': args.seq_lattice,
            'mesh_branch_share': (tm - tr) / tm if tm > 0 else None,
            'seconds_gather_uint8': max(x[4] for x in per_rank), 'gathered_images': None if gathered is None else len(gathered),
            'seconds_wall_incl_host_frame_prep': max(x[2] for x in per_rank),
            'per_rank_seconds_device': [x[0] + x[1] for x in per_rank],
            'timing': 'CUDA events per frame around render (+ uint8 image) and around mesh extraction + 3 normal maps, summed per rank, max over ranks'}


# ------------------------------------------------------------------------------------------------ configs[4]: 1024^2, 128 near samples
def strict_fp32_bench(dev, frame, _unused=None, steps=2, warmup=1):
    """The same 512x512 frame with shade_mode = root_mode = 'fp32': every MLP on fp32 FFMA tiles in the oracle's arithmetic order,
    one launch per iteration (the mode the tensor-core path is checked against; VERDICT r1 asked for a driver-timed number)."""
    import torch
    from tools import ref_layout as rl
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    dvn, rend, skin, sdf = rl.modules_from_frame(frame, dev)
    net = IDHRNetwork(dvn, rend, skin, BodyRayTracing(n_steps=frame.n_steps, near_surface_vol_samples=frame.near_samples,
                                                      far_surface_vol_samples=frame.far_samples), cano_view_dirs=frame.cano_view_dirs,
                      shade_mode='fp32', root_mode='fp32').eval()
    inp = rl.inputs_from_frame(frame, sdf, dev)
    ms = []
    for i in range(warmup + steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net(inp); e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ms.append(e0.elapsed_time(e1))
    st = net.stats()
    m = float(np.mean(ms))
    return {'workload': '512x512 frame, shade_mode fp32 + root_mode fp32 (FFMA tiles, one launch per iteration)', 'rays': int(frame.P), 'ms_per_frame': m,
            'rays_per_s': frame.P / m * 1e3, 'steps': steps, 'warmup': warmup, 'gpu_launches_per_frame': int(st['kernel_launches'])}


def h36m_bench(dev, pk, steps=3, warmup=2):
    """BASELINE configs[4] (im2mesh/config.py:225, ray_tracing.py:336,346): 1024x1024, n_steps 160, 128 near-surface samples, canonical view
    directions: ~10^6 rays x 160 sample slots, the HBM / MLP stress case."""
    import torch
    from arah_release_b200 import synthetic as syn
    from tools import ref_layout as rl
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    fr = syn.make_frame(1024, 1024, seed=4, n_steps=160, near_samples=128, far_samples=16, cano_view_dirs=True, beta=0.003)
    dvn, rend, skin, sdf = rl.modules_from_frame(fr, dev)
    net = IDHRNetwork(dvn, rend, skin, BodyRayTracing(n_steps=fr.n_steps, near_surface_vol_samples=fr.near_samples, far_surface_vol_samples=fr.far_samples),
                      cano_view_dirs=True).eval()
    inp = rl.inputs_from_frame(fr, sdf, dev)
    free0 = torch.cuda.mem_get_info(dev)[0]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        net(inp)
    rr = net._last[0]
    rr.set_profiling(True)
    ms, sts = [], []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net(inp); e1.record()
        sts.append(rr.stats())
        ms.append(e0.elapsed_time(e1))
    free1 = torch.cuda.mem_get_info(dev)[0]
    m = float(np.mean(ms))
    stt = stage_table(sts, sum(ms) / 1e3, pk, rr)
    st = sts[-1]
    return {'workload': '1024x1024, n_steps 160, 128 near + 16 far samples, cano_view_dirs (BASELINE configs[4])', 'rays': int(fr.P),
            'ms_per_frame': m, 'rays_per_s': fr.P / m * 1e3, 'steps': steps, 'warmup': warmup,
            'stages_ms': {k: st[k] for k in ('ms_trace', 'ms_iso', 'ms_sample_corr', 'ms_shade', 'ms_composite')},
            'counters': {k: st[k] for k in ('on_samples', 'corr_skin_evals', 'shaded_samples', 'culled_samples', 'hit_rays')},
            'roofline_stages': {k: {'ms_per_step': v['ms_per_step'], 'frac': v['frac'], 'executed_frac': v['executed_frac']} for k, v in stt.items()},
            'device_memory_in_use_gb': (free0 - free1) / 2 ** 30 + 0.25, 'gpu_launches_per_frame': int(st['kernel_launches'])}


# ------------------------------------------------------------------------------------------------ configs[2]: CPU leg of the training step
def train_step_cpu(frame, rays=192, seed=0):
    """One training step of the CPU port (oracle/arah_oracle.c training-mode tracer on all host threads + oracle/train_oracle.py, the torch
    fp32 autograd restatement pinned to the unmodified reference by tests/golden/train_*.npz) on a bounded ray sample."""
    import copy
    import torch
    from arah_release_b200 import synthetic as syn
    from oracle import oracle as orc, train_oracle as to
    torch.set_num_threads(host_threads())
    rng = np.random.default_rng(1)
    sel = np.sort(rng.choice(frame.P, size=min(rays, frame.P), replace=False))
    fr = copy.copy(frame)
    fr.ray_dirs, fr.near_far, fr.pix = frame.ray_dirs[sel], frame.near_far[sel], frame.pix[sel]
    aux = syn.train_aux_points(fr, seed=seed)
    t = time.perf_counter()
    c = orc.render(fr, train_noise=orc.train_noise(fr, seed), threads=host_threads())
    trace = {k[len('trace.'):]: v for k, v in c.items() if k.startswith('trace.')}
    to.train_step(fr, aux, trace, seed, train_skinning_net=True)
    dt = time.perf_counter() - t
    return {'value': fr.P / dt, 'unit': 'rays/s', 'cores': host_threads(), 'kind': 'port',
            'sample': f'{fr.P} rays of the frame, one forward + backward in {dt:.1f} s (C tracer + torch autograd restatement)'}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from tools import ref_layout as rl
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
    step_flops, stats_last, stage_acc = [], None, []
    for k in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations (outside the event pair)
        ev[k][0].record()
        step_device(args.warmup + k)
        ev[k][1].record()
        st = r.stats()                                  # syncs; counters + stage events of this step
        step_flops.append(algorithmic_flops(st))
        stage_acc.append(st)
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

    # ---------------- reduce over ranks: MAX time, SUM rays; per-rank spread of the device time
    per_rank_ms = [1e3 * t_dev / args.steps]
    if world > 1:
        tl = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
        allt = [torch.zeros_like(tl) for _ in range(world)]
        dist.all_gather(allt, tl)
        per_rank_ms = [1e3 * float(x[0]) / args.steps for x in allt]
        t_dev, t_e2e = max(float(x[0]) for x in allt), max(float(x[1]) for x in allt)
        n = torch.tensor([float(rays_timed)], device=dev, dtype=torch.float64)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        rays_all = float(n[0])
    else:
        rays_all = float(rays_timed)
    # ---------------- BASELINE configs[3]: the 258-frame sequence sharded over the ranks (every rank takes part)
    seq = None
    if args.seq_frames > 0:
        try:
            seq = sequence_bench(args, dev, net, world, rank)
        except Exception as ex:
            seq = {'error': repr(ex)[:300]}
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
    # ---------------- rooflines: every stage against the measured bf16 peak; the stage with the largest share is `roofline`
    stt = stage_table(stage_acc, t_dev, pk, r)
    dom = max(stt, key=lambda k: stt[k]['ms_per_step'])
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get(stt[dom]['main_kernel'], {}).get('dram_bytes_per_launch')
            traffic_src = tj.get('_source')
        except Exception:
            traffic = None
    roof = dict(stt[dom])
    roof.update({'bound': 'tensor', 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s', 'peak_source': pk['src'] + ' bf16 cuBLAS sustained (MEASURED_PEAKS.json)',
                 'stage': dom, 'traffic': traffic, 'traffic_source': traffic_src,
                 'note': 'achieved = ALGORITHMIC flops (SURVEY §8d formula through the device counters; a split-precision product counts once) / '
                         'stage time from CUDA events recorded inside the library at stage boundaries'})
    exec_flops = sum(v['executed_flops_per_step'] for v in stt.values())
    line = {
        'metric': 'rays/sec at 512x512 ZJU-377 render', 'value': rays_all / t_dev, 'unit': 'rays/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_dev / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'fp32 storage/accumulate; tensor-core operands: fp16 (SDF value, gradient, colour), 3xfp16 split ~ fp32 (all root finding)'
                 if r.shade_mode == 'tf32' else 'f32', 'data': 'synthetic',
        'config': workload_config(args.size, f0, args.gpus),
        'e2e': {'value': rays_all / t_e2e, 'unit': 'rays/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'api': 'arah_set_frame(pose_on_host) + arah_render_host via IDHRNetwork host wrapper'},
        'gpu_launches': int((stats_last['kernel_launches'] + stats_last['pack_launches']) * args.steps),
        'gpu_launches_per_frame': {'render': int(stats_last['kernel_launches']), 'set_frame_packing': int(stats_last['pack_launches'])},
        'clocks': clocks,
        'roofline': roof,
        'roofline_stages': stt,
        'whole_step_executed_frac': exec_flops / (1e3 * t_dev / args.steps * 1e-3) / 1e12 / pk['tf_sustained'],
        'whole_step_tflops_reference_equivalent': float(sum(step_flops) / t_dev / 1e12),
        'per_rank_ms_per_step': {'min': min(per_rank_ms), 'max': max(per_rank_ms), 'mean': float(np.mean(per_rank_ms)), 'all': per_rank_ms},
        'stages_ms_last_step': {k: stats_last[k] for k in ('ms_trace', 'ms_iso', 'ms_sample_corr', 'ms_shade', 'ms_composite', 'ms_total')},
        'phase_cycles_last_step': {'corr': phase_clk[:6], 'shade': phase_clk[8:15], 'trace': phase_clk[16:22],
                                   'note': 'SM cycles of one thread per CTA summed over CTAs/launches: corr = [-, layer0, mma_wait, epilogue, out_layer, per_point]; '
                                           'shade = [setup+layer0, fwd_wait, fwd_epi, rev_wait, rev_epi, colour_inputs, colour_mlp]; '
                                           'trace = [1-NN, layer0, mma_wait, epilogue, marching, evaluations]'},
        'counters_last_step': {k: stats_last[k] for k in ('rays', 'trace_sdf_evals', 'iso_rays', 'iso_g_evals', 'on_samples', 'corr_skin_evals',
                                                           'shaded_samples', 'culled_samples', 'hit_rays', 'vol_rays')},
    }
    if seq is not None:
        line['sequence258'] = seq
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
    if args.gpus == 1 and not args.no_h36m:
        try:
            line['strict_fp32'] = strict_fp32_bench(dev, f0, None)
        except Exception as ex:
            line['strict_fp32'] = {'error': repr(ex)[:300]}
    if args.gpus == 1 and not args.no_h36m:
        try:
            del net
            torch.cuda.empty_cache()
            line['h36m_1024'] = h36m_bench(dev, pk)
        except Exception as ex:
            line['h36m_1024'] = {'error': repr(ex)[:300]}
    if args.gpus == 1 and not args.no_cpu_baseline:
        v, cores, n, dt, sel, o_cpu = cpu_rate(f0, args.cpu_sample_seconds)
        line['cpu_baseline'] = {'value': v, 'unit': 'rays/s', 'cores': cores, 'kind': 'port',
                                'sample': f'{n} random bbox rays of the same frame in {dt:.1f} s (oracle/arah_oracle.c, OpenMP)'}
        if ours0 is not None:
            try:
                line['parity'] = parity_on_sample(ours0[0], ours0[1], sel, o_cpu)
            except Exception as ex:
                line['parity'] = {'error': repr(ex)[:300]}
        if not args.no_train_step and isinstance(line.get('train_step'), dict) and 'error' not in line['train_step']:
            try:
                line['train_step']['cpu_baseline'] = train_step_cpu(f0)
                line['train_step']['speedup_vs_cpu_port'] = line['train_step']['rays_per_s'] / line['train_step']['cpu_baseline']['value']
            except Exception as ex:
                line['train_step']['cpu_baseline'] = {'error': repr(ex)[:300]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
