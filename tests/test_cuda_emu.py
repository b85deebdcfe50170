"""CPU: the SOURCE of csrc/arah_image.cu and csrc/arah_loss.cu — kernels, launch geometry and C-ABI entry points, unchanged —
executed on a CPU model of CUDA (tests/native/cuda_emu.h: one OS thread per CUDA thread, barriers for __syncthreads and warp
shuffles / ballots, locked atomics) and driven by the PRODUCT's Python mirrors through the product's ctypes signatures.

The build container has no GPU; this is the closest available check of what nvcc compiles for the GPU box: indexing, grid sizes,
reduction trees, the warp-cooperative rasteriser path, argument marshalling.  It says nothing about performance or the memory model;
the GPU tests (test_gpu_zx_loss.py, test_gpu_zz_images.py) remain the parity tests proper.  Test infrastructure only.
Bars as on the GPU: images / pix_to_face / depth bit-exact against the oracle, goldens of the unmodified reference within their
stated tolerances."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NATIVE = os.path.join(HERE, 'native')
SO = os.path.join(NATIVE, 'libarah_emu.so')
SRCS = [os.path.join(NATIVE, f) for f in ('emu_image.cpp', 'emu_loss.cpp', 'emu_support.cpp')]
DEPS = SRCS + [os.path.join(NATIVE, 'cuda_emu.h'), os.path.join(ROOT, 'include', 'arah_b200.h')] + \
    [os.path.join(ROOT, 'arah_release_b200', 'csrc', f) for f in ('arah_image.cu', 'arah_image_core.h', 'arah_loss.cu', 'arah_loss_core.h')]


@pytest.fixture(scope='module')
def emu_lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in DEPS):
        cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
        subprocess.check_call([cxx, '-std=c++20', '-O1', '-pthread', '-shared', '-fPIC', '-ffp-contract=off', '-o', SO] + SRCS)
    from arah_release_b200 import _lib
    L = C.CDLL(SO)
    L.arah_last_error.restype = C.c_char_p
    _lib.declare_image_and_loss(L)
    return L


@pytest.fixture()
def on_emulator(emu_lib, monkeypatch):
    """Point the product's mirrors at the emulated library and lift the two things that need a GPU (device check, stream)."""
    from arah_release_b200 import _lib, images, loss
    monkeypatch.setattr(_lib, 'lib', lambda: emu_lib)
    monkeypatch.setattr(loss.IDHRLoss, '_require_cuda', staticmethod(lambda dev: None))
    monkeypatch.setattr(loss.IDHRLoss, '_stream', lambda self, dev: None)
    monkeypatch.setattr(torch.cuda, 'synchronize', lambda *a, **k: None)

    class HostFrameImages(images.FrameImages):
        def __init__(self):
            self.device = torch.device('cpu')
            self._ws = {}

        @property
        def _stream(self):
            return None
    return HostFrameImages


def _gpu_image_tests(on_emulator, monkeypatch):
    """tests/test_gpu_zz_images.py with its device set to 'cpu' and its FrameImages factory bound to the emulated library."""
    import test_gpu_zz_images as G
    fi = on_emulator()
    monkeypatch.setattr(G, 'DEV', 'cpu')
    monkeypatch.setattr(G, '_fi', lambda: fi)
    return G


def test_emulator_runs_block_reductions_and_ballots(emu_lib):
    """Sanity of the execution model itself on the PSNR kernels: a two-stage block reduction over 9 CTAs, against numpy."""
    rng = np.random.default_rng(0)
    a, b = rng.random(17001).astype(np.float32), rng.random(17001).astype(np.float32)
    out = np.zeros(2)
    ws = np.zeros(int(emu_lib.arah_psnr_workspace()), np.uint8)
    rc = emu_lib.arah_psnr(a.ctypes.data, b.ctypes.data, a.size, out.ctypes.data, ws.ctypes.data, ws.size, None)
    assert rc == 0
    d = (a - b).astype(np.float32)
    mse = float(np.mean((d * d).astype(np.float64)))
    assert abs(out[0] - mse) <= 1e-7 * mse and abs(out[1] + 10 * np.log10(mse)) <= 1e-5


@pytest.mark.parametrize('seed', [0, 1])
def test_image_tail_kernels_on_emulator(on_emulator, monkeypatch, seed):
    G = _gpu_image_tests(on_emulator, monkeypatch)
    G.test_frame_images_match_reference_validation_step(seed)
    if seed == 0:
        G.test_frame_images_edge_cases()
        G.test_rasterize_degenerate_inputs()


def test_ssim_kernels_on_emulator(on_emulator, monkeypatch):
    G = _gpu_image_tests(on_emulator, monkeypatch)
    G.test_ssim_matches_oracle(light=True)
    G.test_validation_tail_on_golden_batch(seed=0)


@pytest.mark.parametrize('mesh,H,W', [('hand', 48, 64), ('torus', 40, 40)])
def test_rasteriser_kernels_on_emulator(on_emulator, monkeypatch, mesh, H, W):
    """'hand': large triangles -> the warp-cooperative sweep (ballot path); 'torus': small faces -> one lane per face."""
    _gpu_image_tests(on_emulator, monkeypatch).test_rasterize_bit_exact_against_oracle(mesh, H, W)


@pytest.mark.parametrize('seed', [1, 3])          # ZJU-313 weights / patch labels / 2100 rays; every term on + degenerate inputs
def test_loss_kernels_on_emulator(on_emulator, monkeypatch, seed):
    import test_gpu_zx_loss as G
    monkeypatch.setattr(G, 'DEV', 'cpu')
    G.test_fused_loss_matches_reference(seed)


def test_bench_image_tail_entry_runs_on_emulator(on_emulator, monkeypatch):
    """bench.py::image_tail_bench end to end (its Python, the event bookkeeping, the byte accounting, the CPU-port leg) with a
    stand-in renderer and the emulated library: a typo there would otherwise only show on the GPU box."""
    import types
    import bench
    from arah_release_b200 import images
    from helpers_images import iso_mesh, make_camera
    H = W = 48
    rng = np.random.default_rng(0)
    mask = rng.random((H, W)) < 0.4
    pix = np.flatnonzero(mask.reshape(-1))
    P = len(pix)
    v, f = iso_mesh('torus', 20)
    R, T, K = make_camera(H, W)
    T = T + np.array([0, 0, 2.6], np.float32)
    pose = np.eye(4, dtype=np.float32); pose[:3, :3] = R; pose[:3, 3] = T
    frame = types.SimpleNamespace(H=H, W=W, pix=pix.astype(np.int64), pose=pose, K=K)
    out = {'rgb_values': torch.from_numpy(rng.random((1, P, 3)).astype(np.float32)), 'points_cam': torch.from_numpy(rng.random((1, P, 3)).astype(np.float32) + 2)}

    class Net:
        def __call__(self, inp):
            return out

        def extract_canonical_mesh(self, inp, N=256):
            return torch.from_numpy(v), torch.from_numpy(f), torch.from_numpy(v + np.float32(0.01))

    class Event:
        def __init__(self, enable_timing=True):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 1.0
    monkeypatch.setattr(torch.cuda, 'Event', Event)
    monkeypatch.setattr(images, 'FrameImages', lambda dev: on_emulator())
    res = bench.image_tail_bench(Net(), frame, {}, steps=1, warmup=0, N=20)
    assert res['ms_frame_images'] == 1.0 and res['ms_ssim'] == 1.0 and 0.0 < res['ssim'] <= 1.0 and res['psnr_db'] > 0
    assert set(res['frac_of_hbm_peak']) == {'frame_images', 'psnr', 'normal_maps'} and res['cpu_port']['ms_frame_images'] > 0


def test_render_normal_maps_of_the_drop_in_module_on_emulator(on_emulator, monkeypatch):
    """`IDHRNetwork.render_normal_maps` (the `gen_cano_mesh` tail, models/__init__.py:226-309) with a given mesh: reads the reference's
    camera keys, returns the reference's three output keys; values against the oracle."""
    from arah_release_b200 import images, synthetic as syn
    from tools import ref_layout as rl
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    from helpers_images import iso_mesh, make_camera
    from oracle import images_oracle as io
    monkeypatch.setattr(images, 'FrameImages', lambda dev: on_emulator())
    fr = syn.make_frame(8, 8, seed=0)
    dev, rend, skin, _ = rl.modules_from_frame(fr, 'cpu')
    net = IDHRNetwork(dev, rend, skin, BodyRayTracing(n_steps=64), cano_view_dirs=False).eval()
    v, f = iso_mesh('two_spheres', 20)
    posed = (v + np.float32(0.02)).astype(np.float32)
    H, W = 40, 56
    R, T, K = make_camera(H, W)
    T = T + np.array([0, 0, 2.6], np.float32)
    inp = {'cam_rot': torch.from_numpy(R).view(1, 3, 3), 'cam_trans': torch.from_numpy(T).view(1, 3), 'intrinsics': torch.from_numpy(K).view(1, 3, 3)}
    maps = net.render_normal_maps(inp, image_size=(H, W), mesh=(torch.from_numpy(v), torch.from_numpy(f), torch.from_numpy(posed)))
    ref = io.normal_maps(v, f, posed, R, T, K, H, W)
    assert set(maps) == set(ref) == {'output_normal', 'normal_cano_front', 'normal_cano_back'}
    for k in ref:
        assert maps[k].shape == (1, H, W, 3) and np.array_equal(maps[k][0].numpy(), ref[k])
