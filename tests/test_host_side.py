"""CPU: host logic — the C ABI library loads and exports every symbol include/arah_b200.h declares (no compute without a
GPU), the drop-in modules keep the reference's attribute layout, frame sharding works at world_size 2 (gloo)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def test_cabi_exports_every_declared_symbol():
    from arah_release_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'arah_b200.h')).read()
    declared = set(re.findall(r'\b(arah_[a-z_0-9]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f'{name} declared in include/arah_b200.h but not exported by libarah_b200.so'
    assert declared == set(_lib.EXPORTS)
    assert L.arah_version() >= 100


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA (no silent eager/oracle fallback)."""
    from arah_release_b200 import _lib, synthetic as syn
    from tools import ref_layout as rl
    from arah_release_b200.renderer import ArahRenderer, BodyRayTracing, IDHRNetwork
    with pytest.raises(_lib.ArahError):
        ArahRenderer('cpu')
    fr = syn.make_frame(8, 8, seed=0)
    dev, rend, skin, sdf = rl.modules_from_frame(fr, 'cpu')
    net = IDHRNetwork(dev, rend, skin, BodyRayTracing(), cano_view_dirs=False).eval()
    with pytest.raises(_lib.ArahError):
        net(rl.inputs_from_frame(fr, sdf, 'cpu'))
    with pytest.raises(RuntimeError):
        rend(torch.zeros(1, 3))          # containers cannot compute
    import arah_release_b200.renderer as R
    src = open(R.__file__).read()
    assert 'oracle' not in src.replace('the CPU oracle', '')      # product code never touches oracle/


def test_state_dict_layout_matches_reference_names():
    """Aliased keys of the reference's MetaAvatarRender (SURVEY.md §5 checkpoint row) survive our IDHRNetwork."""
    from arah_release_b200 import synthetic as syn
    from tools import ref_layout as rl
    from arah_release_b200.renderer import BodyRayTracing, IDHRNetwork
    fr = syn.make_frame(8, 8, seed=0)
    dev, rend, skin, _ = rl.modules_from_frame(fr, 'cpu')
    net = IDHRNetwork(dev, rend, skin, BodyRayTracing(n_steps=64), cano_view_dirs=False)
    keys = set(net.state_dict().keys())
    for l in range(6):
        for p in ('weight_g', 'weight_v', 'bias'):
            assert f'rendering_network.lin{l}.{p}' in keys
    for l in range(5):
        for p in ('weight_g', 'weight_v', 'bias'):
            assert f'skinning_model.skinning_decoder_fwd.lin{l}.{p}' in keys
    assert 'deviation_network.variance' in keys
    assert net.ray_tracer.n_steps == 64 and net.ray_tracer.near_surface_vol_samples == 16
    # training mode has no CPU path either: CPU tensors are refused before anything is computed
    from arah_release_b200 import _lib
    sdf = rl.sdf_network_from_frame(fr, 'cpu')
    net.train()
    with pytest.raises(_lib.ArahError):
        net(rl.inputs_from_frame(fr, sdf, 'cpu'))


def test_synthetic_frame_is_deterministic_and_sane():
    from arah_release_b200 import synthetic as syn
    a = syn.make_frame(32, 32, seed=5)
    b = syn.make_frame(32, 32, seed=5)
    np.testing.assert_array_equal(a.ray_dirs, b.ray_dirs)
    np.testing.assert_array_equal(a.bone_transforms, b.bone_transforms)
    assert a.P > 100 and (a.near_far[:, 0] < a.near_far[:, 1]).all()
    np.testing.assert_allclose(np.linalg.norm(a.ray_dirs, axis=1), 1.0, atol=1e-5)
    np.testing.assert_allclose(a.smpl_weights.sum(1), 1.0, atol=1e-5)
    np.testing.assert_allclose(a.bone_transforms[:, 3], np.tile([0, 0, 0, 1], (24, 1)), atol=1e-6)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from arah_release_b200 import sharding as sh
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lin = torch.nn.Linear(4, 4)
    with torch.no_grad():
        lin.weight.fill_(float(rank + 1))
    n = sh.broadcast_module_weights([lin], src=0)
    n_frames = 5
    mine = sh.frames_for_rank(n_frames, rank, world)
    imgs = {fi: torch.full((4, 6, 3), fi, dtype=torch.uint8) for fi in mine}
    res = sh.gather_frames(imgs, n_frames, dst=0)
    ok = bool((lin.weight == 1.0).all()) and n == 20
    if rank == 0:
        ok = ok and all(int(res[i][0, 0, 0]) == i for i in range(n_frames))
    else:
        ok = ok and res is None
    # fewer frames than ranks: the rank that owns nothing still enters the collective (shape from H, W)
    one = {0: torch.full((4, 6, 3), 7, dtype=torch.uint8)} if rank == 0 else {}
    res1 = sh.gather_frames(one, 1, dst=0, H=4, W=6)
    ok = ok and ((rank == 0 and len(res1) == 1 and int(res1[0][0, 0, 0]) == 7) or (rank != 0 and res1 is None))
    q.put((rank, mine, ok))
    dist.destroy_process_group()


def test_frame_sharding_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert out[0][1] == [0, 2, 4] and out[1][1] == [1, 3]
    assert all(o[2] for o in out)


def test_to_image_u8():
    from arah_release_b200 import sharding as sh
    rgb = torch.tensor([[0.0, 0.5, 1.0], [1.2, -0.1, 0.25]])
    img = sh.to_image_u8(rgb, torch.tensor([1, 4]), 2, 3)
    assert img.shape == (2, 3, 3) and img[0, 1].tolist() == [0, 127, 255] and img[1, 1].tolist() == [255, 0, 63]
    assert int(img[0, 0].sum()) == 0


def test_image_tail_has_no_cpu_path_and_validates_arguments():
    """arah_release_b200/images.py: no CPU path; the C ABI rejects bad arguments before touching the device (no GPU needed)."""
    import ctypes as C
    from arah_release_b200 import _lib, images
    with pytest.raises(_lib.ArahError):
        images.FrameImages('cpu')
    src = open(images.__file__).read()
    assert 'oracle' not in src
    L = _lib.lib()
    assert L.arah_frame_images_workspace(512, 512) >= 512 * 512 * 12 and L.arah_frame_images_workspace(0, 5) == 0
    assert L.arah_psnr_workspace() >= 8
    assert L.arah_rasterize_mesh_workspace(100, 64, 64) >= 100 * 12 + 64 * 64 * 8
    assert L.arah_frame_images(None, None, None, 4, 0, 8, None, None, None, 0, None) != 0          # bad image size
    assert b'image size' in L.arah_last_error()
    assert L.arah_frame_images(None, None, None, 100, 4, 4, None, None, None, 0, None) != 0        # P > H*W
    assert L.arah_psnr(None, None, 3, None, None, 0, None) != 0
    assert L.arah_rasterize_mesh(None, 0, None, 0, None, 8, 8, None, None, None, 0, None) != 0
    assert L.arah_face_normal_image(None, 0, None, 0, None, 8, 8, C.c_float(1.0), None, C.c_float(0.0), None, None) != 0


def test_image_tail_cameras_match_oracle():
    """The host-side 3x3 camera algebra of images.py (product) against the oracle's restatement of pytorch3d's conventions."""
    from arah_release_b200 import images
    from oracle import images_oracle as io
    for az in (0.0, 180.0, 37.0):
        R, T = images.look_at_view_transform(2.0, 10.0, az)
        Ro, To = io.look_at_view_transform(2.0, 10.0, az)
        assert np.array_equal(R, Ro) and np.array_equal(T, To)
        c, co = images.fov_perspective_camera(R, T), io.fov_camera(Ro, To)
        assert np.array_equal(np.array(c.R[:], np.float32), co['R'].reshape(9)) and c.fx == float(co['fx']) and c.px == 0.0
    rng = np.random.default_rng(0)
    Rm, Tm = np.linalg.qr(rng.normal(size=(3, 3)))[0].astype(np.float32), rng.normal(size=3).astype(np.float32)
    K = np.array([[500.0, 0, 250.5], [0, 510.0, 260.25], [0, 0, 1]], np.float32)
    c, co = images.opencv_camera(Rm, Tm, K, 480, 512), io.opencv_camera(Rm, Tm, K, 480, 512)
    assert np.array_equal(np.array(c.R[:], np.float32), co['R'].reshape(9)) and np.array_equal(np.array(c.T[:], np.float32), co['T'])
    assert (c.fx, c.fy, c.px, c.py) == (float(co['fx']), float(co['fy']), float(co['px']), float(co['py']))


def test_fused_loss_has_no_cpu_path_and_validates_arguments():
    """arah_release_b200/loss.py: no CPU path; arah_idhr_loss rejects bad arguments before touching the device (no GPU needed)."""
    import ctypes as C
    from arah_release_b200 import _lib, loss
    w = dict(rgb_weight=1.0, perceptual_weight=0.0, eikonal_weight=1.0, mask_weight=0.0, off_surface_weight=1.0, inside_weight=0.0, params_weight=0.0,
             skinning_weight=0.0)
    with pytest.raises(_lib.ArahError):
        loss.IDHRLoss(**w)({'rgb_values': torch.zeros(1, 4, 3)}, {})
    assert 'oracle' not in open(loss.__file__).read()
    L = _lib.lib()
    assert L.arah_idhr_loss_workspace() >= 148 * 7 * 8
    cfg, inp = _lib.ArahLossConfig(), _lib.ArahLossInputs()
    terms = (C.c_float * 9)()
    ws = (C.c_char * int(L.arah_idhr_loss_workspace()))()
    call = lambda: L.arah_idhr_loss(C.byref(cfg), C.byref(inp), C.cast(terms, C.c_void_p), None, C.cast(ws, C.c_void_p), len(ws), None)
    assert call() != 0 and b'n_rays' in L.arah_last_error()                      # no rays
    inp.n_rays, inp.body_mask = 4, C.cast(ws, C.c_void_p)
    cfg.perceptual_weight = 1.0
    assert call() != 0 and b'perceptual' in L.arah_last_error()
    cfg.perceptual_weight, cfg.rgb_weight = 0.0, 1.0
    assert call() != 0 and b'rgb term' in L.arah_last_error()                    # weight on, inputs missing
    cfg.rgb_weight, cfg.rgb_loss_type = 0.0, 7
    assert call() != 0 and b'rgb_loss_type' in L.arah_last_error()


def test_bench_stage_table_rooflines():
    """bench.stage_table: algorithmic / executed FLOPs of the SURVEY 8d formula and the MUFU co-roofline, from device counters alone."""
    import bench
    st = {'rays': 262144, 'trace_sdf_evals': 1766730, 'iso_rays': 49793, 'iso_g_evals': 110738, 'on_samples': 15257069,
          'corr_skin_evals': 74452293, 'shaded_samples': 15247192, 'culled_samples': 12436792,
          'ms_trace': 9.2, 'ms_iso': 4.3, 'ms_sample_corr': 31.9, 'ms_shade': 19.4, 'ms_composite': 0.2, 'ms_total': 65.0}

    class R:
        shade_cull, shade_mode = True, 'tf32'
    out = bench.stage_table([st, st], 2 * 0.066, {'tf_sustained': 1414.5}, R())
    assert set(out) == {'trace', 'iso', 'sample_corr', 'shade'}
    c = out['sample_corr']
    assert abs(c['algorithmic_flops_per_step'] - 2.0 * 74452293 * bench.MAC_SKIN) < 1.0
    assert c['executed_flops_per_step'] < c['algorithmic_flops_per_step']          # the kernel evaluates g(x0) once, the reference twice
    assert abs(c['ms_per_step'] - 31.9) < 1e-9 and 0.0 < c['share_of_step'] < 1.0
    assert abs(c['frac'] - c['achieved'] / 1414.5) < 1e-12
    # 1024 exp + log per executed skinning evaluation against 16 / clk / SM
    assert abs(c['mufu_ops_per_step'] - 1024.0 * (74452293 - 15257069)) < 1.0
    assert 0.3 < c['mufu_frac_of_16_per_clk_per_sm'] < 0.6
    s = out['shade']
    assert s['executed_flops_per_step'] < s['algorithmic_flops_per_step']          # the exact cull removes work the reference does
    assert s['frac'] > s['executed_frac'] > 0.0
