"""ctypes binding of libarah_b200.so (include/arah_b200.h).  No fallback: a missing library is an error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, 'libarah_b200.so')
FP = C.c_void_p


class ArahConfig(C.Structure):
    _fields_ = [('device', C.c_int32), ('n_steps', C.c_int32), ('near_samples', C.c_int32), ('far_samples', C.c_int32),
                ('cano_view_dirs', C.c_int32), ('latent_dim', C.c_int32), ('n_verts', C.c_int32), ('max_rays', C.c_int32),
                ('shade_mode', C.c_int32), ('root_mode', C.c_int32), ('shade_cull', C.c_int32), ('render_last_pt', C.c_int32)]


class ArahFrame(C.Structure):
    _fields_ = [('sdf_W', FP * 7), ('sdf_b', FP * 7), ('sdf_freq', FP), ('sdf_phase', FP),
                ('skin_W', FP * 5), ('skin_b', FP * 5), ('col_W', FP * 6), ('col_b', FP * 6),
                ('latent', FP), ('beta', C.c_float),
                ('bone_transforms', FP), ('smpl_verts', FP), ('smpl_weights', FP), ('pose_on_host', C.c_int32),
                ('trans', C.c_float * 3), ('coord_min', C.c_float), ('coord_max', C.c_float), ('center', C.c_float * 3),
                ('cam_loc', C.c_float * 3), ('pose', C.c_float * 16)]


class ArahStats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ('rays', 'trace_sdf_evals', 'iso_rays', 'iso_g_evals', 'on_samples',
                                         'corr_skin_evals', 'shaded_samples', 'hit_rays', 'vol_rays', 'kernel_launches',
                                         'pack_launches')] + \
               [(n, C.c_double) for n in ('ms_trace', 'ms_iso', 'ms_sample_corr', 'ms_shade', 'ms_composite', 'ms_total')] + \
               [('culled_samples', C.c_int64)]

    def as_dict(self):
        return {n: (int(getattr(self, n)) if t is C.c_int64 else float(getattr(self, n))) for n, t in self._fields_}


class ArahTrainGrads(C.Structure):
    _fields_ = [('sdf_W', FP * 7), ('sdf_b', FP * 7), ('sdf_freq', FP), ('sdf_phase', FP), ('skin_W', FP * 5), ('skin_b', FP * 5),
                ('col_W', FP * 6), ('col_b', FP * 6), ('latent', FP), ('beta', FP)]


class ArahHyperWeights(C.Structure):
    _fields_ = [('pe_l0_W', FP), ('pe_l0_b', FP), ('pe_W1', FP), ('pe_b1', FP), ('pe_W2', FP), ('pe_b2', FP),
                ('map_W', FP * 4), ('map_b', FP * 4),
                ('fc1_W', FP * 7), ('fc1_b', FP * 7), ('ln1_g', FP * 7), ('ln1_b', FP * 7),
                ('fc2_W', FP * 7), ('fc2_b', FP * 7), ('ln2_g', FP * 7), ('ln2_b', FP * 7),
                ('out_W', FP * 7), ('out_b', FP * 7), ('init', FP * 7), ('rel_joints', C.c_int32)]


class ArahSdfParams(C.Structure):
    _fields_ = [('sdf_W', FP * 7), ('sdf_b', FP * 7), ('sdf_freq', FP), ('sdf_phase', FP)]


class ArahLossConfig(C.Structure):
    _fields_ = [(n, C.c_float) for n in ('rgb_weight', 'perceptual_weight', 'eikonal_weight', 'mask_weight', 'off_surface_weight', 'inside_weight',
                                         'params_weight', 'skinning_weight')] + [('rgb_loss_type', C.c_int32)]


class ArahLossInputs(C.Structure):
    _fields_ = [(n, FP) for n in ('rgb_values', 'rgb_gt', 'network_body_mask', 'body_mask', 'off_surface_mask', 'sdf_output', 'grad_theta',
                                  'off_surface_sdf', 'inside_sdf', 'pred_weights', 'sampled_weights')] + \
               [('sdf_params', FP * 8), ('sdf_params_count', C.c_int64 * 8)] + \
               [(n, C.c_int32) for n in ('n_rays', 'n_eikonal', 'n_off', 'n_inside', 'n_skin', 'n_joints', 'n_param_tensors')]


class ArahLossGrads(C.Structure):
    _fields_ = [(n, FP) for n in ('rgb_values', 'sdf_output', 'grad_theta', 'off_surface_sdf', 'inside_sdf', 'pred_weights')] + [('sdf_params', FP * 8)]


class ArahRasterCamera(C.Structure):
    _fields_ = [('R', C.c_float * 9), ('T', C.c_float * 3), ('fx', C.c_float), ('fy', C.c_float), ('px', C.c_float), ('py', C.c_float)]


EXPORTS = ['arah_last_error', 'arah_version', 'arah_create', 'arah_destroy', 'arah_set_frame', 'arah_set_profiling', 'arah_render',
           'arah_render_host', 'arah_get_trace', 'arah_get_stats', 'arah_eval_sdf', 'arah_eval_skin', 'arah_debug_umma_gemm', 'arah_debug_umma_f16', 'arah_debug_phase_clocks',
           'arah_set_training', 'arah_train_trace', 'arah_train_shade_forward', 'arah_train_shade_backward', 'arah_train_sdf_forward',
           'arah_train_sdf_backward', 'arah_train_skin_forward', 'arah_train_skin_backward', 'arah_debug_train_gemm',
           'arah_sdf_grid', 'arah_sdf_grid_banded', 'arah_marching_cubes', 'arah_mc_case_table', 'arah_debug_knn', 'arah_marching_cubes_workspace',
           'arah_hyper_forward', 'arah_hyper_workspace', 'arah_pose_smpl', 'arah_frame_rays', 'arah_frame_rays_workspace',
           'arah_frame_images', 'arah_frame_images_workspace', 'arah_psnr', 'arah_psnr_workspace', 'arah_rasterize_mesh',
           'arah_rasterize_mesh_workspace', 'arah_face_normal_image', 'arah_idhr_loss', 'arah_idhr_loss_workspace', 'arah_ssim', 'arah_ssim_workspace']

_lib = None


class ArahError(RuntimeError):
    pass


def declare_image_and_loss(L):
    """Signatures of the image-tail and loss entry points (also applied to the CPU-emulated build of the same sources that
    tests/test_cuda_emu.py loads — test infrastructure)."""
    L.arah_frame_images_workspace.argtypes = [C.c_int32, C.c_int32]
    L.arah_frame_images_workspace.restype = C.c_size_t
    L.arah_frame_images.argtypes = [FP, FP, FP, C.c_int32, C.c_int32, C.c_int32, FP, FP, FP, C.c_size_t, C.c_void_p]
    L.arah_psnr_workspace.argtypes = []
    L.arah_psnr_workspace.restype = C.c_size_t
    L.arah_psnr.argtypes = [FP, FP, C.c_int64, FP, FP, C.c_size_t, C.c_void_p]
    L.arah_rasterize_mesh_workspace.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    L.arah_rasterize_mesh_workspace.restype = C.c_size_t
    L.arah_rasterize_mesh.argtypes = [FP, C.c_int32, FP, C.c_int32, C.POINTER(ArahRasterCamera), C.c_int32, C.c_int32, FP, FP, FP, C.c_size_t, C.c_void_p]
    L.arah_face_normal_image.argtypes = [FP, C.c_int32, FP, C.c_int32, FP, C.c_int32, C.c_int32, C.c_float, C.POINTER(C.c_float), C.c_float, FP,
                                         C.c_void_p]
    L.arah_ssim_workspace.argtypes = []
    L.arah_ssim_workspace.restype = C.c_size_t
    L.arah_ssim.argtypes = [FP, FP, FP, C.c_int32, C.c_int32, FP, FP, C.c_size_t, C.c_void_p]
    L.arah_idhr_loss_workspace.argtypes = []
    L.arah_idhr_loss_workspace.restype = C.c_size_t
    L.arah_idhr_loss.argtypes = [C.POINTER(ArahLossConfig), C.POINTER(ArahLossInputs), FP, C.POINTER(ArahLossGrads), FP, C.c_size_t, C.c_void_p]


def lib():
    """Load the CUDA library.  Raises if it has not been built (python -m arah_release_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise ArahError(f'{SO} is missing: build it with `python -m arah_release_b200.build` '
                        f'(__graft_entry__.build()); there is no CPU fallback')
    L = C.CDLL(SO)
    L.arah_last_error.restype = C.c_char_p
    L.arah_create.argtypes = [C.POINTER(ArahConfig), C.POINTER(C.c_void_p)]
    L.arah_destroy.argtypes = [C.c_void_p]
    L.arah_set_frame.argtypes = [C.c_void_p, C.POINTER(ArahFrame), C.c_void_p]
    L.arah_set_profiling.argtypes = [C.c_void_p, C.c_int32]
    L.arah_render.argtypes = [C.c_void_p, FP, FP, C.c_int32, FP, FP, FP, FP, C.c_void_p]
    L.arah_render_host.argtypes = [C.c_void_p, FP, FP, C.c_int32, FP, FP, FP, C.c_void_p]
    L.arah_get_trace.argtypes = [C.c_void_p, FP, FP, FP, FP, FP, FP, FP, C.c_void_p]
    L.arah_get_stats.argtypes = [C.c_void_p, C.POINTER(ArahStats), C.c_void_p]
    L.arah_eval_sdf.argtypes = [C.c_void_p, FP, C.c_int32, FP, FP, FP, C.c_void_p]
    L.arah_eval_skin.argtypes = [C.c_void_p, FP, C.c_int32, FP, FP, C.c_void_p]
    L.arah_debug_umma_gemm.argtypes = [FP, FP, C.c_int32, C.c_int32, FP, C.c_int32, C.c_void_p]
    L.arah_debug_phase_clocks.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]
    L.arah_set_training.argtypes = [C.c_void_p, C.c_int32]
    L.arah_train_trace.argtypes = [C.c_void_p, FP, FP, C.c_int32, FP, FP, FP, C.c_void_p]
    L.arah_train_shade_forward.argtypes = [C.c_void_p, FP, FP, C.c_int32, C.c_int32, FP, FP, C.c_void_p]
    L.arah_train_shade_backward.argtypes = [C.c_void_p, FP, FP, C.POINTER(ArahTrainGrads), C.c_void_p]
    L.arah_train_sdf_forward.argtypes = [C.c_void_p, C.c_int32, FP, C.c_int32, C.c_int32, FP, FP, C.c_void_p]
    L.arah_train_sdf_backward.argtypes = [C.c_void_p, C.c_int32, FP, FP, C.POINTER(ArahTrainGrads), C.c_void_p]
    L.arah_train_skin_forward.argtypes = [C.c_void_p, FP, C.c_int32, FP, C.c_void_p]
    L.arah_train_skin_backward.argtypes = [C.c_void_p, FP, C.POINTER(ArahTrainGrads), C.c_void_p]
    L.arah_debug_train_gemm.argtypes = [C.c_int32, C.c_int32, C.c_int32, FP, C.c_int64, C.c_int64, FP, C.c_int64, C.c_int64, FP, C.c_int32, FP,
                                        C.c_int32, C.c_int32, C.c_void_p]
    L.arah_sdf_grid.argtypes = [C.c_void_p, C.c_int32, FP, C.c_void_p]
    L.arah_sdf_grid_banded.argtypes = [C.c_void_p, C.c_int32, C.c_float, C.c_float, FP, FP, C.c_void_p]
    L.arah_marching_cubes.argtypes = [FP, C.c_int32, C.c_float, C.c_float, C.POINTER(C.c_float), FP, C.c_int32, FP, C.c_int32, FP, FP, C.c_size_t, C.c_void_p]
    L.arah_marching_cubes_workspace.argtypes = [C.c_int32]
    L.arah_marching_cubes_workspace.restype = C.c_size_t
    L.arah_mc_case_table.argtypes = [C.c_void_p, C.c_void_p]
    L.arah_debug_knn.argtypes = [C.c_void_p, FP, C.c_int32, FP, C.c_void_p]
    L.arah_hyper_forward.argtypes = [C.POINTER(ArahHyperWeights), FP, FP, FP, C.POINTER(ArahSdfParams), FP, C.c_void_p]
    L.arah_hyper_workspace.restype = C.c_size_t
    L.arah_pose_smpl.argtypes = [FP, FP, FP, FP, FP, C.POINTER(C.c_float), C.c_int32, C.c_float, FP, FP, FP, C.c_void_p]
    F9, F3 = C.POINTER(C.c_float), C.POINTER(C.c_float)
    L.arah_frame_rays.argtypes = [F9, F9, F9, F3, F3, FP, C.c_int32, C.c_int32, FP, FP, FP, FP, FP, FP, FP, FP, C.c_size_t, C.c_void_p]
    L.arah_frame_rays_workspace.argtypes = [C.c_int32, C.c_int32]
    L.arah_frame_rays_workspace.restype = C.c_size_t
    declare_image_and_loss(L)
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise ArahError(f'arah_b200 error {rc}: {lib().arah_last_error().decode()}')
