/*
 * arah_b200.h — C ABI of the B200-native ARAH hot path (libarah_b200.so).
 *
 * The reference (taconite/arah-release) is pure Python/PyTorch and has no FFI of its own; the entry points below are
 * what a binding for its renderer modules would call.  Each one names the reference interface it replaces
 * (paths relative to /root/reference/im2mesh).  All pointers are plain fp32 / uint8 / int32 buffers; nothing in this
 * header depends on torch.  Unless stated otherwise pointers are DEVICE pointers borrowed from the caller, must be
 * contiguous and stay alive until the stream has been synchronised; every call is asynchronous on `stream`
 * (a cudaStream_t passed as void*) and performs no host synchronisation.
 *
 * Error model: every function returns 0 on success or a negative ARAH_E* code; arah_last_error() returns a
 * thread-local message.  No C++ exception crosses the boundary.  A handle is not thread-safe; use one per stream.
 */
#ifndef ARAH_B200_H
#define ARAH_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ARAH_OK 0
#define ARAH_EINVAL (-1)   /* bad argument / unsupported configuration */
#define ARAH_ECUDA (-2)    /* CUDA runtime error (message has the CUDA string) */
#define ARAH_ENOMEM (-3)
#define ARAH_ESTATE (-4)   /* call order violated (e.g. render before set_frame) */

typedef struct ArahHandle ArahHandle;

/* Static configuration == constructor arguments of BodyRayTracing / IDHRNetwork
 * (metaavatar_render/renderer/ray_tracing.py:16-49, renderer/implicit_differentiable_renderer.py:18-40)
 * and the network shapes of configs/arah-zju/ZJUMOCAP-377_4gpus.yaml:34-43. */
typedef struct ArahConfig {
    int32_t device;               /* CUDA device ordinal */
    int32_t n_steps;              /* model.n_steps (64) */
    int32_t near_samples;         /* model.near_surface_samples (16) */
    int32_t far_samples;          /* model.far_surface_samples (16) */
    int32_t cano_view_dirs;       /* model.cano_view_dirs */
    int32_t latent_dim;           /* colour-net per-frame latent (128; 0 = none) */
    int32_t n_verts;              /* SMPL vertices (6890) */
    int32_t max_rays;             /* initial workspace size in rays (grown on demand) */
    int32_t shade_mode;           /* ARAH_SHADE_TF32 (default 0; the name is historical): shading MLPs on tcgen05 tensor cores with
                                   * 11-bit-significand operands (fp16 images since round 2, TF32 in round 1), fp32 accumulate;
                                   * ARAH_SHADE_FP32 (1): fp32 FFMA tiles (bit-for-bit the oracle's arithmetic order).
                                   * Independent of root_mode. */
    int32_t root_mode;            /* ARAH_ROOT_3XTF32 (default 0; the name is historical): every MLP of the root-finding stages
                                   * (sphere tracing, joint search, per-sample correspondence search) on tcgen05 in split
                                   * precision (hi/lo fp16 or TF32 operands, 3 products ~ one fp32 product), persistent kernels;
                                   * ARAH_ROOT_FP32 (1): fp32 FFMA tiles, one launch per iteration.  Residual bookkeeping,
                                   * Jacobians and Broyden updates are fp32 in both. */
    int32_t shade_cull;           /* ARAH_CULL_EXACT (default 0): with shade_mode TF32, converged samples whose compositing alpha is
                                   * exactly 0.0f in fp32 skip the SDF-gradient + colour pass — their weight alpha * T is 0, so no
                                   * output bit can depend on their colour (results are bit-identical to ARAH_CULL_OFF, see
                                   * tests/test_gpu_parity.py::test_alpha_cull_is_exact); ARAH_CULL_OFF (1): shade every sample. */
    int32_t render_last_pt;       /* IDHRNetwork(render_last_pt=...), implicit_differentiable_renderer.py:380-381: != 0 gives the last
                                   * converged sample of a ray the interval 1e10 (opaque beyond it) instead of 1 / n_steps */
} ArahConfig;
#define ARAH_CULL_EXACT 0
#define ARAH_CULL_OFF 1
#define ARAH_SHADE_TF32 0
#define ARAH_SHADE_FP32 1
#define ARAH_ROOT_3XTF32 0
#define ARAH_ROOT_FP32 1

/* Per-frame inputs == what IDHRNetwork.forward reads from its `input` dict
 * (renderer/implicit_differentiable_renderer.py:52-71) plus the weights of the modules it owns.
 * Weight matrices are dense fp32 in the reference's own layout [out][in] (weight-norm already applied:
 * w = g * v / ||v||, i.e. what `lin.weight` holds after the forward pre-hook). */
typedef struct ArahFrame {
    /* per-frame SDF FiLM-SIREN produced by the hypernetwork: sdf_network[l][0].weights / .biases / .freq / .phase_shift
     * (hyperlayers.py:391-415); l = 6 is the final BatchLinear (hyperlayers.py:368-388) */
    const float* sdf_W[7];        /* [256][3], 5 x [256][256], [1][256] */
    const float* sdf_b[7];        /* [256] x 6, [1] */
    const float* sdf_freq;        /* [6][256] */
    const float* sdf_phase;       /* [6][256] */
    /* skinning_model.skinning_decoder_fwd.lin0..lin4 (metaavatar/models/decoder.py:133-233) */
    const float* skin_W[5];       /* [128][3], 3 x [128][128], [25][128] */
    const float* skin_b[5];
    /* rendering_network.lin0..lin5 (metaavatar_render/models/decoder.py:10-124), mode 'idr', multires_view 4, skips [3] */
    const float* col_W[6];        /* [256][417], [256][256], [128][256], [256][545], [256][256], [3][256] */
    const float* col_b[6];
    const float* latent;          /* [latent_dim] pose_cond['latent_code'] */
    float beta;                   /* deviation_network.variance (metaavatar_render/models/decoder.py:127-133) */
    /* body / pose buffers; HOST pointers if pose_on_host != 0 (copied with cudaMemcpyAsync), else device pointers */
    const float* bone_transforms; /* [24][4][4] */
    const float* smpl_verts;      /* [n_verts][3]  posed, including trans */
    const float* smpl_weights;    /* [n_verts][24]; may be NULL after the first frame (kept from the previous call) */
    int32_t pose_on_host;
    float trans[3];
    float coord_min, coord_max;   /* scalars (data/zju_mocap_odp.py:326-331) */
    float center[3];
    float cam_loc[3];
    float pose[16];               /* world->camera [R|T], row-major 4x4 */
} ArahFrame;

/* Iteration counters of the last render (algorithmic work is defined through these). */
typedef struct ArahStats {
    int64_t rays;
    int64_t trace_sdf_evals;      /* sphere-tracing SDF evaluations */
    int64_t iso_rays;             /* rays that entered the joint search */
    int64_t iso_g_evals;          /* joint-search g evaluations (excluding the Jacobian init) */
    int64_t on_samples;           /* samples that entered the correspondence search */
    int64_t corr_skin_evals;      /* skinning-net evaluations in the correspondence search (incl. J init + g(x0)) */
    int64_t shaded_samples;       /* converged samples shaded (SDF fwd + grad + colour) */
    int64_t hit_rays;             /* rays with a converged surface point */
    int64_t vol_rays;             /* rays with >= 1 converged sample (rendered) */
    int64_t kernel_launches;      /* CUDA kernels launched by the last render call */
    int64_t pack_launches;        /* CUDA kernels launched by the last arah_set_frame call */
    /* device time of each stage of the last render (ms, CUDA events on the launching stream);
     * only filled when profiling was enabled with arah_set_profiling(h, 1), else 0 */
    double ms_trace, ms_iso, ms_sample_corr, ms_shade, ms_composite, ms_total;
    int64_t culled_samples;       /* of shaded_samples: culled by the exact alpha test (only their SDF value was computed) */
} ArahStats;

const char* arah_last_error(void);
int arah_version(void);

int arah_create(const ArahConfig* cfg, ArahHandle** out);
int arah_destroy(ArahHandle* h);

/* Pack the frame's weights for the kernels (transpose, pad, fold the latent into the colour biases) and stage the
 * pose buffers.  Replaces nothing in the reference one-to-one; it is the boundary crossing that
 * MetaAvatarRender.forward performs by handing modules to idhr_network (metaavatar_render/models/__init__.py:186-200). */
int arah_set_frame(ArahHandle* h, const ArahFrame* frame, void* stream);

/* IDHRNetwork.forward, eval branch (renderer/implicit_differentiable_renderer.py:42-112,141-148,180-259):
 *   ray_dirs [P][3], near_far [P][2] (input['ray_dirs'], input['body_bounds_intersections'])
 *   -> rgb [P][3] ('rgb_values'), mask [P] ('network_body_mask'), points_cam [P][3] ('points_cam'); weights_sum may be NULL. */
int arah_render(ArahHandle* h, const float* ray_dirs, const float* near_far, int32_t P,
                float* rgb, uint8_t* mask, float* points_cam, float* weights_sum, void* stream);

/* Same call with HOST buffers: inputs are copied host->device and results device->host inside the call
 * (pinned buffers recommended); the stream is synchronised before returning. */
int arah_render_host(ArahHandle* h, const float* ray_dirs, const float* near_far, int32_t P,
                     float* rgb, uint8_t* mask, float* points_cam, void* stream);

/* BodyRayTracing.forward (renderer/ray_tracing.py:51-172) outputs of the LAST arah_render call, any pointer may be NULL:
 *   points_hat_norm [P][3], network_body_mask [P], dists [P], sampled_pts [P][S][3], sampled_dists [P][S],
 *   sampled_transforms [P][S][4][4], sampler_converge_mask [P][S]. */
int arah_get_trace(ArahHandle* h, float* points_hat_norm, uint8_t* network_body_mask, float* dists,
                   float* sampled_pts, float* sampled_dists, float* sampled_transforms,
                   uint8_t* sampler_converge_mask, void* stream);

/* Record CUDA events at the stage boundaries of every following arah_render (cheap; no synchronisation). */
int arah_set_profiling(ArahHandle* h, int32_t enable);

/* Synchronises `stream` and reads the device counters of the last render. */
int arah_get_stats(ArahHandle* h, ArahStats* stats, void* stream);

/* Unit-level entry points used by the parity tests (device pointers, n points):
 *   arah_eval_sdf  : sdf_network forward + gradient + feature (hyperlayers.py:412-415; diff_operators.py:39-50)
 *   arah_eval_skin : query_weights + skinning (utils/root_finding_utils.py:54-113, 13-33) */
int arah_eval_sdf(ArahHandle* h, const float* xn, int32_t n, float* sdf, float* grad, float* feat, void* stream);
int arah_eval_skin(ArahHandle* h, const float* x_hat, int32_t n, float* weights, float* x_bar, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Canonical mesh extraction (MetaAvatarRender.forward(gen_cano_mesh=True),
 * metaavatar_render/models/__init__.py:203-224).
 *
 * arah_sdf_grid: utils/sdf_meshing.py:13-58 — the frame's SDF network sampled on the N^3 lattice over [-1,1]^3
 *   (voxel_size = 2/(N-1), lattice point (ix,iy,iz) -> index (ix*N + iy)*N + iz, coordinates ix*voxel_size - 1 with one fp32
 *   rounding per operation as in the reference); sdf is a device buffer of N^3 floats holding the raw network output.
 *   Runs on the tensor cores in split precision (root_mode 3xTF32) or on fp32 FFMA tiles (root_mode FP32).
 * arah_marching_cubes: utils/sdf_meshing.py:69-114 — iso-surface `level` of a device lattice sdf[N][N][N]:
 *   verts [max_verts][3] = origin + (lattice index + t) * voxel_size  (the reference's `voxel_grid_origin + verts`),
 *   faces [max_faces][3] int32 vertex ids, outward (towards larger values) orientation, counts[0..1] (device) = number of
 *   vertices / faces the surface HAS; if a count exceeds its max_* the output was truncated — call again with larger buffers.
 *   One vertex per sign-changing lattice edge (t = v0/(v0-v1)); deterministic order (lattice order).  The reference delegates
 *   this step to skimage 0.18.1 `marching_cubes_lewiner` (absent here: parity against it is unpinned; the checker is
 *   oracle/mc_oracle.c, a CPU restatement of the same published scheme).  Works on the current device; no handle needed. */
int arah_sdf_grid(ArahHandle* h, int32_t N, float* sdf, void* stream);
/* arah_sdf_grid_banded: the same lattice for the caller that only feeds it to arah_marching_cubes at `level`
 *   (utils/sdf_meshing.py:59-114 does exactly that), in two precisions.  Every point is evaluated in one fp16 tensor-core pass;
 *   every cell whose eight coarse corner values satisfy min - eps <= level <= max + eps has all its corners re-evaluated in split
 *   precision (bit-identical to arah_sdf_grid's values).  As long as the coarse error stays below eps, marching cubes finds exact
 *   values wherever it interpolates and exact signs everywhere else, so its vertices and faces are bit-identical to those of the
 *   full-precision lattice.  stats (device, 2 x int32): [0] points refined, [1] refined points whose coarse value was off by more
 *   than eps — a run-time check of the bound on the ~2 % of the lattice nearest the surface; non-zero means: call arah_sdf_grid.
 *   Values far from the level keep their fp16-pass accuracy (~1e-3).  Needs root_mode ARAH_ROOT_3XTF32. */
int arah_sdf_grid_banded(ArahHandle* h, int32_t N, float level, float eps, float* sdf, int32_t* stats, void* stream);
int arah_marching_cubes(const float* sdf, int32_t N, float level, float voxel_size, const float* origin3 /* host [3] */,
                        float* verts, int32_t max_verts, int32_t* faces, int32_t max_faces, int32_t* counts,
                        void* workspace /* device, 16-byte aligned */, size_t workspace_bytes, void* stream);
/* Bytes of device scratch arah_marching_cubes needs for an N^3 lattice (~9.03 N^3); the library never allocates behind the call. */
size_t arah_marching_cubes_workspace(int32_t N);
/* Host-only: the generated marching-cubes case table, tri[256][16] (edge ids, -1 terminated) and ntri[256]. */
int arah_mc_case_table(int8_t* tri, uint8_t* ntri);

/* ------------------------------------------------------------------------------------------------------------------
 * Hypernetwork forward: pose -> the frame's SDF parameters, i.e. HyperBVPNet.forward up to the
 * assembled decoder (metaavatar/models/siren_modules.py:280-312; called at metaavatar_render/models/__init__.py:181-183).
 * All pointers are DEVICE pointers in the reference's own state_dict layout ([out][in] row-major), borrowed for the call —
 * the 341 MB of output-layer matrices are read in place, never copied.  Shapes are those of
 * configs/arah-zju/ZJUMOCAP-377_4gpus.yaml:34 (hyper_in_ch 144, hidden 256, 5 hidden SIREN layers, use_FiLM). */
typedef struct ArahHyperWeights {
    /* pose_encoder.* — HierarchicalPoseEncoder (siren_modules.py:196-244) */
    const float* pe_l0_W;         /* layer_0.weight [6][288] */
    const float* pe_l0_b;         /* [6] */
    const float* pe_W1;           /* layers.j.0.weight stacked [24][19][19] */
    const float* pe_b1;           /* [24][19] */
    const float* pe_W2;           /* layers.j.2.weight stacked [24][6][19] */
    const float* pe_b2;           /* [24][6] */
    /* net.mapping_network.network.{0,2,4,6} — CustomMappingNetwork (hyperlayers.py:107-139) */
    const float* map_W[4];        /* [256][128], [256][256], [256][256], [3072][256] */
    const float* map_b[4];
    /* net.layers.l.(hyper_linear.)hypo_params.net.{0,1,2} — FCBlock (pytorch_prototyping.py:50-81), l = 0..6 */
    const float* fc1_W[7];        /* net.0.net.0.weight [256][144] */
    const float* fc1_b[7];
    const float* ln1_g[7];        /* net.0.net.1.weight (LayerNorm) [256] */
    const float* ln1_b[7];
    const float* fc2_W[7];        /* net.1.net.0.weight [256][256] */
    const float* fc2_b[7];
    const float* ln2_g[7];
    const float* ln2_b[7];
    const float* out_W[7];        /* net.2.weight [n_l][256], n_l = in_l*out_l + out_l: 1024, 5 x 65792, 257 */
    const float* out_b[7];        /* [n_l] */
    const float* init[7];         /* hypo_params_init [n_l] (hyperlayers.py:443,487); may be NULL (= zeros) */
    int32_t rel_joints;           /* HierarchicalPoseEncoder(rel_joints=...) */
} ArahHyperWeights;

/* Outputs == the tensors HyperFCFiLM.forward hands to BatchLinearFiLM / BatchLinear (hyperlayers.py:270-285, 453-510) and
 * ArahFrame consumes: device buffers owned by the caller. */
typedef struct ArahSdfParams {
    float* sdf_W[7];              /* [256][3], 5 x [256][256], [1][256] */
    float* sdf_b[7];              /* [256] x 6, [1] */
    float* sdf_freq;              /* [6][256] */
    float* sdf_phase;             /* [6][256] */
} ArahSdfParams;

/* rots [24][9] ('rots': rotation matrices, root = identity), Jtrs [24][3] ('Jtrs': normalised joints), latent [128] (the
 * geometry latent code; NULL = zeros).  workspace: arah_hyper_workspace() bytes of device scratch, 16-byte aligned.
 * Two kernel launches on `stream`, no synchronisation. */
int arah_hyper_forward(const ArahHyperWeights* w, const float* rots, const float* Jtrs, const float* latent,
                       const ArahSdfParams* out, void* workspace, void* stream);
size_t arah_hyper_workspace(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Per-frame ray set-up: what the dataset classes compute on the CPU for every frame
 * (data/zju_mocap_odp.py:250-315).  Device pointers unless marked host; asynchronous on `stream`; no allocation inside.
 *
 * arah_pose_smpl (:268-289): minimal_shape [n][3] + posedirs [3n][207] . pose_feature [207] (DOUBLE, as scipy hands it over; the
 *   product is accumulated in fp64 like numpy does), T = skinning_weights [n][24] . bone_transforms [24][16], posed vertices
 *   = T [v; 1] + trans -> verts [n][3] (== ArahFrame.smpl_verts), bounds [2][3] = min / max -/+ box_margin.
 *   workspace: 32 bytes of device scratch.
 * arah_frame_rays (:291-315, utils/utils.py:17-73): K, K_inv, R, T, cam_loc are HOST arrays (row-major 3x3 / 3; K already rescaled,
 *   K_inv = inverse(K), cam_loc = -R^T T as the caller's numpy computes them); bounds [2][3] on the device.  mask_in (H*W
 *   bytes) == NULL: the bounding-box mask of get_bound_2d_mask is rasterised into bound_mask (six cv2.fillPoly calls restated
 *   in integer arithmetic; identical to OpenCV 4.13 for boxes that project inside the image, see oracle/rays_oracle.py).
 *   Outputs, in np.where order (row-major): pix [P] = y*W + x, ray_dirs [P][3], near_far [P][2] (the rows with near < far),
 *   image_mask [H*W] bytes, count[0] = P (device).  Caller sizes pix / ray_dirs / near_far for H*W rays. */
int arah_pose_smpl(const float* minimal_shape, const float* posedirs, const double* pose_feature, const float* skinning_weights,
                   const float* bone_transforms, const float* trans3 /* host */, int32_t n_verts, float box_margin, float* verts,
                   float* bounds, void* workspace, void* stream);
int arah_frame_rays(const float* K, const float* K_inv, const float* R, const float* T, const float* cam_loc /* 5 host arrays */,
                    const float* bounds, int32_t H, int32_t W, const uint8_t* mask_in, uint8_t* bound_mask, int32_t* pix,
                    float* ray_dirs, float* near_far, uint8_t* image_mask, int32_t* count, void* workspace, size_t workspace_bytes,
                    void* stream);
size_t arah_frame_rays_workspace(int32_t H, int32_t W);

/* ------------------------------------------------------------------------------------------------------------------
 * Image-space tail of the validation / test step (the parts after the renderer).  Device pointers
 * unless noted; launches are asynchronous on `stream`; caller-owned workspaces sized by the *_workspace functions.
 *
 * arah_frame_images (im2mesh/metaavatar_render/lightning_model.py:176-205): rgb / points_cam [P][3] in ray order, pix [P] =
 *   y*W + x of ray k (np.where order of image_mask — arah_frame_rays' `pix`): `masked_scatter_` into pred_pixels [H][W][3]
 *   (background 0) and the finite-difference normal map of the scattered point image -> pred_normals [H][W][3] in [0,1]
 *   (NaN -> -1 before the (n+1)/2 clip, as the reference).  Either output may be NULL.
 * arah_psnr (:218-221, im2mesh/utils/eval.py:6-9): mse_psnr[0] = mean((pred - gt)^2) rounded to fp32, [1] = -10 log10(mse),
 *   both stored as doubles on the device; n = number of floats (3 P for a ray list).  Bit-reproducible.
 * arah_rasterize_mesh (im2mesh/metaavatar_render/models/__init__.py:238-254,265-277,291-299 — pytorch3d MeshRasterizer with
 *   image_size (H, W), faces_per_pixel 1, blur_radius 0, perspective-correct depth): verts [n_verts][3] world space, faces
 *   [n_faces][3] int32 -> pix_to_face [H][W] (index of the nearest face covering the pixel centre, lowest index on depth
 *   ties, -1 = background), zbuf [H][W] (view depth, -1 = background; may be NULL).  `cam` is a HOST struct in pytorch3d's
 *   convention: row vectors, view = X R + T, +X left / +Y up / +Z forward, ndc = (fx x + px z, fy y + py z) / z
 *   (FoVPerspectiveCameras: fx = fy = 1/tan(fov/2), px = py = 0; cameras_from_opencv_projection: see images.py).
 *   Faces with a vertex at or behind the camera plane are not drawn.
 * arah_face_normal_image (:256-263, 279-286, 301-308): image [H][W][3] = clip((v + 1) / 2, 0, 1) with v = rot3x3 (HOST,
 *   row-major, may be NULL) . (sign * unit face normal of pix_to_face) for covered pixels, v = background otherwise. */
typedef struct ArahRasterCamera { float R[9]; float T[3]; float fx, fy, px, py; } ArahRasterCamera;
size_t arah_frame_images_workspace(int32_t H, int32_t W);
int arah_frame_images(const float* rgb, const float* points_cam, const int32_t* pix, int32_t P, int32_t H, int32_t W, float* pred_pixels,
                      float* pred_normals, void* workspace, size_t workspace_bytes, void* stream);
size_t arah_psnr_workspace(void);
int arah_psnr(const float* pred, const float* gt, int64_t n, double* mse_psnr, void* workspace, size_t workspace_bytes, void* stream);
/* arah_ssim (lightning_model.py:222, im2mesh/utils/eval.py:11-19): crop pred_image / gt_image [H][W][3] to cv2.boundingRect(mask [H*W]
 *   bytes) and take skimage.metrics.structural_similarity(multichannel=True) with its defaults (uniform 7x7 window, sample covariance,
 *   float64, data_range 2 for float images): out5[0] = SSIM (NaN if the crop is smaller than the window — skimage raises),
 *   out5[1..4] = x, y, w, h of the rectangle; doubles on the device.  Bit-reproducible; no host synchronisation. */
size_t arah_ssim_workspace(void);
int arah_ssim(const float* pred_image, const float* gt_image, const uint8_t* mask, int32_t H, int32_t W, double* out5, void* workspace,
              size_t workspace_bytes, void* stream);
size_t arah_rasterize_mesh_workspace(int32_t n_verts, int32_t H, int32_t W);
int arah_rasterize_mesh(const float* verts, int32_t n_verts, const int32_t* faces, int32_t n_faces, const ArahRasterCamera* cam, int32_t H,
                        int32_t W, int32_t* pix_to_face, float* zbuf, void* workspace, size_t workspace_bytes, void* stream);
int arah_face_normal_image(const float* verts, int32_t n_verts, const int32_t* faces, int32_t n_faces, const int32_t* pix_to_face, int32_t H,
                           int32_t W, float sign, const float* rot3x3, float background, float* image, void* stream);

/* Unit-level: pytorch3d.ops.knn_points(K=1) as used at renderer/ray_tracing.py:386,407 — index of the nearest posed SMPL vertex
 * (exact fp32 argmin of (x-v).(x-v), lowest index on ties) for n device points [n][3] -> idx [n] int32. */
int arah_debug_knn(ArahHandle* h, const float* pts, int32_t n, int32_t* idx, void* stream);

/* Debug: SM-clock cycles spent per kernel phase by one designated thread per CTA, summed over CTAs and launches of the last
 * profiled render (arah_set_profiling(h,1)): out32[0..5] correspondence step (gather, layer 0, MMA wait, epilogues, output
 * layer, per-point phase), out16[8..14] shading (setup+layer0, fwd wait, fwd epilogue, rev wait, rev epilogue, colour inputs,
 * colour MLP), out32[16..20] sphere-tracing step (gather, layer 0, MMA wait, epilogues, marching).  32 entries. */
int arah_debug_phase_clocks(ArahHandle* h, uint64_t* out32, void* stream);

/* Debug/bring-up: D[128][N] = A[128][K] . W[N][K]^T through the tcgen05 TF32 tile used by the shading kernel
 * (device pointers, K multiple of 32 <= 256, N in {128, 256}; a_in_tmem != 0 stages A in tensor memory and uses the
 * `.ts` MMA form); synchronises the stream. */
int arah_debug_umma_gemm(const float* A, const float* W, int32_t K, int32_t N, float* D, int32_t a_in_tmem, void* stream);

/* Debug/bring-up: the same product through the fp16 split-precision path of the persistent root-finding kernels
 * (tcgen05 kind::f16, A in tensor memory as packed half pairs, weights as pre-scaled hi / lo chunk images;
 * csrc/arah_f16x3.cuh).  K in {64, 128}, N in {32, 128, 256}; mode bit 0: three-pass split product
 * (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, fp32-grade) instead of hi.hi only; synchronises the stream. */
int arah_debug_umma_f16(const float* A, const float* W, int32_t K, int32_t N, float* D, int32_t mode, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Training (BASELINE configs[2]; IDHRNetwork.forward with self.training == True,
 * renderer/implicit_differentiable_renderer.py:73-78,117-178,235-249, get_rbg_value_vol_sdf :261-396).
 *
 * The tracer runs without gradient (the reference wraps it in torch.no_grad, :84); the differentiable part — implicit-
 * gradient LBS correction, SDF forward, SDF input gradient as normal (differentiated a second time in backward), colour
 * MLP, sigma-from-SDF compositing, plus the auxiliary SDF / skinning evaluations of the regularisers — is a hand-written
 * forward/backward pair per entry point.  The host side (renderer.py) wraps each pair in a torch.autograd.Function.
 * Weights are those of the last arah_set_frame issued AFTER arah_set_training(h, 1) (raw reference-layout copies are kept
 * for the backward).  Gradient buffers are caller-zeroed fp32 device buffers in the layout of the corresponding ArahFrame
 * member ([out][in], weight-norm applied: d L / d (g v / ||v||)); entry points ACCUMULATE into them; NULL members are skipped.
 * One forward may be outstanding per context: shading (1), SDF evaluation slots 0..2, skinning evaluation (1).
 * arah_train_shade_forward synchronises the stream once (it reads the number of converged samples). */
typedef struct ArahTrainGrads {
    float* sdf_W[7]; float* sdf_b[7]; float* sdf_freq; float* sdf_phase;
    float* skin_W[5]; float* skin_b[5];
    float* col_W[6]; float* col_b[6];
    float* latent;                /* [latent_dim] */
    float* beta;                  /* [1]  d L / d ||variance|| */
} ArahTrainGrads;

/* mode: 0 = off (frees the training state); ARAH_TRAIN_3XTF32 (default) = the engine's GEMMs on tcgen05 tensor cores in split
 * precision (hi/lo TF32 operands, three MMAs per K-step, fp32 accumulation in TMEM: fp32-class results);
 * ARAH_TRAIN_FP32 = fp32 SIMT FFMA GEMMs; ARAH_TRAIN_TF32 = single-pass TF32 operands (fastest; gradient cosine vs fp32 drops
 * to ~0.96 on the first SIREN layer, see tests/test_gpu_train.py). */
#define ARAH_TRAIN_3XTF32 1
#define ARAH_TRAIN_FP32 2
#define ARAH_TRAIN_TF32 3
int arah_set_training(ArahHandle* h, int32_t mode);

/* BodyRayTracing.forward(eval_mode=False) (renderer/ray_tracing.py:51-172): every ray enters the joint search (:249) and the
 * z samples are jittered (:298-311) with the caller's three uniform draws u_all [P][n_steps], u_near [P][near+1],
 * u_far [P][far] (the reference draws them with torch.rand on the CPU generator in this order).  Results: arah_get_trace. */
int arah_train_trace(ArahHandle* h, const float* ray_dirs, const float* near_far, int32_t P, const float* u_all, const float* u_near,
                     const float* u_far, void* stream);

/* Differentiable shading of the samples of the last arah_train_trace.  view_dirs [P][3]: ray directions after view
 * augmentation (:152-162), view_dirs_orig [P][3] (may be NULL = same); ray_augm: apply the back-facing test of :342-350.
 * -> rgb [P][3] ('rgb_values'), weights_sum [P] ('sdf_output'); both 0 for rays without a converged sample. */
int arah_train_shade_forward(ArahHandle* h, const float* view_dirs, const float* view_dirs_orig, int32_t ray_augm,
                             int32_t train_skinning_net, float* rgb, float* weights_sum, void* stream);
int arah_train_shade_backward(ArahHandle* h, const float* g_rgb, const float* g_weights_sum, const ArahTrainGrads* grads, void* stream);

/* sdf_network(points) and (with_grad) gradient(sdf, points) for the eikonal / off-surface / inside terms (:117-140):
 * points [n][3] normalised -> sdf [n] (raw network output), grad [n][3].  Backward takes d L / d sdf and d L / d grad. */
int arah_train_sdf_forward(ArahHandle* h, int32_t slot, const float* points, int32_t n, int32_t with_grad, float* sdf, float* grad, void* stream);
int arah_train_sdf_backward(ArahHandle* h, int32_t slot, const float* g_sdf, const float* g_grad, const ArahTrainGrads* grads, void* stream);

/* query_weights(points_skinning) (:73-78; utils/root_finding_utils.py:54-113): points [n][3] in metres -> weights [n][24]. */
int arah_train_skin_forward(ArahHandle* h, const float* points, int32_t n, float* weights, void* stream);
int arah_train_skin_backward(ArahHandle* h, const float* g_weights, const ArahTrainGrads* grads, void* stream);

/* IDHRLoss.forward (im2mesh/metaavatar_render/renderer/loss.py:122-200) and its gradient in one fused pass (SURVEY.md §8 row f2,
 * "fused loss reductions"): terms[9] (device floats) = loss, rgb_loss, perceptual_loss (always 0: LPIPS is outside this path and
 * perceptual_weight > 0 is rejected), eikonal_loss, mask_loss, off_surface_loss, inside_loss, sdf_params_loss, skinning_loss;
 * grads (may be NULL = forward only; NULL members are skipped) receive d loss / d input, weights included, written (not
 * accumulated) over the full extent of each buffer.  Inputs are the step's outputs with the batch dimension dropped and cut to
 * the first 2048 rays as the reference does (:124-127,132): rgb_values / rgb_gt [n_rays][3], network_body_mask / body_mask /
 * off_surface_mask [n_rays] bytes (body_mask keeps its values: 0 / 1, 100 = patch border, :52-54), sdf_output [n_rays]
 * ('sdf_output' = the rays' weight sums), grad_theta [n_eikonal][3], off_surface_sdf [n_off], inside_sdf [n_inside], pred_weights /
 * sampled_weights [n_skin][n_joints], sdf_params = the hypernetwork's weight matrices flattened (siren_modules.py:310-314).
 * A term whose weight is <= 0 is 0 and its inputs are not read (the reference returns torch.zeros(1) for it).  The mask term is
 * the 2-norm of the whole off-surface difference vector over N, as torch.norm(dim=-1) of the reference's 1-D difference gives
 * (:100-101; see csrc/arah_loss_core.h).  No host synchronisation; results are bit-reproducible (integer atomic + fixed trees).  Config / input / gradient structs are HOST structs of device pointers. */
#define ARAH_LOSS_MAX_PARAM_TENSORS 8
typedef struct ArahLossConfig {
    float rgb_weight, perceptual_weight, eikonal_weight, mask_weight, off_surface_weight, inside_weight, params_weight, skinning_weight;
    int32_t rgb_loss_type;                     /* 0 'l1', 1 'mse', 2 'smoothed_l1' (beta 0.1), loss.py:34-41 */
} ArahLossConfig;
typedef struct ArahLossInputs {
    const float* rgb_values; const float* rgb_gt; const uint8_t* network_body_mask; const uint8_t* body_mask; const uint8_t* off_surface_mask;
    const float* sdf_output; const float* grad_theta; const float* off_surface_sdf; const float* inside_sdf;
    const float* pred_weights; const float* sampled_weights;
    const float* sdf_params[ARAH_LOSS_MAX_PARAM_TENSORS]; int64_t sdf_params_count[ARAH_LOSS_MAX_PARAM_TENSORS];
    int32_t n_rays, n_eikonal, n_off, n_inside, n_skin, n_joints, n_param_tensors;
} ArahLossInputs;
typedef struct ArahLossGrads {
    float* rgb_values; float* sdf_output; float* grad_theta; float* off_surface_sdf; float* inside_sdf; float* pred_weights;
    float* sdf_params[ARAH_LOSS_MAX_PARAM_TENSORS];
} ArahLossGrads;
size_t arah_idhr_loss_workspace(void);
int arah_idhr_loss(const ArahLossConfig* cfg, const ArahLossInputs* in, float* terms, const ArahLossGrads* grads, void* workspace,
                   size_t workspace_bytes, void* stream);

/* Debug/bring-up: the training engine's strided GEMM, C[i][j] (+)= bias[j] + sum_k A[i sa_i + k sa_k] B[k sb_k + j sb_j]
 * (device pointers, element strides), mode = ARAH_TRAIN_3XTF32 / ARAH_TRAIN_TF32 (tcgen05) or ARAH_TRAIN_FP32 (SIMT). */
int arah_debug_train_gemm(int32_t M, int32_t N, int32_t K, const float* A, int64_t sa_i, int64_t sa_k, const float* B, int64_t sb_k,
                          int64_t sb_j, float* C, int32_t ldc, const float* bias, int32_t accumulate, int32_t mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif
