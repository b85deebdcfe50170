// arah_image.cu — image-space tail of the validation / test step on the GPU (the parts after the
// renderer): what the reference does per frame with torch index ops, numpy and pytorch3d once `IDHRNetwork.forward` returned.
//
//   arah_frame_images   : masked_scatter_ of rgb / points_cam into H x W images and the finite-difference normal map of the
//                         depth image                                  (im2mesh/metaavatar_render/lightning_model.py:176-205)
//   arah_psnr           : mean squared error + PSNR of two float lists  (:218-221, im2mesh/utils/eval.py:6-9)
//   arah_ssim           : SSIM of the two images inside the mask's bounding rectangle  (:222, im2mesh/utils/eval.py:11-19)
//   arah_rasterize_mesh : nearest face per pixel of a triangle mesh seen by a pytorch3d-convention perspective camera
//                         (MeshRasterizer, faces_per_pixel 1, blur 0: im2mesh/metaavatar_render/models/__init__.py:238-254,
//                         265-277, 291-299; algorithm restated from pytorch3d 0.6.1, see oracle/images_oracle.py)
//   arah_face_normal_image : per-pixel face normal -> colour            (models/__init__.py:256-263, 279-286, 301-308)
//
// All of it is HBM-bound byte / index work on a few MB per frame: one thread per ray / pixel / vertex / face, coalesced
// accesses, the z-buffer is a 64-bit atomicMin on (depth bits, face index) keys — deterministic, lowest face index on ties.
// The arithmetic lives in arah_image_core.h (shared with the host test harness); the kernels here only index.
#ifndef ARAH_CUDA_EMU                      // tests/native/cuda_emu.h runs this file's source on the CPU (test infrastructure)
#include <cuda_runtime.h>
#define ARAH_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif
#include <stdint.h>
#include <string>

#include "../../include/arah_b200.h"
#include "arah_image_core.h"

extern "C" int arah_internal_fail(int code, const char* msg);

namespace arah_img {

constexpr int BLK = 256;
static inline unsigned nblk(size_t n) { return (unsigned)((n + BLK - 1) / BLK); }
static inline size_t ialign(size_t v) { return (v + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------ validation images
// ray k -> pixel pix[k]: rows arrive in row-major mask order (pix ascending), so the 12-byte stores are nearly contiguous
__global__ void __launch_bounds__(BLK) k_img_scatter(const float* __restrict__ rgb, const float* __restrict__ pts, const int32_t* __restrict__ pix,
                                                     int P, int n_pix, float* __restrict__ img_rgb, float* __restrict__ img_pts) {
    const int k = blockIdx.x * BLK + threadIdx.x;
    if (k >= P) return;
    const int p = pix[k];
    if (p < 0 || p >= n_pix) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (img_rgb) img_rgb[(size_t)p * 3 + c] = rgb[(size_t)k * 3 + c];
        if (img_pts) img_pts[(size_t)p * 3 + c] = pts[(size_t)k * 3 + c];
    }
}

__global__ void __launch_bounds__(BLK) k_img_normals(const float* __restrict__ img_pts, int H, int W, float* __restrict__ normals) {
    const size_t i = (size_t)blockIdx.x * BLK + threadIdx.x;
    if (i >= (size_t)H * W) return;
    float o[3];
    depth_normal(img_pts, H, W, (int)(i / W), (int)(i % W), o);
    normals[i * 3 + 0] = o[0]; normals[i * 3 + 1] = o[1]; normals[i * 3 + 2] = o[2];
}

// ------------------------------------------------------------------------------------------------ PSNR
constexpr int PSNR_MAX_BLOCKS = 1184;                  // 148 SMs x 8 resident CTAs

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < BLK / 32 ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;                                          // valid in thread 0
}

// fixed grid-stride partition + fixed reduction trees: the result does not depend on scheduling (bit-reproducible)
__global__ void __launch_bounds__(BLK) k_sqdiff_partial(const float* __restrict__ a, const float* __restrict__ b, long long n, double* __restrict__ partial) {
    __shared__ double sh[BLK / 32];
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * BLK;
    for (long long i = (long long)blockIdx.x * BLK + threadIdx.x; i < n; i += stride) {
        const float d = sub(__ldg(a + i), __ldg(b + i));
        acc += (double)mul(d, d);                      // (pred - gt) ** 2 in fp32, as numpy; the mean accumulates in fp64
    }
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(BLK) k_psnr_finish(const double* __restrict__ partial, int nb, long long n, double* __restrict__ out) {
    __shared__ double sh[BLK / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nb; i += BLK) acc += partial[i];
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) {
        const float mse = (float)(t / (double)n);      // np.mean of a float32 array is a float32
        out[0] = (double)mse;
        out[1] = -10.0 * log((double)mse) / log(10.0);
    }
}

// ------------------------------------------------------------------------------------------------ SSIM
constexpr int SSIM_BLOCKS = 148;                       // 1 CTA per SM; the crop is not known on the host (no synchronisation)
struct SsimScratch { int x0, x1, y0, y1; double partial[SSIM_BLOCKS]; };

__global__ void k_ssim_init(SsimScratch* s) { s->x0 = 0x7fffffff; s->x1 = -1; s->y0 = 0x7fffffff; s->y1 = -1; }

// cv2.boundingRect of the mask: integer atomics, order-independent
__global__ void __launch_bounds__(BLK) k_ssim_rect(const uint8_t* __restrict__ mask, int H, int W, SsimScratch* s) {
    const size_t i = (size_t)blockIdx.x * BLK + threadIdx.x;
    if (i >= (size_t)H * W || mask[i] == 0) return;
    const int y = (int)(i / W), x = (int)(i % W);
    atomicMin(&s->x0, x); atomicMax(&s->x1, x); atomicMin(&s->y0, y); atomicMax(&s->y1, y);
}

// one item = one interior pixel of the crop x one channel; fixed grid-stride partition, fp64, fixed trees: bit-reproducible
__global__ void __launch_bounds__(BLK) k_ssim_partial(const float* __restrict__ X, const float* __restrict__ Y, int W, SsimScratch* s) {
    __shared__ double sh[BLK / 32];
    // empty mask: x0 = INT_MAX, x1 = -1 -- guard before subtracting (signed overflow otherwise), exactly as k_ssim_finish does
    const int x0 = s->x0, y0 = s->y0;
    const int w = (s->x1 >= s->x0) ? s->x1 - s->x0 + 1 : 0, h = (s->y1 >= s->y0) ? s->y1 - s->y0 + 1 : 0;
    const int iw = w - 2 * SSIM_PAD, ih = h - 2 * SSIM_PAD;
    double acc = 0.0;
    if (iw > 0 && ih > 0) {
        const long long items = 3ll * iw * ih, stride = (long long)SSIM_BLOCKS * BLK;
        for (long long i = (long long)blockIdx.x * BLK + threadIdx.x; i < items; i += stride) {
            const int ch = (int)(i % 3);
            const long long p = i / 3;
            acc += ssim_pixel(X, Y, W, y0 + SSIM_PAD + (int)(p / iw), x0 + SSIM_PAD + (int)(p % iw), ch);
        }
    }
    const double t = block_sum(acc, sh);
    if (threadIdx.x == 0) s->partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(BLK) k_ssim_finish(const SsimScratch* s, double* __restrict__ out) {
    __shared__ double sh[BLK / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < SSIM_BLOCKS; i += BLK) acc += s->partial[i];
    const double t = block_sum(acc, sh);
    if (threadIdx.x != 0) return;
    const int w = s->x1 >= s->x0 ? s->x1 - s->x0 + 1 : 0, h = s->y1 >= s->y0 ? s->y1 - s->y0 + 1 : 0;
    const long long n = 3ll * (w - 2 * SSIM_PAD) * (h - 2 * SSIM_PAD);
    out[0] = (w >= SSIM_WIN && h >= SSIM_WIN) ? t / (double)n : nan("");      // skimage raises for a crop smaller than the window
    out[1] = w ? s->x0 : 0; out[2] = h ? s->y0 : 0; out[3] = w; out[4] = h;
}

// ------------------------------------------------------------------------------------------------ rasteriser
__global__ void __launch_bounds__(BLK) k_project(const float* __restrict__ verts, int n, Camera cam, float* __restrict__ ndc) {
    const int v = blockIdx.x * BLK + threadIdx.x;
    if (v >= n) return;
    const float in[3] = {verts[(size_t)v * 3], verts[(size_t)v * 3 + 1], verts[(size_t)v * 3 + 2]};
    float o[3];
    project(cam, in, o);
    ndc[(size_t)v * 3] = o[0]; ndc[(size_t)v * 3 + 1] = o[1]; ndc[(size_t)v * 3 + 2] = o[2];
}

// Load + set up face f; false if it cannot be drawn (bad indices, non-finite / behind the camera / zero area) or misses the image.
__device__ __forceinline__ bool raster_face(const float* __restrict__ ndc, const int32_t* __restrict__ faces, int f, int n_verts, int H, int W,
                                            FaceSetup* s, int* x_lo, int* x_hi, int* y_lo, int* y_hi) {
    const int i0 = faces[(size_t)f * 3], i1 = faces[(size_t)f * 3 + 1], i2 = faces[(size_t)f * 3 + 2];
    if ((unsigned)i0 >= (unsigned)n_verts || (unsigned)i1 >= (unsigned)n_verts || (unsigned)i2 >= (unsigned)n_verts) return false;
    *s = face_setup(ndc + (size_t)i0 * 3, ndc + (size_t)i1 * 3, ndc + (size_t)i2 * 3);
    if (!s->drawable) return false;
    pixel_range(s->xmin, s->xmax, W, H, x_lo, x_hi);
    pixel_range(s->ymin, s->ymax, H, W, y_lo, y_hi);
    return *x_lo <= *x_hi && *y_lo <= *y_hi;
}

// One thread per face: iso-surface triangles of a 256^3 lattice cover a few pixels each at 512 x 512, so a lane sweeps its own
// face's pixel box.  A face whose box exceeds RASTER_COOP_PIXELS is instead swept by the whole warp (lanes stride over the box;
// every lane rebuilds the face's constants from the same global loads), so that a screen-filling triangle costs box/32 tests per
// lane instead of serialising in one thread.  The per-pixel test and the (depth, face) key are the same in both shapes and the
// z-buffer is an order-independent minimum: the result does not depend on which shape a face takes.
constexpr int RASTER_COOP_PIXELS = 256;
__global__ void __launch_bounds__(BLK) k_raster_faces(const float* __restrict__ ndc, const int32_t* __restrict__ faces, int n_faces, int n_verts,
                                                      int H, int W, unsigned long long* __restrict__ keys) {
    const int f = blockIdx.x * BLK + threadIdx.x, lane = threadIdx.x & 31;
    FaceSetup s;
    int x_lo = 0, x_hi = -1, y_lo = 0, y_hi = -1;
    const bool draw = f < n_faces && raster_face(ndc, faces, f, n_verts, H, W, &s, &x_lo, &x_hi, &y_lo, &y_hi);
    const bool big = draw && (long long)(x_hi - x_lo + 1) * (y_hi - y_lo + 1) > RASTER_COOP_PIXELS;
    if (draw && !big)
        for (int y = y_lo; y <= y_hi; ++y) {
            const float py = pix_to_ndc(H - 1 - y, H, W);
            if (py < s.ymin || py > s.ymax) continue;
            for (int x = x_lo; x <= x_hi; ++x) {
                float pz;
                if (face_covers(s, pix_to_ndc(W - 1 - x, W, H), py, &pz)) atomicMin(keys + (size_t)y * W + x, raster_key(pz, f));
            }
        }
    unsigned todo = __ballot_sync(0xffffffffu, big);            // every lane of the warp reaches this point
    while (todo) {
        const int fb = f - lane + (__ffs(todo) - 1);
        todo &= todo - 1;
        FaceSetup sb;
        int bx_lo, bx_hi, by_lo, by_hi;
        if (!raster_face(ndc, faces, fb, n_verts, H, W, &sb, &bx_lo, &bx_hi, &by_lo, &by_hi)) continue;   // uniform across the warp
        const int bw = bx_hi - bx_lo + 1;
        const long long npx = (long long)bw * (by_hi - by_lo + 1);
        for (long long i = lane; i < npx; i += 32) {
            const int y = by_lo + (int)(i / bw), x = bx_lo + (int)(i % bw);
            float pz;
            if (face_covers(sb, pix_to_ndc(W - 1 - x, W, H), pix_to_ndc(H - 1 - y, H, W), &pz)) atomicMin(keys + (size_t)y * W + x, raster_key(pz, fb));
        }
    }
}

__global__ void __launch_bounds__(BLK) k_raster_resolve(const unsigned long long* __restrict__ keys, size_t n, int32_t* __restrict__ pix_to_face,
                                                        float* __restrict__ zbuf) {
    const size_t i = (size_t)blockIdx.x * BLK + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    const bool bg = k == KEY_EMPTY;
    pix_to_face[i] = bg ? -1 : (int32_t)(unsigned)(k & 0xffffffffull);
    if (zbuf) zbuf[i] = bg ? -1.0f : __uint_as_float((unsigned)(k >> 32));
}

struct Rot { float m[9]; int use; };

__global__ void __launch_bounds__(BLK) k_normal_image(const float* __restrict__ verts, const int32_t* __restrict__ faces, int n_faces, int n_verts,
                                                      const int32_t* __restrict__ pix_to_face, size_t n, float sign, Rot rot, float background,
                                                      float* __restrict__ image) {
    const size_t i = (size_t)blockIdx.x * BLK + threadIdx.x;
    if (i >= n) return;
    const int f = pix_to_face[i];
    float o[3];
    o[0] = o[1] = o[2] = to_unit(background);
    if (f >= 0 && f < n_faces) {
        const int i0 = faces[(size_t)f * 3], i1 = faces[(size_t)f * 3 + 1], i2 = faces[(size_t)f * 3 + 2];
        if ((unsigned)i0 < (unsigned)n_verts && (unsigned)i1 < (unsigned)n_verts && (unsigned)i2 < (unsigned)n_verts)
            face_normal_pixel(verts + (size_t)i0 * 3, verts + (size_t)i1 * 3, verts + (size_t)i2 * 3, sign, rot.use ? rot.m : nullptr, o);
    }
    image[i * 3] = o[0]; image[i * 3 + 1] = o[1]; image[i * 3 + 2] = o[2];
}

}  // namespace arah_img

using namespace arah_img;

#define ICU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return arah_internal_fail(ARAH_ECUDA, (std::string(#x) + ": " + cudaGetErrorString(e_)).c_str()); } while (0)

static bool bad_image(int32_t H, int32_t W) { return H <= 0 || W <= 0 || (long long)H * W > (1ll << 28); }

extern "C" size_t arah_frame_images_workspace(int32_t H, int32_t W) {
    return bad_image(H, W) ? 0 : ialign((size_t)H * W * 3 * sizeof(float));
}

extern "C" int arah_frame_images(const float* rgb, const float* points_cam, const int32_t* pix, int32_t P, int32_t H, int32_t W, float* pred_pixels,
                                 float* pred_normals, void* workspace, size_t workspace_bytes, void* stream) {
    if (bad_image(H, W)) return arah_internal_fail(ARAH_EINVAL, "bad image size");
    if (P < 0 || (long long)P > (long long)H * W) return arah_internal_fail(ARAH_EINVAL, "P must be in [0, H*W]");
    if (!pred_pixels && !pred_normals) return arah_internal_fail(ARAH_EINVAL, "no output requested");
    if (P > 0 && (!pix || (pred_pixels && !rgb) || (pred_normals && !points_cam))) return arah_internal_fail(ARAH_EINVAL, "null argument");
    if (pred_normals && (!workspace || workspace_bytes < arah_frame_images_workspace(H, W)))
        return arah_internal_fail(ARAH_EINVAL, "workspace smaller than arah_frame_images_workspace(H, W)");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)H * W;
    float* img_pts = pred_normals ? (float*)workspace : nullptr;
    if (pred_pixels) ICU(cudaMemsetAsync(pred_pixels, 0, n * 3 * sizeof(float), st));
    if (img_pts) ICU(cudaMemsetAsync(img_pts, 0, n * 3 * sizeof(float), st));
    if (P > 0) ARAH_LAUNCH(k_img_scatter, nblk((size_t)P), BLK, st, rgb, points_cam, pix, P, (int)n, pred_pixels, img_pts);
    if (pred_normals) ARAH_LAUNCH(k_img_normals, nblk(n), BLK, st, img_pts, H, W, pred_normals);
    ICU(cudaGetLastError());
    return ARAH_OK;
}

extern "C" size_t arah_psnr_workspace(void) { return ialign(PSNR_MAX_BLOCKS * sizeof(double)); }

extern "C" int arah_psnr(const float* pred, const float* gt, int64_t n, double* mse_psnr, void* workspace, size_t workspace_bytes, void* stream) {
    if (!pred || !gt || !mse_psnr || !workspace) return arah_internal_fail(ARAH_EINVAL, "null argument");
    if (n <= 0) return arah_internal_fail(ARAH_EINVAL, "n <= 0 (the mean of an empty list is undefined)");
    if (workspace_bytes < arah_psnr_workspace()) return arah_internal_fail(ARAH_EINVAL, "workspace smaller than arah_psnr_workspace()");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t want = ((size_t)n + (size_t)BLK * 8 - 1) / ((size_t)BLK * 8);
    const int nb = (int)(want < 1 ? 1 : (want > (size_t)PSNR_MAX_BLOCKS ? (size_t)PSNR_MAX_BLOCKS : want));
    ARAH_LAUNCH(k_sqdiff_partial, nb, BLK, st, pred, gt, (long long)n, (double*)workspace);
    ARAH_LAUNCH(k_psnr_finish, 1, BLK, st, (const double*)workspace, nb, (long long)n, mse_psnr);
    ICU(cudaGetLastError());
    return ARAH_OK;
}

extern "C" size_t arah_ssim_workspace(void) { return ialign(sizeof(SsimScratch)); }

extern "C" int arah_ssim(const float* pred_image, const float* gt_image, const uint8_t* mask, int32_t H, int32_t W, double* out5, void* workspace,
                         size_t workspace_bytes, void* stream) {
    if (!pred_image || !gt_image || !mask || !out5 || !workspace) return arah_internal_fail(ARAH_EINVAL, "null argument");
    if (bad_image(H, W)) return arah_internal_fail(ARAH_EINVAL, "bad image size");
    if (workspace_bytes < arah_ssim_workspace()) return arah_internal_fail(ARAH_EINVAL, "workspace smaller than arah_ssim_workspace()");
    cudaStream_t st = (cudaStream_t)stream;
    SsimScratch* s = (SsimScratch*)workspace;
    ARAH_LAUNCH(k_ssim_init, 1, 1, st, s);
    ARAH_LAUNCH(k_ssim_rect, nblk((size_t)H * W), BLK, st, mask, H, W, s);
    ARAH_LAUNCH(k_ssim_partial, SSIM_BLOCKS, BLK, st, pred_image, gt_image, W, s);
    ARAH_LAUNCH(k_ssim_finish, 1, BLK, st, (const SsimScratch*)s, out5);
    ICU(cudaGetLastError());
    return ARAH_OK;
}

extern "C" size_t arah_rasterize_mesh_workspace(int32_t n_verts, int32_t H, int32_t W) {
    if (n_verts <= 0 || bad_image(H, W)) return 0;
    return ialign((size_t)n_verts * 3 * sizeof(float)) + ialign((size_t)H * W * sizeof(unsigned long long));
}

extern "C" int arah_rasterize_mesh(const float* verts, int32_t n_verts, const int32_t* faces, int32_t n_faces, const ArahRasterCamera* cam, int32_t H,
                                   int32_t W, int32_t* pix_to_face, float* zbuf, void* workspace, size_t workspace_bytes, void* stream) {
    if (!cam || !pix_to_face || !workspace) return arah_internal_fail(ARAH_EINVAL, "null argument");
    if (bad_image(H, W)) return arah_internal_fail(ARAH_EINVAL, "bad image size");
    if (n_verts < 0 || n_faces < 0 || ((n_verts > 0) && !verts) || ((n_faces > 0) && !faces)) return arah_internal_fail(ARAH_EINVAL, "bad mesh");
    if (workspace_bytes < arah_rasterize_mesh_workspace(n_verts > 0 ? n_verts : 1, H, W))
        return arah_internal_fail(ARAH_EINVAL, "workspace smaller than arah_rasterize_mesh_workspace(n_verts, H, W)");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)H * W;
    float* ndc = (float*)workspace;
    unsigned long long* keys = (unsigned long long*)((uint8_t*)workspace + ialign((size_t)(n_verts > 0 ? n_verts : 1) * 3 * sizeof(float)));
    Camera c;
    for (int i = 0; i < 9; ++i) c.R[i] = cam->R[i];
    for (int i = 0; i < 3; ++i) c.T[i] = cam->T[i];
    c.fx = cam->fx; c.fy = cam->fy; c.px = cam->px; c.py = cam->py;
    ICU(cudaMemsetAsync(keys, 0xff, n * sizeof(unsigned long long), st));
    if (n_verts > 0 && n_faces > 0) {
        ARAH_LAUNCH(k_project, nblk((size_t)n_verts), BLK, st, verts, n_verts, c, ndc);
        ARAH_LAUNCH(k_raster_faces, nblk((size_t)n_faces), BLK, st, ndc, faces, n_faces, n_verts, H, W, keys);
    }
    ARAH_LAUNCH(k_raster_resolve, nblk(n), BLK, st, keys, n, pix_to_face, zbuf);
    ICU(cudaGetLastError());
    return ARAH_OK;
}

extern "C" int arah_face_normal_image(const float* verts, int32_t n_verts, const int32_t* faces, int32_t n_faces, const int32_t* pix_to_face, int32_t H,
                                      int32_t W, float sign, const float* rot3x3, float background, float* image, void* stream) {
    if (!pix_to_face || !image) return arah_internal_fail(ARAH_EINVAL, "null argument");
    if (bad_image(H, W)) return arah_internal_fail(ARAH_EINVAL, "bad image size");
    if (n_verts < 0 || n_faces < 0 || ((n_faces > 0) && (!verts || !faces))) return arah_internal_fail(ARAH_EINVAL, "bad mesh");
    Rot r;
    r.use = rot3x3 != nullptr;
    for (int i = 0; i < 9; ++i) r.m[i] = rot3x3 ? rot3x3[i] : 0.0f;
    const size_t n = (size_t)H * W;
    ARAH_LAUNCH(k_normal_image, nblk(n), BLK, (cudaStream_t)stream, verts, faces, n_faces, n_verts, pix_to_face, n, sign, r, background, image);
    ICU(cudaGetLastError());
    return ARAH_OK;
}
