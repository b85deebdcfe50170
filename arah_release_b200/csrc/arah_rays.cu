// arah_rays.cu — per-frame ray set-up on the GPU: what the reference's dataset classes compute with numpy
// and cv2 on the CPU for every frame before the renderer can start (im2mesh/data/zju_mocap_odp.py:250-315).
//
//   arah_pose_smpl  : pose blend shapes (a [3 n_verts][207] GEMV, fp64 accumulation as numpy does with scipy's float64 pose
//                     feature) + linear blend skinning + translation -> posed vertices; bounding box with margin   (:268-289)
//   arah_frame_rays : 2-D mask of the projected bounding box — six cv2.fillPoly calls (utils/utils.py:43-52) restated in integer
//                     arithmetic (16.16 fixed-point scanline spans + clipped 8-connected Bresenham outlines, see
//                     oracle/rays_oracle.py for what is pinned) —, pixel rays uv = [x y 1] K^-T, d = normalise(uv R), slab
//                     intersection with the box (utils/utils.py:54-73), and the ORDERED compaction of the pixels with
//                     near < far (row-major, as np.where + boolean indexing produce them)                         (:291-315)
// HBM-bound byte / index work: ~45 B per ray out, 17 MB of blend-shape basis in per frame.  Floating point is fp32 with one
// rounding per numpy operation (explicit _rn intrinsics; dot products in the k order of an FMA sgemm micro-kernel), so the
// integer outputs (mask, pixel list) match the reference bit for bit except at exact numerical ties.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <string>

#include "../../include/arah_b200.h"

extern "C" int arah_internal_fail(int code, const char* msg);

namespace arah_rays {

constexpr int XY_SHIFT = 16;
constexpr long long XY_ONE = 1ll << XY_SHIFT;
constexpr int MAX_EDGES = 5, NPOLY = 6, NLINES = 30, BLK = 256;

struct Edge { int y0, y1; long long x, dx; };
struct MaskPlan {                      // built by one thread, consumed by the raster kernels
    int corners[8][2];
    int nedge[NPOLY];
    Edge edge[NPOLY][MAX_EDGES];
    int line[NLINES][4];               // clipped end points; line[i][0] = INT_MIN: invisible
};

// ---------------------------------------------------------------------------------------------- posed SMPL (zju_mocap_odp.py:268-289)
__device__ __forceinline__ unsigned enc(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float dec(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_bounds_init(unsigned* mm) { if (threadIdx.x < 3) mm[threadIdx.x] = 0xffffffffu; else if (threadIdx.x < 6) mm[threadIdx.x] = 0u; }

// one warp per vertex: three 207-long fp64 dot products (coalesced rows of the blend-shape basis), then LBS by lane 0
__global__ void __launch_bounds__(256) k_pose_smpl(const float* __restrict__ shape, const float* __restrict__ posedirs, const double* __restrict__ pf,
                                                   const float* __restrict__ w, const float* __restrict__ B, float tx, float ty, float tz, int n,
                                                   float* __restrict__ verts, unsigned* __restrict__ mm) {
    __shared__ double spf[207];
    __shared__ float sB[24 * 16];
    for (int i = threadIdx.x; i < 207; i += blockDim.x) spf[i] = pf[i];
    for (int i = threadIdx.x; i < 24 * 16; i += blockDim.x) sB[i] = B[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (v >= n) return;
    double acc[3] = {0.0, 0.0, 0.0};
    for (int k = lane; k < 207; k += 32) {
        const double p = spf[k];
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = fma((double)__ldcs(posedirs + (size_t)(3 * v + c) * 207 + k), p, acc[c]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (lane) return;
    float ms[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) ms[c] = (float)((double)shape[3 * v + c] + acc[c]);          // float32 += float64 (:272)
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
    for (int j = 0; j < 24; ++j) {                                                            // T = W . B (:276), k-sequential FMA
        const float wj = w[(size_t)v * 24 + j];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = fmaf(wj, sB[16 * j + e], T[e]);
    }
    const float t3[3] = {tx, ty, tz};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float s = T[4 * r] * ms[0];
        s = fmaf(T[4 * r + 1], ms[1], s); s = fmaf(T[4 * r + 2], ms[2], s); s = fmaf(T[4 * r + 3], 1.0f, s);
        const float o = __fadd_rn(s, t3[r]);                                                  // + trans (:280)
        verts[3 * (size_t)v + r] = o;
        atomicMin(&mm[r], enc(o)); atomicMax(&mm[3 + r], enc(o));
    }
}
__global__ void k_bounds_finish(const unsigned* mm, float margin, float* bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = __fsub_rn(dec(mm[threadIdx.x]), margin);       // :288-289
    else if (threadIdx.x < 6) bounds[threadIdx.x] = __fadd_rn(dec(mm[threadIdx.x]), margin);
}

// ---------------------------------------------------------------------------------------------- cv2.fillPoly, restated
__device__ __forceinline__ long long trunc_d(double v) { return (long long)v; }
// cv::clipLine: end points are modified even when the segment is invisible
__device__ bool clip_line(int W, int H, long long& x1, long long& y1, long long& x2, long long& y2) {
    const long long right = W - 1, bottom = H - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) { a = c1 < 8 ? 0 : bottom; x1 += trunc_d((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1)); y1 = a; c1 = (x1 < 0) + (x1 > right) * 2; }
        if (c2 & 12) { a = c2 < 8 ? 0 : bottom; x2 += trunc_d((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1)); y2 = a; c2 = (x2 < 0) + (x2 > right) * 2; }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) { a = c1 == 1 ? 0 : right; y1 += trunc_d((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1)); x1 = a; c1 = 0; }
            if (c2) { a = c2 == 1 ? 0 : right; y2 += trunc_d((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1)); x2 = a; c2 = 0; }
        }
    }
    return (c1 | c2) == 0;
}

struct RayArgs { float K[9], Kinv[9], R[9], T[3], cam[3]; int H, W; };

// one thread: corners (utils/utils.py:17-45), the six index lists (:47-52), CollectPolyEdges for each
__global__ void k_mask_setup(RayArgs a, const float* __restrict__ bounds, MaskPlan* plan) {
    if (threadIdx.x || blockIdx.x) return;
    const float mn[3] = {bounds[0], bounds[1], bounds[2]}, mx[3] = {bounds[3], bounds[4], bounds[5]};
    int c2[8][2];
    for (int c = 0; c < 8; ++c) {
        const float p[3] = {(c & 4) ? mx[0] : mn[0], (c & 2) ? mx[1] : mn[1], (c & 1) ? mx[2] : mn[2]};       // get_bound_corners order
        float cam[3], q[3];
        for (int r = 0; r < 3; ++r)                       // xyz . RT[:, :3]^T + RT[:, 3]  (sgemm k order, then one add)
            cam[r] = __fadd_rn(fmaf(p[2], a.R[3 * r + 2], fmaf(p[1], a.R[3 * r + 1], __fmul_rn(p[0], a.R[3 * r]))), a.T[r]);
        for (int r = 0; r < 3; ++r) q[r] = fmaf(cam[2], a.K[3 * r + 2], fmaf(cam[1], a.K[3 * r + 1], __fmul_rn(cam[0], a.K[3 * r])));
        c2[c][0] = (int)rintf(__fdiv_rn(q[0], q[2]));     // np.round: half to even
        c2[c][1] = (int)rintf(__fdiv_rn(q[1], q[2]));
        plan->corners[c][0] = c2[c][0]; plan->corners[c][1] = c2[c][1];
    }
    const int faces[NPOLY][5] = {{0, 1, 3, 2, 0}, {4, 5, 7, 6, 5}, {0, 1, 5, 4, 0}, {2, 3, 7, 6, 2}, {0, 2, 6, 4, 0}, {1, 3, 7, 5, 1}};
    int nl = 0;
    for (int f = 0; f < NPOLY; ++f) {
        int ne = 0;
        int p0 = faces[f][4];
        for (int i = 0; i < 5; ++i) {
            const int p1 = faces[f][i];
            const long long x0 = c2[p0][0], y0 = c2[p0][1], x1 = c2[p1][0], y1 = c2[p1][1];
            // outline segment (cv::Line clips first)
            long long lx0 = x0, ly0 = y0, lx1 = x1, ly1 = y1;
            const bool vis = clip_line(a.W, a.H, lx0, ly0, lx1, ly1);
            plan->line[nl][0] = vis ? (int)lx0 : INT_MIN; plan->line[nl][1] = (int)ly0; plan->line[nl][2] = (int)lx1; plan->line[nl][3] = (int)ly1;
            ++nl;
            // scanline edge; partially visible edges take slope and anchor from their clipped end points
            long long c0x = x0 << XY_SHIFT, c0y = y0, c1x = x1 << XY_SHIFT, c1y = y1;
            if (!(x0 >= 0 && x0 < a.W && x1 >= 0 && x1 < a.W && y0 >= 0 && y0 < a.H && y1 >= 0 && y1 < a.H)) {
                if (ly0 != ly1) { c0y = ly0; c1y = ly1; c0x = lx0 << XY_SHIFT; c1x = lx1 << XY_SHIFT; }
            }
            if (y0 != y1) {
                Edge e;
                e.dx = (c1x - c0x) / (c1y - c0y);         // C++ division: towards zero
                if (y0 < y1) { e.y0 = (int)y0; e.y1 = (int)y1; e.x = c0x + (y0 - c0y) * e.dx; }
                else { e.y0 = (int)y1; e.y1 = (int)y0; e.x = c1x + (y1 - c1y) * e.dx; }
                plan->edge[f][ne++] = e;
            }
            p0 = p1;
        }
        plan->nedge[f] = ne;
    }
}
// one thread per outline segment: 8-connected Bresenham from the left end point (cv::LineIterator, leftToRight)
__global__ void k_mask_lines(const MaskPlan* __restrict__ plan, int W, uint8_t* __restrict__ mask) {
    const int i = threadIdx.x;
    if (i >= NLINES || plan->line[i][0] == INT_MIN) return;
    int x0 = plan->line[i][0], y0 = plan->line[i][1], x1 = plan->line[i][2], y1 = plan->line[i][3];
    int dx = x1 - x0, dy = y1 - y0;
    if (dx < 0) { int t = x0; x0 = x1; x1 = t; t = y0; y0 = y1; y1 = t; dx = -dx; dy = -dy; }
    const int adx = dx, ady = dy < 0 ? -dy : dy, sy = dy >= 0 ? 1 : -1;
    int x = x0, y = y0;
    if (adx >= ady) {
        int err = adx - 2 * ady;
        for (int k = 0; k <= adx; ++k) { mask[(size_t)y * W + x] = 1; if (err < 0) { err += 2 * adx - 2 * ady; y += sy; } else err -= 2 * ady; ++x; }
    } else {
        int err = ady - 2 * adx;
        for (int k = 0; k <= ady; ++k) { mask[(size_t)y * W + x] = 1; if (err < 0) { err += 2 * ady - 2 * adx; ++x; } else err -= 2 * adx; y += sy; }
    }
}
// grid (H, 6): FillEdgeCollection for one polygon and one row: even-odd pairs of the sorted active edges, ceil .. floor
__global__ void __launch_bounds__(128) k_mask_spans(const MaskPlan* __restrict__ plan, int W, int H, uint8_t* __restrict__ mask) {
    const int y = blockIdx.x, f = blockIdx.y;
    __shared__ int span[4];
    if (threadIdx.x == 0) {
        long long xs[MAX_EDGES];
        int m = 0, ymax = INT_MIN;
        const int ne = plan->nedge[f];
        for (int i = 0; i < ne; ++i) ymax = max(ymax, plan->edge[f][i].y1);
        for (int i = 0; i < ne; ++i) { const Edge e = plan->edge[f][i]; if (e.y0 <= y && y < e.y1) xs[m++] = e.x + (long long)(y - e.y0) * e.dx; }
        for (int i = 1; i < m; ++i) { const long long k = xs[i]; int j = i - 1; while (j >= 0 && xs[j] > k) { xs[j + 1] = xs[j]; --j; } xs[j + 1] = k; }
        span[0] = span[2] = 1; span[1] = span[3] = 0;
        if (y < min(ymax, H))
            for (int k = 0; k + 1 < m && k < 4; k += 2) {
                long long xa = (xs[k] + XY_ONE - 1) >> XY_SHIFT, xb = xs[k + 1] >> XY_SHIFT;
                if (xa < W && xb >= 0) { span[k] = (int)max(xa, 0ll); span[k + 1] = (int)min(xb, (long long)W - 1); }
            }
    }
    __syncthreads();
    for (int s = 0; s < 4; s += 2)
        for (int x = span[s] + threadIdx.x; x <= span[s + 1]; x += blockDim.x) mask[(size_t)y * W + x] = 1;
}

// ---------------------------------------------------------------------------------------------- rays (zju_mocap_odp.py:295-315)
struct RayOut { float d[3], near_, far_; bool hit; };
__device__ __forceinline__ RayOut pixel_ray(const RayArgs& a, const float* __restrict__ bounds, int px, int py) {
    const float x = (float)px, y = (float)py;
    float uv[3], r[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)                            // [x y 1] . K_inv^T  (:298)
        uv[j] = fmaf(1.0f, a.Kinv[3 * j + 2], fmaf(y, a.Kinv[3 * j + 1], __fmul_rn(x, a.Kinv[3 * j])));
#pragma unroll
    for (int j = 0; j < 3; ++j)                            // uv . R  (:176)
        r[j] = fmaf(uv[2], a.R[6 + j], fmaf(uv[1], a.R[3 + j], __fmul_rn(uv[0], a.R[j])));
    const float n0 = __fadd_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[0], r[0]), __fmul_rn(r[1], r[1])), __fmul_rn(r[2], r[2]))), 1e-12f);
    RayOut o;
#pragma unroll
    for (int j = 0; j < 3; ++j) o.d[j] = __fdiv_rn(r[j], n0);                                   // normalize_vectors (:165-169)
    // get_near_far (utils/utils.py:54-73)
    const float nd = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(o.d[0], o.d[0]), __fmul_rn(o.d[1], o.d[1])), __fmul_rn(o.d[2], o.d[2])));
    float nr = -INFINITY, fr = INFINITY;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float v = __fdiv_rn(o.d[j], nd);
        if (v < 1e-5f && v > -1e-10f) v = 1e-5f;
        if (v > -1e-5f && v < 1e-10f) v = -1e-5f;
        const float t0 = __fdiv_rn(__fsub_rn(bounds[j], a.cam[j]), v), t1 = __fdiv_rn(__fsub_rn(bounds[3 + j], a.cam[j]), v);
        nr = fmaxf(nr, fminf(t0, t1)); fr = fminf(fr, fmaxf(t0, t1));
    }
    o.hit = nr < fr;
    o.near_ = __fdiv_rn(nr, nd); o.far_ = __fdiv_rn(fr, nd);
    return o;
}
__device__ __forceinline__ unsigned block_scan(unsigned v, unsigned* sh, unsigned& total) {   // exclusive, BLK threads
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    unsigned base = 0; total = 0;
#pragma unroll
    for (int w = 0; w < BLK / 32; ++w) { if (w < warp) base += sh[w]; total += sh[w]; }
    __syncthreads();
    return base + inc - v;
}
__global__ void __launch_bounds__(BLK) k_rays_count(RayArgs a, const float* __restrict__ bounds, const uint8_t* __restrict__ mask,
                                                    uint8_t* __restrict__ image_mask, unsigned* __restrict__ block_sums) {
    __shared__ unsigned sh[BLK / 32];
    const int p = blockIdx.x * BLK + threadIdx.x, n = a.H * a.W;
    unsigned hit = 0;
    if (p < n) {
        if (mask[p]) hit = pixel_ray(a, bounds, p % a.W, p / a.W).hit ? 1u : 0u;
        image_mask[p] = (uint8_t)hit;
    }
    unsigned tot;
    (void)block_scan(hit, sh, tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) k_rays_scan(unsigned* __restrict__ block_sums, int nblocks, int32_t* __restrict__ count) {
    __shared__ unsigned s[1024];
    const int t = threadIdx.x, per = (nblocks + 1023) / 1024, b0 = t * per, b1 = min(nblocks, b0 + per);
    unsigned a = 0;
    for (int i = b0; i < b1; ++i) a += block_sums[i];
    s[t] = a;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { unsigned x = 0; if (t >= o) x = s[t - o]; __syncthreads(); s[t] += x; __syncthreads(); }
    unsigned r = s[t] - a;
    for (int i = b0; i < b1; ++i) { const unsigned v = block_sums[i]; block_sums[i] = r; r += v; }
    if (t == 1023) count[0] = (int32_t)s[t];
}
__global__ void __launch_bounds__(BLK) k_rays_emit(RayArgs a, const float* __restrict__ bounds, const uint8_t* __restrict__ image_mask,
                                                   const unsigned* __restrict__ block_offs, int32_t* __restrict__ pix,
                                                   float* __restrict__ dirs, float* __restrict__ near_far) {
    __shared__ unsigned sh[BLK / 32];
    const int p = blockIdx.x * BLK + threadIdx.x, n = a.H * a.W;
    const unsigned hit = (p < n && image_mask[p]) ? 1u : 0u;
    unsigned tot;
    const unsigned off = block_offs[blockIdx.x] + block_scan(hit, sh, tot);
    if (!hit) return;
    const RayOut o = pixel_ray(a, bounds, p % a.W, p / a.W);
    pix[off] = p;
    dirs[3 * (size_t)off] = o.d[0]; dirs[3 * (size_t)off + 1] = o.d[1]; dirs[3 * (size_t)off + 2] = o.d[2];
    near_far[2 * (size_t)off] = o.near_; near_far[2 * (size_t)off + 1] = o.far_;
}

}  // namespace arah_rays

using namespace arah_rays;

#define RCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return arah_internal_fail(ARAH_ECUDA, (std::string(#x) + ": " + cudaGetErrorString(e_)).c_str()); } while (0)

extern "C" int arah_pose_smpl(const float* minimal_shape, const float* posedirs, const double* pose_feature, const float* skinning_weights,
                              const float* bone_transforms, const float* trans3, int32_t n_verts, float box_margin, float* verts,
                              float* bounds, void* workspace, void* stream) {
    if (!minimal_shape || !posedirs || !pose_feature || !skinning_weights || !bone_transforms || !trans3 || !verts || !bounds || !workspace)
        return arah_internal_fail(ARAH_EINVAL, "null argument");
    if (n_verts <= 0) return arah_internal_fail(ARAH_EINVAL, "n_verts <= 0");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned* mm = (unsigned*)workspace;
    k_bounds_init<<<1, 32, 0, st>>>(mm);
    k_pose_smpl<<<(unsigned)(((size_t)n_verts * 32 + 255) / 256), 256, 0, st>>>(minimal_shape, posedirs, pose_feature, skinning_weights, bone_transforms,
                                                                                 trans3[0], trans3[1], trans3[2], n_verts, verts, mm);
    k_bounds_finish<<<1, 32, 0, st>>>(mm, box_margin, bounds);
    RCU(cudaGetLastError());
    return ARAH_OK;
}

static inline size_t ralign(size_t v) { return (v + 255) & ~(size_t)255; }
extern "C" size_t arah_frame_rays_workspace(int32_t H, int32_t W) {
    if (H <= 0 || W <= 0) return 0;
    const size_t n = (size_t)H * W, nb = (n + BLK - 1) / BLK;
    return ralign(sizeof(MaskPlan)) + ralign(nb * 4) + 256;
}

extern "C" int arah_frame_rays(const float* K, const float* K_inv, const float* R, const float* T, const float* cam_loc, const float* bounds,
                               int32_t H, int32_t W, const uint8_t* mask_in, uint8_t* bound_mask, int32_t* pix, float* ray_dirs, float* near_far,
                               uint8_t* image_mask, int32_t* count, void* workspace, size_t workspace_bytes, void* stream) {
    if (!K || !K_inv || !R || !T || !cam_loc || !bounds || !pix || !ray_dirs || !near_far || !image_mask || !count || !workspace)
        return arah_internal_fail(ARAH_EINVAL, "null argument");
    if (!mask_in && !bound_mask) return arah_internal_fail(ARAH_EINVAL, "either a mask or a buffer for the bounding-box mask is required");
    if (H <= 0 || W <= 0 || (long long)H * W > (1ll << 30)) return arah_internal_fail(ARAH_EINVAL, "bad image size");
    if (workspace_bytes < arah_frame_rays_workspace(H, W)) return arah_internal_fail(ARAH_EINVAL, "workspace smaller than arah_frame_rays_workspace(H, W)");
    cudaStream_t st = (cudaStream_t)stream;
    RayArgs a;
    for (int i = 0; i < 9; ++i) { a.K[i] = K[i]; a.Kinv[i] = K_inv[i]; a.R[i] = R[i]; }
    for (int i = 0; i < 3; ++i) { a.T[i] = T[i]; a.cam[i] = cam_loc[i]; }
    a.H = H; a.W = W;
    MaskPlan* plan = (MaskPlan*)workspace;
    unsigned* bs = (unsigned*)((uint8_t*)workspace + ralign(sizeof(MaskPlan)));
    const size_t n = (size_t)H * W;
    const unsigned nb = (unsigned)((n + BLK - 1) / BLK);
    const uint8_t* mask = mask_in;
    if (!mask_in) {
        RCU(cudaMemsetAsync(bound_mask, 0, n, st));
        k_mask_setup<<<1, 32, 0, st>>>(a, bounds, plan);
        k_mask_lines<<<1, 32, 0, st>>>(plan, W, bound_mask);
        k_mask_spans<<<dim3((unsigned)H, NPOLY), 128, 0, st>>>(plan, W, H, bound_mask);
        mask = bound_mask;
    }
    k_rays_count<<<nb, BLK, 0, st>>>(a, bounds, mask, image_mask, bs);
    k_rays_scan<<<1, 1024, 0, st>>>(bs, (int)nb, count);
    k_rays_emit<<<nb, BLK, 0, st>>>(a, bounds, image_mask, bs, pix, ray_dirs, near_far);
    RCU(cudaGetLastError());
    return ARAH_OK;
}
