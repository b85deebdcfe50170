// arah_kernels.cuh — the stage kernels of the ARAH hot path (sm_100a).
//
// Stage map (reference file:line each kernel replaces; all under /root/reference/im2mesh):
//   k_trace_begin        ray_tracing.py:178-196            per-ray init, active list
//   k_knn_rays/_samples  ray_tracing.py:382-400, 403-421   brute-force 1-NN over the posed SMPL verts + NN-skinning inverse
//   k_trace_iter         ray_tracing.py:198-241            one sphere-tracing step for the still-active rays
//   k_iso_prepare/_init  utils/root_finding_utils.py:365-418  joint-search Jacobian (full LBS jac + grad sdf), J^-1, g(u0)
//   k_iso_iter           utils/root_finding_utils.py:426-461 + utils/broyden.py:47-76   4-D Broyden step
//   k_trace_finish       ray_tracing.py:266-294, 313-350   accept/reject, z-sample placement (merge of 2 sorted runs)
//   k_corr_init/_iter    utils/root_finding_utils.py:267-362 + utils/broyden.py   3-D Broyden per sample
//   k_shade              renderer/implicit_differentiable_renderer.py:284-361    SDF fwd + grad, colour MLP
//   k_composite          renderer/implicit_differentiable_renderer.py:366-394, 225-257
// Iterations are separate launches over a compacted active list that lives in HBM (the reference compacts with
// boolean masks too); no host synchronisation anywhere — list sizes are read from device counters.
#pragma once
#include "arah_work.cuh"

namespace arah {

// ================================================================================================ tracing
__global__ void k_trace_begin(Work w) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    bool unfinished = false;
    if (r < w.P) {
        const float nr = w.near_far[2 * r], fr = w.near_far[2 * r + 1];
        unfinished = nr < fr;
        w.ray_t[r] = nr;
        w.ray_flags[r] = unfinished ? 1 : 2;
        RayCur c;
        c.xn[0] = c.xn[1] = c.xn[2] = 0.0f; c.s = 0.0f;
#pragma unroll
        for (int e = 0; e < 12; ++e) c.T[e] = 0.0f;
        w.ray_cur[r] = c;
    }
    warp_append(unfinished, r, w.listA, &w.counters[C_TRACE]);
}

// single block, 1024 threads, dynamic smem = 8192 * 8 bytes; n <= 8192
__global__ void __launch_bounds__(1024) k_knn_build(const float* __restrict__ v3, int n, float4* __restrict__ sv, float4* __restrict__ cmin, float4* __restrict__ cmax) {
    extern __shared__ uint32_t sk[];
    uint32_t* keys = sk;
    uint32_t* idx = sk + 8192;
    __shared__ float red[6][32];
    __shared__ float bb[6];
    const int tid = threadIdx.x;
    float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = tid; i < n; i += blockDim.x)
        for (int k = 0; k < 3; ++k) { const float c = v3[3 * i + k]; lo[k] = fminf(lo[k], c); hi[k] = fmaxf(hi[k], c); }
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) { lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o)); }
        if ((tid & 31) == 0) { red[k][tid >> 5] = lo[k]; red[3 + k][tid >> 5] = hi[k]; }
    }
    __syncthreads();
    if (tid < 6) {
        float a = red[tid][0];
        for (int i = 1; i < 32; ++i) a = (tid < 3) ? fminf(a, red[tid][i]) : fmaxf(a, red[tid][i]);
        bb[tid] = a;
    }
    __syncthreads();
    for (int i = tid; i < 8192; i += blockDim.x) {
        uint32_t key = 0xFFFFFFFFu;
        if (i < n) {
            uint32_t q[3];
            for (int k = 0; k < 3; ++k) {
                const float t = (v3[3 * i + k] - bb[k]) / fmaxf(bb[3 + k] - bb[k], 1e-12f);
                q[k] = (uint32_t)fminf(fmaxf(t * 1023.0f, 0.0f), 1023.0f);
            }
            key = morton_spread10(q[0]) | (morton_spread10(q[1]) << 1) | (morton_spread10(q[2]) << 2);
        }
        keys[i] = key; idx[i] = (uint32_t)i;
    }
    __syncthreads();
    for (int k = 2; k <= 8192; k <<= 1) {                       // bitonic sort of (key, idx), ties broken by idx
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < 8192; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const bool up = (i & k) == 0;
                    const uint32_t ka = keys[i], kb = keys[p], ia = idx[i], ib = idx[p];
                    const bool gt = (ka > kb) || (ka == kb && ia > ib);
                    if (gt == up) { keys[i] = kb; keys[p] = ka; idx[i] = ib; idx[p] = ia; }
                }
            }
            __syncthreads();
        }
    }
    const int nc = (n + KNN_CLUSTER - 1) / KNN_CLUSTER;
    for (int i = tid; i < nc * KNN_CLUSTER; i += blockDim.x) {
        float4 o = make_float4(1e30f, 1e30f, 1e30f, __int_as_float(0x7fffffff));
        if (i < n) { const int s_ = (int)idx[i]; o = make_float4(v3[3 * s_], v3[3 * s_ + 1], v3[3 * s_ + 2], __int_as_float(s_)); }
        sv[i] = o;
    }
    __syncthreads();
    for (int c = tid; c < nc; c += blockDim.x) {
        float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
        for (int i = c * KNN_CLUSTER; i < min(n, (c + 1) * KNN_CLUSTER); ++i) {
            const int s_ = (int)idx[i];
            for (int k = 0; k < 3; ++k) { const float v = v3[3 * s_ + k]; mn[k] = fminf(mn[k], v); mx[k] = fmaxf(mx[k], v); }
        }
        cmin[c] = make_float4(mn[0], mn[1], mn[2], 0.f);
        cmax[c] = make_float4(mx[0], mx[1], mx[2], 0.f);
    }
}
__global__ void __launch_bounds__(512, 1) k_knn_rays(FrameParams fp, KnnIndex ix, Work w, int iter) {
    extern __shared__ float4 sv[];
    const int n = w.counters[C_TRACE + iter];
    const int B = knn_batch_size<KNN_COOP_RAYS>(n);
    if ((int)(blockIdx.x * (blockDim.x >> 5)) * B >= n) return;
    const KnnSmem kk = load_knn(sv, ix);
    const int* list = (iter & 1) ? w.listB : w.listA;
    knn_warp_batches<KNN_COOP_RAYS>(kk, n, B,
        [&](int i, float* x) {
            const int r = list[i];
            const float t = w.ray_t[r];
#pragma unroll
            for (int k = 0; k < 3; ++k) x[k] = w.ray_dirs[3 * r + k] * t + fp.cam_loc[k];
        },
        [&](int i, const float* x, int idx) {
            const int r = list[i];
            RayCur c;
            float xh[3];
            nn_inverse_skinning(fp, idx, x, c.T, &c.s, xh);
            normalize3(fp, xh, c.xn);
            w.ray_cur[r] = c;
        });
}

// unit-level entry (tests): nearest posed-vertex index of n arbitrary points
__global__ void __launch_bounds__(512, 1) k_knn_points(KnnIndex ix, const float* __restrict__ pts, int n, int* __restrict__ out_idx) {
    extern __shared__ float4 sv[];
    const int B = knn_batch_size<KNN_COOP_SAMPLES>(n);
    if ((int)(blockIdx.x * (blockDim.x >> 5)) * B >= n) return;
    const KnnSmem kk = load_knn(sv, ix);
    knn_warp_batches<KNN_COOP_SAMPLES>(kk, n, B,
        [&](int i, float* x) { x[0] = pts[3 * i]; x[1] = pts[3 * i + 1]; x[2] = pts[3 * i + 2]; },
        [&](int i, const float*, int idx) { out_idx[i] = idx; });
}

// smem layout helpers -------------------------------------------------------------------------------------------
struct TileSmem {
    float* A; float* wbuf; float (*xs)[4]; float (*logits)[32]; float* sdfo; uint64_t* bars;
};
constexpr int LDA_SDF = 260;      // 256 + 4
constexpr int LDA_SHADE = 308;    // 304 + 4
__host__ __device__ constexpr size_t tile_smem_bytes(int lda) {
    return (size_t)(TM * lda + WBUF_FLOATS + TM * 4 + TM * 32 + TM) * 4 + 64;
}
__device__ __forceinline__ TileSmem carve(float* base, int lda) {
    TileSmem s;
    s.wbuf = base;                               // 128-byte aligned start (dynamic smem base)
    s.A = base + WBUF_FLOATS;
    float* p = s.A + TM * lda;
    s.xs = reinterpret_cast<float (*)[4]>(p); p += TM * 4;
    s.logits = reinterpret_cast<float (*)[32]>(p); p += TM * 32;
    s.sdfo = p; p += TM;
    s.bars = reinterpret_cast<uint64_t*>(p);
    return s;
}

__global__ void __launch_bounds__(256, 2) k_trace_iter(FrameParams fp, Work w, int iter) {
    extern __shared__ __align__(128) float smem[];
    const int n = w.counters[C_TRACE + iter];
    if ((int)blockIdx.x * TM >= n) return;
    TileSmem s = carve(smem, LDA_SDF);
    WPipe wp;
    wpipe_init(wp, s.wbuf, s.bars);
    const int* list = (iter & 1) ? w.listB : w.listA;
    int* next = (iter & 1) ? w.listA : w.listB;
    const int tid = threadIdx.x;
    for (int tile = blockIdx.x; tile * TM < n; tile += gridDim.x) {
        int r = -1;
        if (tid < TM) {
            const int i = tile * TM + tid;
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) { r = list[i]; const RayCur& c = w.ray_cur[r]; xn[0] = c.xn[0]; xn[1] = c.xn[1]; xn[2] = c.xn[2]; }
            s.xs[tid][0] = xn[0]; s.xs[tid][1] = xn[1]; s.xs[tid][2] = xn[2]; s.xs[tid][3] = 0.f;
        }
        __syncthreads();
        sdf_tile_forward<PLAIN>(fp, s.xs, s.A, LDA_SDF, wp, s.sdfo, nullptr, 0.f);
        __syncthreads();
        if (tid < TM) {       // warps 0,1: marching logic (ray_tracing.py:228-241)
            bool still = false;
            if (r >= 0) {
                const float sdf = sdf_to_metres(s.sdfo[tid], fp.cmin, fp.cmax);
                float t = w.ray_t[r];
                const float far_ = w.near_far[2 * r + 1];
                const float sm = fminf(fmaxf(sdf, -0.1f), 0.1f);
                bool diverge = false;
                if (fabsf(sm) > CVG_THRESH && fabsf(sdf) < 1e6f) { t = t + sm; diverge = t >= far_; w.ray_t[r] = t; }
                still = !(fabsf(sdf) <= CVG_THRESH || diverge);
                w.ray_flags[r] = (still ? 1 : 0) | (diverge ? 2 : 0);
            }
            if (iter + 1 < TRACE_ITERS) warp_append(still, r, next, &w.counters[C_TRACE + iter + 1]);
            warp_stat_add(r >= 0 ? 1 : 0, &w.counters[C_STAT_TRACE_EVALS]);
        }
        __syncthreads();
    }
}

// ================================================================================================ joint iso search
__global__ void k_iso_prepare(Work w) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool go = (r < w.P) && (w.train || !(w.ray_flags[r] & 2));   // eval: non-diverged rays; training: all (ray_tracing.py:249)
    warp_append(go, r, w.listA, &w.counters[C_ISO]);
}

// 16 rays per tile (value + 3 tangent rows each): full LBS Jacobian, grad sdf, 4x4 inverse, g(u0)
__global__ void __launch_bounds__(256, 1) k_iso_init(FrameParams fp, Work w) {
    extern __shared__ __align__(128) float smem[];
    const int n = w.counters[C_ISO];
    constexpr int PTS = TM / 4;
    if ((int)blockIdx.x * PTS >= n) return;
    TileSmem s = carve(smem, LDA_SDF);
    WPipe wp;
    wpipe_init(wp, s.wbuf, s.bars);
    const int tid = threadIdx.x;
    const float dn = 2.0f / (fp.cmax - fp.cmin) / 1.1f;
    for (int tile = blockIdx.x; tile * PTS < n; tile += gridDim.x) {
        int r = -1;
        float x0[3] = {0.f, 0.f, 0.f};
        if (tid < PTS) {
            const int i = tile * PTS + tid;
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) {
                r = w.listA[i];
                const RayCur& c = w.ray_cur[r];
                unnormalize3(fp, c.xn, x0);              // ray_tracing.py:245
                normalize3(fp, x0, xn);                  // root_finding_utils.py:75 (inside query_weights)
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { s.xs[4 * tid + q][0] = xn[0]; s.xs[4 * tid + q][1] = xn[1]; s.xs[4 * tid + q][2] = xn[2]; s.xs[4 * tid + q][3] = 0.f; }
        }
        __syncthreads();
        skin_tile_forward<DUAL>(fp, s.xs, s.A, LDA_SDF, wp, s.logits, dn);
        __syncthreads();
        sdf_tile_forward<DUAL>(fp, s.xs, s.A, LDA_SDF, wp, s.sdfo, nullptr, dn);
        __syncthreads();
        if (r >= 0) {
            Dual3 lx[25], pw[NJ];
#pragma unroll
            for (int c = 0; c < 25; ++c) {
                lx[c].v = s.logits[4 * tid][c] * 20.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) lx[c].d[k] = s.logits[4 * tid + 1 + k][c] * 20.0f;
            }
            hierarchical_softmax_dual(lx, pw);
            float J[16], T12[12], xb[3];
#pragma unroll
            for (int e = 0; e < 16; ++e) J[e] = 0.0f;
#pragma unroll
            for (int e = 0; e < 12; ++e) T12[e] = 0.0f;
            for (int j = 0; j < NJ; ++j) {
                const float* B = fp.bone_T + j * 16;
                float bx[3];
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) bx[rr] = B[rr * 4] * x0[0] + B[rr * 4 + 1] * x0[1] + B[rr * 4 + 2] * x0[2] + B[rr * 4 + 3];
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) J[(rr + 1) * 4 + c] += pw[j].v * B[rr * 4 + c] + bx[rr] * pw[j].d[c];
#pragma unroll
                    for (int c = 0; c < 4; ++c) T12[rr * 4 + c] += pw[j].v * B[rr * 4 + c];
                }
            }
            apply_T(T12, x0, xb);
            const float so = 1.0f / 2.0f * 1.1f * (fp.cmax - fp.cmin);        // d(sdf metres)/d(sdf raw)
#pragma unroll
            for (int c = 0; c < 3; ++c) J[c] = s.sdfo[4 * tid + 1 + c] * so;
            J[3] = 0.0f;
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) J[(rr + 1) * 4 + 3] = -w.ray_dirs[3 * r + rr];
            float Ji[16];
            invert_gj<4>(J, Ji);
            const float z0 = w.ray_t[r];
            const float u0[4] = {x0[0], x0[1], x0[2], z0};
            float g0[4];
#pragma unroll
            for (int k = 0; k < 3; ++k) g0[1 + k] = xb[k] - ((w.ray_dirs[3 * r + k] * z0 + fp.cam_loc[k]) - fp.trans[k]);
            g0[0] = sdf_to_metres(s.sdfo[4 * tid], fp.cmin, fp.cmax);
            BroydenState<4> st;
            broyden_begin<4>(st, u0, g0, Ji, w.ray_cur[r].T);
            st.owner = r;
            st.tgt[0] = st.tgt[1] = st.tgt[2] = 0.f;
            state_store(&w.iso_state[r], st);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256, 2) k_iso_iter(FrameParams fp, Work w, int iter) {
    extern __shared__ __align__(128) float smem[];
    const int n = w.counters[C_ISO + iter];
    if ((int)blockIdx.x * TM >= n) return;
    TileSmem s = carve(smem, LDA_SDF);
    WPipe wp;
    wpipe_init(wp, s.wbuf, s.bars);
    const int* list = (iter & 1) ? w.listB : w.listA;
    int* next = (iter & 1) ? w.listA : w.listB;
    const int tid = threadIdx.x;
    for (int tile = blockIdx.x; tile * TM < n; tile += gridDim.x) {
        int r = -1;
        BroydenState<4> st;
        float dx[4];
        if (tid < TM) {
            const int i = tile * TM + tid;
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) {
                r = list[i];
                state_load(st, &w.iso_state[r]);
                broyden_advance<4>(st, dx);
                normalize3(fp, st.x, xn);
            }
            s.xs[tid][0] = xn[0]; s.xs[tid][1] = xn[1]; s.xs[tid][2] = xn[2]; s.xs[tid][3] = 0.f;
        }
        __syncthreads();
        skin_tile_forward<PLAIN>(fp, s.xs, s.A, LDA_SDF, wp, s.logits, 0.f);
        __syncthreads();
        sdf_tile_forward<PLAIN>(fp, s.xs, s.A, LDA_SDF, wp, s.sdfo, nullptr, 0.f);
        __syncthreads();
        if (tid < TM) {
            bool active = false;
            if (r >= 0) {
                float g[4], T12[12];
                iso_residual(fp, w, r, st.x, s.logits[tid], s.sdfo[tid], g, T12);
                active = broyden_update<4>(st, dx, g, T12);
                if (iter + 1 >= BROYDEN_ITERS) active = false;
                state_store(&w.iso_state[r], st);
            }
            if (iter + 1 < BROYDEN_ITERS) warp_append(active, r, next, &w.counters[C_ISO + iter + 1]);
            warp_stat_add(r >= 0 ? 1 : 0, &w.counters[C_STAT_ISO_EVALS]);
        }
        __syncthreads();
    }
}

// accept/reject the joint-search result, emit the tracer outputs, place the z samples and enqueue the on-samples
__global__ void k_trace_finish(FrameParams fp, Work w) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int count = 0;
    if (r < w.P) {
        const int S = w.S;
        const float nr = w.near_far[2 * r], fr = w.near_far[2 * r + 1];
        const uint8_t fl = w.ray_flags[r];
        float xopt[3], zopt;
        bool conv = false;
        if (w.train || !(fl & 2)) {
            const BroydenState<4>& st = w.iso_state[r];
            xopt[0] = st.best_x[0]; xopt[1] = st.best_x[1]; xopt[2] = st.best_x[2]; zopt = st.best_x[3];
            conv = st.best_n < CVG_THRESH;
        } else {
            unnormalize3(fp, w.ray_cur[r].xn, xopt);
            zopt = w.ray_t[r];
        }
        conv = conv && (zopt >= nr) && (zopt <= fr);                      // ray_tracing.py:266
        float pn[3];
        normalize3(fp, xopt, pn);
        const float dist = conv ? zopt : nr;                              // :274-278
        w.ray_conv[r] = conv ? 1 : 0;
        w.ray_dist[r] = dist;
        w.ray_pnorm[3 * r] = pn[0]; w.ray_pnorm[3 * r + 1] = pn[1]; w.ray_pnorm[3 * r + 2] = pn[2];
        if (w.out_points_cam) {                                           // implicit_differentiable_renderer.py:114-115,142-143,251
            const bool surf = conv && fabsf(pn[0]) <= 1.0f && fabsf(pn[1]) <= 1.0f && fabsf(pn[2]) <= 1.0f;
            float pc[3] = {0.f, 0.f, 0.f};
            if (surf) {
                float pw[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) pw[k] = ((fp.cam_loc[k] + dist * w.ray_dirs[3 * r + k]) - fp.trans[k]) + fp.trans[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) pc[k] = pw[0] * fp.pose[k * 4] + pw[1] * fp.pose[k * 4 + 1] + pw[2] * fp.pose[k * 4 + 2] + fp.pose[k * 4 + 3];
            }
            w.out_points_cam[3 * r] = pc[0]; w.out_points_cam[3 * r + 1] = pc[1]; w.out_points_cam[3 * r + 2] = pc[2];
        }
        // ---- z placement (ray_sampler, ray_tracing.py:317-350); training adds perturb_z_vals (:298-311)
        float* z = w.z_vals + (size_t)r * S;
        const int nn = fp.near_samples + 1, nf = fp.far_samples;
        const bool tr = w.train != 0;
        // stratified jitter of an analytic ascending run f(0..n-1): lower/upper = mid points to the neighbours
        auto jitter = [&](auto f, int i, int n, const float* u, int fix) -> float {
            const float zi = f(i);
            if (!tr) return zi;
            const float lo = (i == 0) ? zi : 0.5f * (zi + f(i - 1));
            const float up = (i == n - 1) ? zi : 0.5f * (f(i + 1) + zi);
            const float t = (i == fix) ? 0.5f : u[i];
            return __fadd_rn(lo, __fmul_rn(up - lo, t));
        };
        auto f_all = [&](int i) -> float { return dist + (fr - dist) * linspace01(i, S); };
        const float* ua = tr ? w.u_all + (size_t)r * S : nullptr;
        if (!conv) {
            for (int i = 0; i < S; ++i) z[i] = jitter(f_all, i, S, ua, -1);
            count = S;
        } else {
            count = nn + nf;
            for (int i = count; i < S; ++i) z[i] = jitter(f_all, i, S, ua, -1);
            // merge of two ascending runs == torch.sort of their concatenation (:348)
            const float zs0 = dist - 0.05f;
            const float span = fmaxf(dist - 0.05f - nr, 1e-5f);
            auto f_near = [&](int i) -> float { return zs0 + 0.1f * linspace01(i, nn); };
            auto f_far = [&](int i) -> float { return nr + span * linspace01(i, nf); };
            const float* un = tr ? w.u_near + (size_t)r * nn : nullptr;
            const float* uf = (tr && nf > 0) ? w.u_far + (size_t)r * nf : nullptr;
            int a = 0, b = 0;
            float va = jitter(f_near, 0, nn, un, fp.near_samples / 2);
            float vb = (nf > 0) ? jitter(f_far, 0, nf, uf, -1) : INFINITY;
            for (int k = 0; k < count; ++k) {
                if (b >= nf || (a < nn && va <= vb)) { z[k] = va; ++a; va = (a < nn) ? jitter(f_near, a, nn, un, fp.near_samples / 2) : INFINITY; }
                else { z[k] = vb; ++b; vb = (b < nf) ? jitter(f_far, b, nf, uf, -1) : INFINITY; }
            }
        }
        // slots that are off keep zeros / false (generate_point_samples_opt scatters into zeros, :549-555)
        for (int i = count; i < S; ++i) {
            const size_t sl = (size_t)r * S + i;
            w.smp_conv[sl] = 0;
            w.smp_xn[3 * sl] = 0.f; w.smp_xn[3 * sl + 1] = 0.f; w.smp_xn[3 * sl + 2] = 0.f;
        }
        if (conv) atomicAdd(&w.counters[C_STAT_HIT_RAYS], 1);
    }
    // enqueue on-samples: the first `count` slots of the ray; one warp-aggregated reservation per warp
    const int lane = threadIdx.x & 31;
    int incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 0 && total) base = atomicAdd(&w.counters[C_ON], total);
    base = __shfl_sync(0xffffffffu, base, 0) + incl - count;
    if (r < w.P) w.ray_on_base[r] = base;
    for (int i = 0; i < count; ++i) w.on_list[base + i] = r * w.S + i;
}

// ================================================================================================ correspondences
__global__ void __launch_bounds__(512, 1) k_knn_samples(FrameParams fp, KnnIndex ix, Work w) {
    extern __shared__ float4 sv[];
    const int n = w.counters[C_ON];
    if (blockIdx.x == 0 && threadIdx.x == 0) w.counters[C_CORR] = n;
    const int B = knn_batch_size<KNN_COOP_SAMPLES>(n);
    if ((int)(blockIdx.x * (blockDim.x >> 5)) * B >= n) return;
    KnnSmem kk = load_knn(sv, ix);
    load_knn_quarters(sv + knn_smem_bytes(fp.n_verts) / 16, kk);
    auto finish = [&](int i, const float* x, int idx) {
        BroydenState<3> st;
        float s_, xh[3];
        nn_inverse_skinning(fp, idx, x, st.best_T, &s_, xh);
        if (w.corr_seed) {                                       // persistent correspondence kernel: 80-byte start record
            float4* sp = reinterpret_cast<float4*>(w.corr_seed + i);
            sp[0] = make_float4(xh[0], xh[1], xh[2], __int_as_float(w.on_list[i]));
            sp[1] = make_float4(st.best_T[0], st.best_T[1], st.best_T[2], st.best_T[3]);
            sp[2] = make_float4(st.best_T[4], st.best_T[5], st.best_T[6], st.best_T[7]);
            sp[3] = make_float4(st.best_T[8], st.best_T[9], st.best_T[10], st.best_T[11]);
            sp[4] = make_float4(x[0] - fp.trans[0], x[1] - fp.trans[1], x[2] - fp.trans[2], 0.f);
            return;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) { st.x[k] = xh[k]; st.best_x[k] = xh[k]; st.tgt[k] = x[k] - fp.trans[k]; st.gx[k] = 0.f; st.upd[k] = 0.f; }
#pragma unroll
        for (int k = 0; k < 9; ++k) st.Jinv[k] = 0.f;
        st.best_n = 0.f; st.owner = w.on_list[i]; st.g_evals = 0;
        state_store(&w.corr_state[i], st);
    };
    // (2 = default: ray-major only when the rays fill at least half of the grid's lanes — a 2048-ray training batch would leave
    // 97 % of them idle and walk 64 samples serially per lane; 3 = always, for the tests)
    if (B == 32 && (w.knn_seed == 3 || (w.knn_seed == 2 && 2 * w.P >= (int)(gridDim.x * blockDim.x)))) {
        // throughput regime, ray-major: a lane walks ALL on-samples of one ray front to back, each query seeded with the previous
        // winner; the lanes of a warp hold adjacent rays (adjacent pixels), so at every step their queries lie within centimetres
        // of each other and scan the same clusters: the per-lane search stays converged (round 1's run-of-4 mapping put two whole
        // rays, ~2 m of depth range, into one warp)
        const int total = gridDim.x * blockDim.x;
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < w.P; r += total) {
            const int cnt = w.ray_conv[r] ? (fp.near_samples + 1 + fp.far_samples) : w.S;
            const int base = w.ray_on_base[r];
            const float d0 = w.ray_dirs[3 * r], d1 = w.ray_dirs[3 * r + 1], d2 = w.ray_dirs[3 * r + 2];
            const float* zr = w.z_vals + (size_t)r * w.S;
            int slot = -1;
            for (int i = 0; i < cnt; ++i) {
                const float z = zr[i];
                const float x[3] = {d0 * z + fp.cam_loc[0], d1 * z + fp.cam_loc[1], d2 * z + fp.cam_loc[2]};
                const int idx = knn_scan_seeded(kk, x[0], x[1], x[2], slot);
                finish(base + i, x, idx);
            }
        }
        return;
    }
    if (B == 32 && w.knn_seed) {
        // throughput regime: a lane walks a run of KNN_RUN consecutive on-samples (neighbours on one ray), each query seeded by
        // the previous winner while the ray stays the same
        constexpr int KNN_RUN = 4;
        const int total = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
        for (long long base = (long long)t0 * KNN_RUN; base < n; base += (long long)total * KNN_RUN) {
            int slot = -1, prev_ray = -1;
            for (int j = 0; j < KNN_RUN; ++j) {
                const int i = (int)base + j;
                if (i >= n) break;
                const int sl = w.on_list[i];
                const int r = sl / w.S;
                const float z = w.z_vals[sl];
                float x[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) x[k] = w.ray_dirs[3 * r + k] * z + fp.cam_loc[k];
                if (r != prev_ray) slot = -1;
                prev_ray = r;
                const int idx = knn_scan_seeded(kk, x[0], x[1], x[2], slot);
                finish(i, x, idx);
            }
        }
        return;
    }
    knn_warp_batches<KNN_COOP_SAMPLES>(kk, n, B,
        [&](int i, float* x) {
            const int sl = w.on_list[i];
            const int r = sl / w.S;
            const float z = w.z_vals[sl];
#pragma unroll
            for (int k = 0; k < 3; ++k) x[k] = w.ray_dirs[3 * r + k] * z + fp.cam_loc[k];
        },
        finish);
}

// iter == -1: initial evaluation g(x0) + J^-1 init (root_finding_utils.py:327-328) over all on-samples;
// iter >= 0 : Broyden step `iter` over the active list.
__global__ void __launch_bounds__(256, 2) k_corr_step(FrameParams fp, Work w, int iter) {
    extern __shared__ __align__(128) float smem[];
    const int n = (iter < 0) ? w.counters[C_ON] : w.counters[C_CORR + iter];
    if ((int)blockIdx.x * TM >= n) return;
    TileSmem s = carve(smem, LDA_SKIN);
    WPipe wp;
    wpipe_init(wp, s.wbuf, s.bars);
    const int* list = (iter <= 0) ? nullptr : ((iter & 1) ? w.listB : w.listA);     // step 0 runs on every on-sample
    int* next = (iter & 1) ? w.listA : w.listB;
    const int tid = threadIdx.x;
    for (int tile = blockIdx.x; tile * TM < n; tile += gridDim.x) {
        int id = -1;
        BroydenState<3> st;
        float dx[3];
        if (tid < TM) {
            const int i = tile * TM + tid;
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) {
                id = list ? list[i] : i;
                state_load(st, &w.corr_state[id]);
                if (iter >= 0) broyden_advance<3>(st, dx);
                normalize3(fp, st.x, xn);
            }
            s.xs[tid][0] = xn[0]; s.xs[tid][1] = xn[1]; s.xs[tid][2] = xn[2]; s.xs[tid][3] = 0.f;
        }
        __syncthreads();
        skin_tile_forward<PLAIN>(fp, s.xs, s.A, LDA_SKIN, wp, s.logits, 0.f);
        __syncthreads();
        if (tid < TM) {
            bool active = false;
            if (id >= 0) {
                float T12[12], xb[3], g[3];
                skin_point(fp, s.logits[tid], st.x, T12, xb);
#pragma unroll
                for (int k = 0; k < 3; ++k) g[k] = xb[k] - st.tgt[k];
                if (iter < 0) {
                    float A3[9], Ai[9], Tinit[12];
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                        for (int c = 0; c < 3; ++c) A3[rr * 3 + c] = T12[rr * 4 + c];
                    invert3(A3, Ai);
#pragma unroll
                    for (int e = 0; e < 12; ++e) Tinit[e] = st.best_T[e];
                    const float x0[3] = {st.x[0], st.x[1], st.x[2]};
                    const int owner = st.owner;
                    const float tg[3] = {st.tgt[0], st.tgt[1], st.tgt[2]};
                    broyden_begin<3>(st, x0, g, Ai, Tinit);
                    st.owner = owner; st.tgt[0] = tg[0]; st.tgt[1] = tg[1]; st.tgt[2] = tg[2];
                    st.g_evals = 2;                         // J-init evaluation + g(x0) (the reference evaluates twice)
                    state_store(&w.corr_state[id], st);
                } else {
                    active = broyden_update<3>(st, dx, g, T12);
                    if (iter + 1 >= BROYDEN_ITERS) active = false;
                    if (active) state_store(&w.corr_state[id], st);
                    else corr_finalize(fp, w, st);
                }
            }
            if (iter >= 0) {
                if (iter + 1 < BROYDEN_ITERS) warp_append(active, id, next, &w.counters[C_CORR + iter + 1]);
                const bool done = (id >= 0) && !active;
                warp_append(done && st.best_n < CVG_THRESH, done ? st.owner : 0, w.shade_list, &w.counters[C_SHADE]);
                warp_stat_add(done ? st.g_evals : 0, &w.counters[C_STAT_CORR_EVALS]);
            }
        }
        __syncthreads();
    }
}

// ================================================================================================ shading
// smem: wbuf | A[TM][308] | xs[TM][4] | cin[TM][36] | grad[TM][4] | sdfo[TM] | bars
__host__ __device__ constexpr size_t shade_smem_bytes() {
    return (size_t)(WBUF_FLOATS + TM * LDA_SHADE + TM * 4 + TM * 36 + TM * 4 + TM) * 4 + 64;
}

__global__ void __launch_bounds__(256, 1) k_shade(FrameParams fp, Work w) {
    extern __shared__ __align__(128) float smem[];
    const int n = w.counters[C_SHADE];
    if ((int)blockIdx.x * TM >= n) return;
    float* wbuf = smem;
    float* A = smem + WBUF_FLOATS;
    float* p = A + TM * LDA_SHADE;
    float (*xs)[4] = reinterpret_cast<float (*)[4]>(p); p += TM * 4;
    float (*cin)[36] = reinterpret_cast<float (*)[36]>(p); p += TM * 36;
    float (*grad)[4] = reinterpret_cast<float (*)[4]>(p); p += TM * 4;
    float* sdfo = p; p += TM;
    uint64_t* bars = reinterpret_cast<uint64_t*>(p);
    WPipe wp;
    wpipe_init(wp, wbuf, bars);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* cf = w.scratch + (size_t)blockIdx.x * 7 * TM * SDF_H;      // [6][TM][256] cos factors + [TM][256] features
    float* feat = cf + (size_t)6 * TM * SDF_H;
    float* Arow = A + warp * 8 * LDA_SHADE;
    for (int tile = blockIdx.x; tile * TM < n; tile += gridDim.x) {
        int sl = -1;
        if (tid < TM) {
            const int i = tile * TM + tid;
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) { sl = w.shade_list[i]; xn[0] = w.smp_xn[3 * (size_t)sl]; xn[1] = w.smp_xn[3 * (size_t)sl + 1]; xn[2] = w.smp_xn[3 * (size_t)sl + 2]; }
            xs[tid][0] = xn[0]; xs[tid][1] = xn[1]; xs[tid][2] = xn[2]; xs[tid][3] = 0.f;
        }
        __syncthreads();
        // ---- SDF forward (keeps 30 f cos(arg) per layer) and the 256-d feature (implicit_differentiable_renderer.py:336-337)
        sdf_tile_forward<PLAIN>(fp, xs, A, LDA_SHADE, wp, sdfo, cf, 0.f);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float h[8];
            const float* row = Arow + r * LDA_SHADE;
            const float4 a = *reinterpret_cast<const float4*>(row + 4 * lane), b = *reinterpret_cast<const float4*>(row + 128 + 4 * lane);
            h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w; h[4] = b.x; h[5] = b.y; h[6] = b.z; h[7] = b.w;
            store_row<256>(feat + (size_t)(warp * 8 + r) * SDF_H, h, lane);
        }
        __syncwarp();
        // ---- gradient wrt the normalised point (autograd in the reference, :338)
        sdf_tile_backward(fp, A, LDA_SHADE, wp, cf, grad);
        __syncthreads();
        // ---- per-sample colour inputs: xn, PE(view), normal (rotated to posed space unless cano_view_dirs)
        if (tid < TM) {
            float v[3] = {0.f, 0.f, 0.f}, nrm[3] = {0.f, 0.f, 0.f};
            if (sl >= 0) {
                const int r = sl / w.S;
                const float* T = w.smp_T + 12 * (size_t)sl;
                const float d[3] = {w.ray_dirs[3 * r], w.ray_dirs[3 * r + 1], w.ray_dirs[3 * r + 2]};
                const float g[3] = {grad[tid][0], grad[tid][1], grad[tid][2]};
                if (fp.cano_view_dirs) {
                    float A3[9], Ai[9];
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                        for (int c = 0; c < 3; ++c) A3[rr * 3 + c] = T[rr * 4 + c];
                    invert3(A3, Ai);
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) { v[rr] = Ai[rr * 3] * -d[0] + Ai[rr * 3 + 1] * -d[1] + Ai[rr * 3 + 2] * -d[2]; nrm[rr] = g[rr]; }
                } else {
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) { v[rr] = -d[rr]; nrm[rr] = T[rr * 4] * g[0] + T[rr * 4 + 1] * g[1] + T[rr * 4 + 2] * g[2]; }
                }
                w.smp_sdf[sl] = sdf_to_metres(sdfo[tid], fp.cmin, fp.cmax);
            }
            float* c = cin[tid];
            c[0] = xs[tid][0]; c[1] = xs[tid][1]; c[2] = xs[tid][2];
            c[3] = v[0]; c[4] = v[1]; c[5] = v[2];
            int q = 6;
#pragma unroll
            for (int l = 0; l < 4; ++l) {                      // embedder.py:11-36, multires_view = 4
                const float fr = (float)(1 << l);
#pragma unroll
                for (int k = 0; k < 3; ++k) c[q++] = sinf(v[k] * fr);
#pragma unroll
                for (int k = 0; k < 3; ++k) c[q++] = cosf(v[k] * fr);
            }
            c[30] = nrm[0]; c[31] = nrm[1]; c[32] = nrm[2]; c[33] = 0.f; c[34] = 0.f; c[35] = 0.f;
        }
        __syncthreads();
        // ---- colour MLP (decoder.py:69-124); A[row] = [feat | cin(33) | 0...]
        auto fill_input = [&]() {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float h[8];
                load_cols_rw<256>(h, feat + (size_t)(warp * 8 + r) * SDF_H, lane);   // same thread wrote these values
                float* row = Arow + r * LDA_SHADE;
                store_row<256>(row, h, lane);
                for (int k = lane; k < COL_IN_PAD - 256; k += 32) row[256 + k] = (k < 33) ? cin[warp * 8 + r][k] : 0.f;
            }
            __syncwarp();
        };
        fill_input();
        float acc[8][8];
        auto relu_store = [&](const float* bias) {
            float b[8];
            load_cols<256>(b, bias, lane);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float h[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) h[c] = fmaxf(acc[r][c] + b[c], 0.f);
                store_row<256>(Arow + r * LDA_SHADE, h, lane);
            }
            __syncwarp();
        };
        tile_gemm<256>(acc, Arow, LDA_SHADE, COL_IN_PAD, fp.col_Wt0, wp);
        relu_store(fp.col_b[0]);
        tile_gemm<256>(acc, Arow, LDA_SHADE, 256, fp.col_Wt1, wp);
        relu_store(fp.col_b[1]);
        {
            float acc2[8][4], b[4];
            tile_gemm<128>(acc2, Arow, LDA_SHADE, 256, fp.col_Wt2, wp);
            load_cols<128>(b, fp.col_b[2], lane);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float h[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) h[c] = fmaxf(acc2[r][c] + b[c], 0.f);
                store_row<128>(Arow + r * LDA_SHADE, h, lane);
            }
            __syncwarp();
        }
        tile_gemm<256>(acc, Arow, LDA_SHADE, 128, fp.col_Wt3b, wp);             // skip connection, lin2 part ...
        fill_input();
        tile_gemm<256, true>(acc, Arow, LDA_SHADE, COL_IN_PAD, fp.col_Wt3a, wp);  // ... + network-input part (:113-115)
        relu_store(fp.col_b[3]);
        tile_gemm<256>(acc, Arow, LDA_SHADE, 256, fp.col_Wt4, wp);
        relu_store(fp.col_b[4]);
        float rgb[3][8];
#pragma unroll
        for (int j = 0; j < 3; ++j) warp_rows_dot(Arow, LDA_SHADE, 256, fp.col_W5 + j * 256, rgb[j], lane);
        if (lane < 8) {
            const int row = warp * 8 + lane;
            const int i = tile * TM + row;
            if (i < n) {
                const int sl2 = w.shade_list[i];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float v = rgb[j][0];
#pragma unroll
                    for (int r = 1; r < 8; ++r) if (lane == r) v = rgb[j][r];
                    w.smp_rgb[3 * (size_t)sl2 + j] = sigmoid_(v + __ldg(fp.col_b[5] + j));
                }
            }
        }
        __syncthreads();
    }
}

// ================================================================================================ compositing
// one warp per ray: converged samples are compacted in slot order into a per-warp shared-memory strip
// (ballot/popc ranks), then alpha, the transmittance product (warp shuffle scan) and the weighted colour sum run
// 32 samples at a time (implicit_differentiable_renderer.py:366-394).
constexpr int COMP_WARPS = 4;
// alpha of compacted sample e of a ray (implicit_differentiable_renderer.py:379-387); shared by k_composite and k_alpha_cull so
// that both evaluate the identical fp32 expression
__device__ __forceinline__ float sample_alpha(const float* cz, const float* cd, int e, int len, int S, int last_pt) {
    const float dz = (e + 1 < len) ? (cz[e + 1] - cz[e]) : (last_pt ? 1e10f : 1.0f / (float)S);     // :379-385
    return 1.0f - expf(-cd[e] * dz);
}

// Exact alpha cull.  A converged sample whose alpha is EXACTLY 0.0f has compositing weight alpha * T == 0, so its colour (and
// the SDF gradient that only feeds the colour network) cannot influence any output bit: rgb += 0 * c.  Given smp_sdf of every
// converged sample (k_sdf_fwd16), this kernel recomputes alpha with k_composite's own expression and builds the list of
// samples that still need the full shading pass; culled samples get rgb = 0 (k_composite multiplies it by 0).  Far from the
// surface sigma = exp(-sdf/beta)/(2 beta) underflows quickly: ~80 % of the samples of a bounding-box frame are culled.
__global__ void __launch_bounds__(32 * COMP_WARPS) k_alpha_cull(FrameParams fp, Work w, int* __restrict__ out_list) {
    __shared__ float strip[COMP_WARPS][2][MAX_STEPS];
    __shared__ int slot_of[COMP_WARPS][MAX_STEPS];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * COMP_WARPS + wib;
    if (r >= w.P) return;
    const int S = w.S;
    float beta = fabsf(fp.beta);
    beta = fminf(fmaxf(beta, 1e-6f), 1e6f);
    const float inv_beta = 1.0f / beta;
    float* cz = strip[wib][0]; float* cd = strip[wib][1];
    int* cs = slot_of[wib];
    int len = 0;
    for (int base = 0; base < S; base += 32) {
        const int i = base + lane;
        const size_t sl = (size_t)r * S + i;
        const bool valid = (i < S) && w.smp_conv[sl];
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int k = len + __popc(m & ((1u << lane) - 1u));
            cz[k] = w.z_vals[sl];
            cd[k] = laplace_density(w.smp_sdf[sl], inv_beta);
            cs[k] = i;
        }
        len += __popc(m);
    }
    __syncwarp();
    for (int base = 0; base < len; base += 32) {
        const int e = base + lane;
        const bool in = e < len;
        bool keep = false;
        int sl = 0;
        if (in) {
            sl = r * S + cs[e];
            keep = sample_alpha(cz, cd, e, len, S, fp.render_last_pt) != 0.0f;
            if (!keep) { w.smp_rgb[3 * (size_t)sl] = 0.f; w.smp_rgb[3 * (size_t)sl + 1] = 0.f; w.smp_rgb[3 * (size_t)sl + 2] = 0.f; }
        }
        warp_append(keep, sl, out_list, &w.counters[C_SHADE2]);
    }
}

__global__ void __launch_bounds__(32 * COMP_WARPS) k_composite(FrameParams fp, Work w) {
    __shared__ float strip[COMP_WARPS][5][MAX_STEPS];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * COMP_WARPS + wib;
    if (r >= w.P) return;
    const int S = w.S;
    float beta = fabsf(fp.beta);
    beta = fminf(fmaxf(beta, 1e-6f), 1e6f);
    const float inv_beta = 1.0f / beta;
    float* cz = strip[wib][0]; float* cd = strip[wib][1];
    float* cr = strip[wib][2]; float* cg = strip[wib][3]; float* cb = strip[wib][4];
    int len = 0;
    for (int base = 0; base < S; base += 32) {
        const int i = base + lane;
        const size_t sl = (size_t)r * S + i;
        const bool valid = (i < S) && w.smp_conv[sl];
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const int k = len + __popc(m & ((1u << lane) - 1u));
            cz[k] = w.z_vals[sl];
            cd[k] = laplace_density(w.smp_sdf[sl], inv_beta);
            cr[k] = w.smp_rgb[3 * sl]; cg[k] = w.smp_rgb[3 * sl + 1]; cb[k] = w.smp_rgb[3 * sl + 2];
        }
        len += __popc(m);
    }
    __syncwarp();
    float Tr = 1.0f, acc_r = 0.f, acc_g = 0.f, acc_b = 0.f, acc_w = 0.f;
    for (int base = 0; base < len; base += 32) {
        const int e = base + lane;
        const bool in = e < len;
        float alpha = 0.f;
        if (in) alpha = sample_alpha(cz, cd, e, len, S, fp.render_last_pt);
        const float fac = in ? (1.0f - alpha + 1e-7f) : 1.0f;
        float incl = fac;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl *= t; }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float wgt = in ? alpha * (Tr * excl) : 0.f;
        float sr = in ? wgt * cr[e] : 0.f, sg = in ? wgt * cg[e] : 0.f, sb = in ? wgt * cb[e] : 0.f, sw = wgt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, o); sg += __shfl_xor_sync(0xffffffffu, sg, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o); sw += __shfl_xor_sync(0xffffffffu, sw, o);
        }
        acc_r += sr; acc_g += sg; acc_b += sb; acc_w += sw;
        Tr *= __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
        w.out_rgb[3 * r] = acc_r; w.out_rgb[3 * r + 1] = acc_g; w.out_rgb[3 * r + 2] = acc_b;
        w.out_mask[r] = len > 0 ? 1 : 0;
        if (w.out_wsum) w.out_wsum[r] = fminf(fmaxf(acc_w, 0.f), 1.f);
        if (len > 0) atomicAdd(&w.counters[C_STAT_VOL_RAYS], 1);
    }
}

// stage export for the BodyRayTracing.forward sub-boundary (ray_tracing.py:166-172): 4x4 transforms, zeros where off
__global__ void k_export_samples(Work w, int n_on_hit, float* pts, float* dists, float* T16, uint8_t* conv) {
    const size_t sl = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sl >= (size_t)w.P * w.S) return;
    const int r = (int)(sl / w.S), i = (int)(sl % w.S);
    const int n_on = w.ray_conv[r] ? n_on_hit : w.S;
    const bool on = i < n_on;
    if (pts) { pts[3 * sl] = w.smp_xn[3 * sl]; pts[3 * sl + 1] = w.smp_xn[3 * sl + 1]; pts[3 * sl + 2] = w.smp_xn[3 * sl + 2]; }
    if (dists) dists[sl] = w.z_vals[sl];
    if (conv) conv[sl] = w.smp_conv[sl];
    if (T16) {
        float* o = T16 + 16 * sl;
#pragma unroll
        for (int e = 0; e < 12; ++e) o[e] = on ? w.smp_T[12 * sl + e] : 0.f;
        o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = on ? 1.f : 0.f;
    }
}

}  // namespace arah
