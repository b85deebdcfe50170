// arah_iso_init_tc.cuh — k_iso_init_tc3: the one-off Jacobian initialisation of the joint iso-surface / correspondence search
// (utils/root_finding_utils.py:401-418: forward_skinning_jac with d w / d x_hat, grad sdf, 4x4 inverse, g(u0)) on the tensor
// cores.  k_iso_init spent 4 ms per frame on FP32 FFMA tiles; here the skinning MLP and the SDF run on the 3xTF32 engine of
// arah_sdf3x.cuh in FORWARD MODE: a ray owns four consecutive rows of the 128-row tile (= four adjacent lanes of a warp):
// row 4p = value at x_hat, rows 4p+1..3 = the tangents d/dx_hat_k seeded with dn e_k (dn = d x_norm / d x_hat).  A linear layer
// maps value and tangent rows alike (one GEMM); the epilogue applies h = act(a) to value rows and h_t = act'(a) D_t to
// tangent rows, act'(a) coming from the ray's value lane by one warp shuffle per column.  32 rays per tile.
#pragma once
#include "arah_sdf3x.cuh"

namespace arah {

// d softplus(beta = 100) / d a = sigmoid(100 a)
__device__ __forceinline__ float sigmoid100(float a) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a * -144.26950408889634f));
    return __fdividef(1.0f, 1.0f + e);
}

// skinning MLP, value + 3 tangent rows per ray: logits[row][0..24] (value rows incl. bias; tangent rows d/dx_hat_k)
__device__ __forceinline__ void s3_compute_skin_dual(const SkinTC& sk, const float* xs3, const S3Bars& bar, uint32_t& done_par, uint32_t tbase,
                                                     float (*logits)[LGS], float dn) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2, r = 32 * q + lane, ty = lane & 3, vl = lane & ~3;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    auto put = [&](int chunk, const float (&v)[32]) {
        a_tmem_store_split(trow + 32u * chunk, trow + 128u + 32u * chunk, v);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar.ready[chunk]);
    };
    auto wait_done = [&]() { mbar_wait(bar.done, done_par); done_par ^= 1u; __syncwarp(); tc_fence_after(); };
    {
        const float x = xs3[3 * r], y = xs3[3 * r + 1], z = xs3[3 * r + 2];          // the four rows of a ray hold the same point
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
            const int col0 = 64 * half + 32 * b;
            float h[32], w0[32], w1[32], w2[32], pb[32];
            ldg32(sk.Wt0 + col0, w0); ldg32(sk.Wt0 + 128 + col0, w1); ldg32(sk.Wt0 + 256 + col0, w2); ldg32(sk.b[0] + col0, pb);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float a = fmaf(w2[i], z, fmaf(w1[i], y, w0[i] * x)) + pb[i];
                const float seed = dn * ((ty == 1) ? w0[i] : (ty == 2) ? w1[i] : w2[i]);
                h[i] = (ty == 0) ? softplus100_fast(a) : sigmoid100(a) * seed;
            }
            put(col0 / 32, h);
        }
    }
    for (int l = 1; l < 4; ++l) {
        wait_done();
        const uint32_t tD = trow + ((l & 1) ? 256u : 384u);
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
            const int col0 = 64 * half + 32 * b;
            float v[32], pb[32];
            ldg32(sk.b[l] + col0, pb);
            tmem_ld32(tD + (uint32_t)col0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float a = v[i] + pb[i];                                           // meaningful on value lanes
                const float sg = __shfl_sync(0xffffffffu, sigmoid100(a), vl);
                v[i] = (ty == 0) ? softplus100_fast(a) : sg * v[i];
            }
            put(col0 / 32, v);
        }
    }
    wait_done();
    if (half == 0) {
        float v[32], pb[32];
        ldg32(sk.b[4], pb);
        tmem_ld32(trow + 384u, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) logits[r][i] = (ty == 0) ? v[i] + pb[i] : v[i];
    }
    tc_fence_before();
}

// SDF, value + 3 tangent rows per ray; returns this thread's partial of w6 . h5 over its 128 columns
__device__ __forceinline__ float s3_compute_sdf_dual(const SdfTC& sd, const float* xs3, float* A_lo, const S3Bars& bar, uint32_t& done_par,
                                                     uint32_t tbase, float dn) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2, r = 32 * q + lane, ty = lane & 3, vl = lane & ~3;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    auto put = [&](int reg, int chunk, const float (&v)[32]) {
        float hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { hi[i] = tf32_rn(v[i]); lo[i] = v[i] - hi[i]; }
        tmem_st32(trow + 256u * reg + 32u * chunk, hi);
        a_store_chunk(A_lo, r, chunk, lo);
        fence_async_smem();
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar.ready[chunk]);
    };
    {
        const float x = xs3[3 * r], y = xs3[3 * r + 1], z = xs3[3 * r + 2];
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            const int col0 = 128 * half + 32 * b;
            float h[32], w0[32], w1[32], w2[32], pf[32], pb[32], pp[32];
            ldg32(sd.Wt0 + col0, w0); ldg32(sd.Wt0 + 256 + col0, w1); ldg32(sd.Wt0 + 512 + col0, w2);
            ldg32(sd.freq + col0, pf); ldg32(sd.b[0] + col0, pb); ldg32(sd.phase + col0, pp);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float a = fmaf(w2[i], z, fmaf(w1[i], y, w0[i] * x));
                float sn, cs;
                sincos_cw(30.0f * (pf[i] * (a + pb[i]) + pp[i]), sn, cs);
                const float seed = dn * ((ty == 1) ? w0[i] : (ty == 2) ? w1[i] : w2[i]);
                h[i] = (ty == 0) ? sn : (30.0f * pf[i] * cs) * seed;
            }
            put(0, col0 / 32, h);
        }
    }
    float dot = 0.f;
    for (int L = 1; L <= 5; ++L) {
        mbar_wait(bar.done, done_par);
        done_par ^= 1u;
        __syncwarp();
        tc_fence_after();
        const int dreg = L & 1;
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            const int col0 = 128 * half + 32 * b;
            float v[32], pf[32], pb[32], pp[32];
            ldg32(sd.freq + L * 256 + col0, pf); ldg32(sd.b[L] + col0, pb); ldg32(sd.phase + L * 256 + col0, pp);
            tmem_ld32(trow + 256u * dreg + (uint32_t)col0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float sn, cs;
                sincos_cw(30.0f * (pf[i] * (v[i] + pb[i]) + pp[i]), sn, cs);           // meaningful on value lanes
                const float cf = __shfl_sync(0xffffffffu, 30.0f * pf[i] * cs, vl);
                v[i] = (ty == 0) ? sn : cf * v[i];
            }
            if (L < 5) put(dreg, col0 / 32, v);
            else {
                ldg32(sd.w6 + col0, pf);
#pragma unroll
                for (int i = 0; i < 32; ++i) dot = fmaf(v[i], pf[i], dot);
            }
        }
    }
    return dot;
}

constexpr int ISO_TC_PTS = UM / 4;      // rays per tile

__global__ void __launch_bounds__(TC3_THREADS, 1) k_iso_init_tc3(FrameParams fp, SdfTC sd, SkinTC sk, Work w) {
    extern __shared__ __align__(1024) uint8_t raw_smem[];
    const int n = w.counters[C_ISO];
    if ((int)blockIdx.x * ISO_TC_PTS >= n) return;
    if (smem_u32(raw_smem) & 1023u) __trap();
    float* A_lo = reinterpret_cast<float*>(raw_smem);
    float (*logits)[LGS] = reinterpret_cast<float (*)[LGS]>(A_lo);        // aliases A_lo: only alive between the two MLPs
    float* ring = A_lo + S3_ALO_FLOATS;
    float* xs3 = ring + S3_RING_FLOATS;
    float (*part)[UM] = reinterpret_cast<float (*)[UM]>(xs3 + UM * 3);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xs3 + UM * 3 + 2 * UM);
    S3Bars bar; bar.full = bars; bar.empty = bars + S3_NSLOTS; bar.ready = bars + 2 * S3_NSLOTS; bar.done = bar.ready + 8;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar.done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < S3_NSLOTS; ++i) { mbar_init(&bar.full[i], 1); mbar_init(&bar.empty[i], 1); }
        for (int i = 0; i < 8; ++i) mbar_init(&bar.ready[i], 4);
        mbar_init(bar.done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    const int ntiles = (n + ISO_TC_PTS - 1) / ISO_TC_PTS;
    if (warp == 8) {
        if (lane == 0) {
            uint32_t slot = 0, use = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) { s3_produce_skin(ring, bar, slot, use, sk); s3_produce_sdf(ring, bar, slot, use, sd); }
        }
        return;
    }
    if (warp == 9) {
        if (lane == 0) {
            uint32_t slot = 0, use = 0, rpar = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) { s3_mma_skin(ring, bar, slot, use, rpar, tbase); s3_mma_sdf(ring, A_lo, bar, slot, use, rpar, tbase); }
        }
        return;
    }
    const int half = warp >> 2, r = 32 * (warp & 3) + lane;
    uint32_t done_par = 0;
    const float dn = 2.0f / (fp.cmax - fp.cmin) / 1.1f;                    // d x_norm / d x_hat
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int ray = -1;
        float x0[3] = {0.f, 0.f, 0.f};
        if (tid < ISO_TC_PTS) {
            const int i = tile * ISO_TC_PTS + tid;
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) {
                ray = w.listA[i];
                const RayCur& c = w.ray_cur[ray];
                unnormalize3(fp, c.xn, x0);              // ray_tracing.py:245
                normalize3(fp, x0, xn);                  // root_finding_utils.py:75 (inside query_weights)
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) { xs3[3 * (4 * tid + t)] = xn[0]; xs3[3 * (4 * tid + t) + 1] = xn[1]; xs3[3 * (4 * tid + t) + 2] = xn[2]; }
        }
        cta_sync_compute();
        s3_compute_skin_dual(sk, xs3, bar, done_par, tbase, logits, dn);
        cta_sync_compute();
        tc_fence_after();
        // ---- LBS value + full Jacobian incl. d w / d x_hat (root_finding_utils.py:406-418), by the ray's thread
        float Jl[9], xb[3], Tk[12];
        if (ray >= 0) {
            Dual3 lx[25], pw[NJ];
#pragma unroll
            for (int c = 0; c < 25; ++c) {
                lx[c].v = logits[4 * tid][c] * 20.0f;
#pragma unroll
                for (int k = 0; k < 3; ++k) lx[c].d[k] = logits[4 * tid + 1 + k][c] * 20.0f;
            }
            hierarchical_softmax_dual(lx, pw);
#pragma unroll
            for (int e = 0; e < 9; ++e) Jl[e] = 0.0f;
#pragma unroll
            for (int e = 0; e < 12; ++e) Tk[e] = 0.0f;
            for (int j = 0; j < NJ; ++j) {
                const float* B = fp.bone_T + j * 16;
                float bx[3];
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) bx[rr] = B[rr * 4] * x0[0] + B[rr * 4 + 1] * x0[1] + B[rr * 4 + 2] * x0[2] + B[rr * 4 + 3];
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) Jl[rr * 3 + c] += pw[j].v * B[rr * 4 + c] + bx[rr] * pw[j].d[c];
#pragma unroll
                    for (int c = 0; c < 4; ++c) Tk[rr * 4 + c] += pw[j].v * B[rr * 4 + c];
                }
            }
            apply_T(Tk, x0, xb);
        }
        cta_sync_compute();                                            // logits consumed: A_lo may be overwritten
        const float dot = s3_compute_sdf_dual(sd, xs3, A_lo, bar, done_par, tbase, dn);
        part[half][r] = dot;
        cta_sync_compute();
        if (ray >= 0) {
            float J[16], Ji[16], so4[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) so4[t] = part[0][4 * tid + t] + part[1][4 * tid + t];
            const float so = 1.0f / 2.0f * 1.1f * (fp.cmax - fp.cmin);        // d(sdf metres)/d(sdf raw)
#pragma unroll
            for (int c = 0; c < 3; ++c) J[c] = so4[1 + c] * so;
            J[3] = 0.0f;
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
                for (int c = 0; c < 3; ++c) J[(rr + 1) * 4 + c] = Jl[rr * 3 + c];
                J[(rr + 1) * 4 + 3] = -w.ray_dirs[3 * ray + rr];
            }
            invert_gj<4>(J, Ji);
            const float z0 = w.ray_t[ray];
            const float u0[4] = {x0[0], x0[1], x0[2], z0};
            float g0[4];
#pragma unroll
            for (int k = 0; k < 3; ++k) g0[1 + k] = xb[k] - ((w.ray_dirs[3 * ray + k] * z0 + fp.cam_loc[k]) - fp.trans[k]);
            g0[0] = sdf_to_metres(so4[0] + __ldg(sd.b6), fp.cmin, fp.cmax);
            BroydenState<4> st;
            broyden_begin<4>(st, u0, g0, Ji, w.ray_cur[ray].T);
            st.owner = ray;
            st.tgt[0] = st.tgt[1] = st.tgt[2] = 0.f;
            state_store(&w.iso_state[ray], st);
        }
        cta_sync_compute();
    }
    tc_fence_before();
    cta_sync_compute();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace arah
