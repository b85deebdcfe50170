// arah_shade_tc2.cuh — k_shade_tc2: shading on tcgen05 with engine v2 (arah_tc2.cuh): activations in TMEM (.ts MMA),
// 6 x 32 KB weight ring filled by a TMA producer warp that runs ahead across layers, 288 threads.
// Algorithm, scratch layout and epilogues are those of k_shade_tc (arah_shade_tc.cuh).
#pragma once
#include "arah_shade_tc.cuh"
#include "arah_tc2.cuh"

namespace arah {

__host__ __device__ constexpr size_t shade_tc2_smem_bytes() {
    return (size_t)(TC_NSLOTS * RING_SLOT_FLOATS + UM * 36 + 2 * 256 + UM * 4 + 2 * UM * 4) * 4 + 256 + 1024;
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_shade_tc2(FrameParams fp, ShadeTC tc, Work w) {
    extern __shared__ uint8_t raw_smem[];
    const int n = w.counters[C_SHADE];
    if ((int)blockIdx.x * UM >= n) return;
    const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;
    float* sm = reinterpret_cast<float*>(raw_smem + (base - smem_u32(raw_smem)));
    float* ring = sm;
    float (*cin)[36] = reinterpret_cast<float (*)[36]>(ring + TC_NSLOTS * RING_SLOT_FLOATS);
    float* lp0 = reinterpret_cast<float*>(cin) + UM * 36;     // per-layer column parameters
    float* lp1 = lp0 + 256;
    float (*xs)[4] = reinterpret_cast<float (*)[4]>(lp1 + 256);
    float (*part)[UM][4] = reinterpret_cast<float (*)[UM][4]>(reinterpret_cast<float*>(xs) + UM * 4);   // [2][UM][4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(part) + 2 * UM * 4);         // full[6] empty[6] done
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * TC_NSLOTS + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2;
    const int r = 32 * q + lane;                 // this thread's row (TMEM lane)
    TCRing rg; rg.buf = ring; rg.full = bars; rg.empty = bars + TC_NSLOTS;
    uint64_t* done_bar = bars + 2 * TC_NSLOTS;
    if (tid == 0) { tcring_init(rg); mbar_init(done_bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    if (warp == 8) {                 // ===== TMA producer warp: runs ahead through the static chunk schedule =====
        if (lane == 0) {
            RingPos pp; pp.slot = 0; pp.use = 0;
            for (int tile = blockIdx.x; tile * UM < n; tile += gridDim.x) {
                for (int l = 0; l < 5; ++l) tcring_produce(rg, pp, tc.sdf_fwd[l], 8, 32768u);
                for (int l = 4; l >= 0; --l) tcring_produce(rg, pp, tc.sdf_bwd[l], 8, 32768u);
                tcring_produce(rg, pp, tc.col0, 10, 32768u);
                tcring_produce(rg, pp, tc.col1, 8, 32768u);
                tcring_produce(rg, pp, tc.col2, 8, 16384u);
                tcring_produce(rg, pp, tc.col3b, 4, 32768u);
                tcring_produce(rg, pp, tc.col3a, 10, 32768u);
                tcring_produce(rg, pp, tc.col4, 8, 32768u);
            }
        }
        return;
    }
    const uint32_t trowA = tbase + ((uint32_t)(32 * q) << 16);          // activations: TMEM columns [0, 256)
    const uint32_t trow = trowA + 256u;                                    // accumulators: TMEM columns [256, 512)
    RingPos cp; cp.slot = 0; cp.use = 0;
    uint32_t done_par = 0;

    // per-thread private scratch, slot-major [slot][256 threads] uint4: a warp access = 512 contiguous bytes.
    // slots 0..95: bf16 cos factors (layer l, batch b, 4 x uint4); slots 96..127: fp32 feature (batch b, 8 x float4)
    uint4* scr = reinterpret_cast<uint4*>(w.scratch + (size_t)blockIdx.x * TC_SCRATCH_FLOATS) + tid;
    auto cf_put = [&](int l, int b, const float (&v)[32]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * i], v[8 * i + 1]), p1 = __floats2bfloat162_rn(v[8 * i + 2], v[8 * i + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * i + 4], v[8 * i + 5]), p3 = __floats2bfloat162_rn(v[8 * i + 6], v[8 * i + 7]);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
            u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
            scr[(size_t)((l * 4 + b) * 4 + i) * 256] = u;
        }
    };
    auto cf_get = [&](int l, int b, float (&v)[32]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 u = scr[(size_t)((l * 4 + b) * 4 + i) * 256];
            const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wd[j]));
                v[8 * i + 2 * j] = f.x; v[8 * i + 2 * j + 1] = f.y;
            }
        }
    };
    auto feat_put = [&](int b, const float (&v)[32]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) scr[(size_t)(96 + b * 8 + j) * 256] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
    };
    auto feat_get = [&](int b, float (&v)[32]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 u = scr[(size_t)(96 + b * 8 + j) * 256];
            v[4 * j] = __uint_as_float(u.x); v[4 * j + 1] = __uint_as_float(u.y); v[4 * j + 2] = __uint_as_float(u.z); v[4 * j + 3] = __uint_as_float(u.w);
        }
    };

    // A is complete (generic-proxy writes) and TMEM reads are done -> hand over to the MMA issuer
    auto handoff = [&]() { tmem_st_wait(); tc_fence_before(); cta_sync_compute(); tc_fence_after(); };
    auto gemm = [&](const float* Wsw, int nchunks, int N, uint32_t accumulate) {
        (void)Wsw;
        if (tid == 0) tcring_mma_layer(rg, cp, tbase, nchunks, N, tbase + 256u, accumulate, done_bar);
        mbar_wait(done_bar, done_par);
        done_par ^= 1u;
        __syncwarp();
        tc_fence_after();
    };
    auto load_params = [&](const float* p0, const float* p1) {
        lp0[tid] = p0 ? __ldg(p0 + tid) : 0.f;
        lp1[tid] = p1 ? __ldg(p1 + tid) : 0.f;
        cta_sync_compute();
    };

    for (int tile = blockIdx.x; tile * UM < n; tile += gridDim.x) {
        int sl = -1;
        if (tid < UM) {
            const int i = tile * UM + tid;
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) { sl = w.shade_list[i]; xn[0] = w.smp_xn[3 * (size_t)sl]; xn[1] = w.smp_xn[3 * (size_t)sl + 1]; xn[2] = w.smp_xn[3 * (size_t)sl + 2]; }
            xs[tid][0] = xn[0]; xs[tid][1] = xn[1]; xs[tid][2] = xn[2]; xs[tid][3] = 0.f;
        }
        // ================= SDF forward =================
        load_params(tc.sdf_F, tc.sdf_G);                       // also orders xs
        {   // layer 0 (K = 3) directly on the FP32 pipe
            const float x = xs[r][0], y = xs[r][1], z = xs[r][2];
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int col0 = 128 * half + 32 * b;
                float h[32], c[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int cc = col0 + i;
                    const float a = fmaf(__ldg(tc.sdf_Wt0 + 512 + cc), z, fmaf(__ldg(tc.sdf_Wt0 + 256 + cc), y, __ldg(tc.sdf_Wt0 + cc) * x));
                    float s_, c_;
                    __sincosf(fmaf(a, lp0[cc], lp1[cc]), &s_, &c_);
                    h[i] = s_; c[i] = c_ * lp0[cc];
                }
                a_tmem_store(trowA + (uint32_t)col0, h);
                cf_put(0, b, c);
            }
        }
        handoff();
        for (int l = 1; l < 6; ++l) {
            load_params(tc.sdf_F + l * 256, tc.sdf_G + l * 256);
            gemm(tc.sdf_fwd[l - 1], 8, 256, 0u);
            float dot = 0.f;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int col0 = 128 * half + 32 * b;
                float v[32], c[32];
                tmem_ld32(trow + (uint32_t)col0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float s_, c_;
                    __sincosf(fmaf(v[i], lp0[col0 + i], lp1[col0 + i]), &s_, &c_);
                    v[i] = s_; c[i] = c_ * lp0[col0 + i];
                }
                cf_put(l, b, c);
                if (l < 5) a_tmem_store(trowA + (uint32_t)col0, v);
                else {
                    feat_put(b, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) dot = fmaf(v[i], __ldg(tc.sdf_w6 + col0 + i), dot);
                }
            }
            if (l == 5) part[half][r][0] = dot;
            handoff();
        }
        if (tid < UM && sl >= 0) w.smp_sdf[sl] = sdf_to_metres(part[0][tid][0] + part[1][tid][0] + tc.sdf_b6, fp.cmin, fp.cmax);
        // ================= reverse pass: d sdf / d xn =================
        load_params(tc.sdf_w6, nullptr);
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {                            // g_a5 = w6 * cf5
            const int col0 = 128 * half + 32 * b;
            float c[32];
            cf_get(5, b, c);
#pragma unroll
            for (int i = 0; i < 32; ++i) c[i] *= lp0[col0 + i];
            a_tmem_store(trowA + (uint32_t)col0, c);
        }
        handoff();
        float g3[3] = {0.f, 0.f, 0.f};
        for (int l = 5; l >= 1; --l) {
            gemm(tc.sdf_bwd[l - 1], 8, 256, 0u);                 // g_h(l-1) = g_a(l) @ W_l
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int col0 = 128 * half + 32 * b;
                float v[32], c[32];
                tmem_ld32(trow + (uint32_t)col0, v);
                cf_get(l - 1, b, c);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= c[i];
                if (l > 1) a_tmem_store(trowA + (uint32_t)col0, v);
                else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float* w0 = tc.sdf_W0 + (col0 + i) * 3;
                        g3[0] = fmaf(v[i], __ldg(w0), g3[0]); g3[1] = fmaf(v[i], __ldg(w0 + 1), g3[1]); g3[2] = fmaf(v[i], __ldg(w0 + 2), g3[2]);
                    }
                }
            }
            if (l == 1) { part[half][r][0] = g3[0]; part[half][r][1] = g3[1]; part[half][r][2] = g3[2]; }
            handoff();
        }
        // ================= colour inputs =================
        if (tid < UM) {
            float v[3] = {0.f, 0.f, 0.f}, nrm[3] = {0.f, 0.f, 0.f};
            if (sl >= 0) {
                const int ray = sl / w.S;
                const float* T = w.smp_T + 12 * (size_t)sl;
                const float d[3] = {w.ray_dirs[3 * ray], w.ray_dirs[3 * ray + 1], w.ray_dirs[3 * ray + 2]};
                const float g[3] = {part[0][tid][0] + part[1][tid][0], part[0][tid][1] + part[1][tid][1], part[0][tid][2] + part[1][tid][2]};
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
            }
            float* c = cin[tid];
            c[0] = xs[tid][0]; c[1] = xs[tid][1]; c[2] = xs[tid][2];
            c[3] = v[0]; c[4] = v[1]; c[5] = v[2];
            int k = 6;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const float fr = (float)(1 << l);
#pragma unroll
                for (int j = 0; j < 3; ++j) c[k++] = sinf(v[j] * fr);
#pragma unroll
                for (int j = 0; j < 3; ++j) c[k++] = cosf(v[j] * fr);
            }
            c[30] = nrm[0]; c[31] = nrm[1]; c[32] = nrm[2]; c[33] = 0.f; c[34] = 0.f; c[35] = 0.f;
        }
        auto fill_feat = [&]() {                                  // A chunks 0..7 <- feature (this thread wrote these addresses)
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int col0 = 128 * half + 32 * b;
                float v[32];
                feat_get(b, v);
                a_tmem_store(trowA + (uint32_t)col0, v);
            }
        };
        auto fill_cin = [&]() {                                   // A chunks 0..1 <- [x, PE(view), n | 0 ...] (64 wide)
            if (half == 0) {
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const int k = 32 * c + i; v[i] = (k < 33) ? cin[r][k] : 0.f; }
                    a_tmem_store(trowA + (uint32_t)(32 * c), v);
                }
            }
        };
        auto relu_epilogue = [&](int N, const float* bias, bool store) {
            const int per = N / 2;                                // columns per thread
            float acc3[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
            for (int b = 0; b < per / 32; ++b) {
                const int col0 = per * half + 32 * b;
                float v[32];
                tmem_ld32(trow + (uint32_t)col0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + lp0[col0 + i], 0.f);
                if (store) a_tmem_store(trowA + (uint32_t)col0, v);
                else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        acc3[0] = fmaf(v[i], __ldg(tc.col_W5 + col0 + i), acc3[0]);
                        acc3[1] = fmaf(v[i], __ldg(tc.col_W5 + 256 + col0 + i), acc3[1]);
                        acc3[2] = fmaf(v[i], __ldg(tc.col_W5 + 512 + col0 + i), acc3[2]);
                    }
                }
            }
            if (!store) { part[half][r][0] = acc3[0]; part[half][r][1] = acc3[1]; part[half][r][2] = acc3[2]; }
            (void)bias;
        };
        // ================= colour MLP (decoder.py:69-124) =================
        fill_feat();
        handoff();                                                // (also publishes cin)
        load_params(tc.col_b[0], nullptr);
        gemm(tc.col0, 8, 256, 0u);                                // feature part
        fill_cin();
        handoff();
        gemm(tc.col0 + (size_t)8 * 256 * UK, 2, 256, 1u);         // + [x | PE | n] part
        relu_epilogue(256, tc.col_b[0], true);
        handoff();
        load_params(tc.col_b[1], nullptr);
        gemm(tc.col1, 8, 256, 0u);
        relu_epilogue(256, tc.col_b[1], true);
        handoff();
        load_params(tc.col_b[2], nullptr);
        gemm(tc.col2, 8, 128, 0u);
        relu_epilogue(128, tc.col_b[2], true);                    // 128 outputs -> A chunks 0..3
        handoff();
        load_params(tc.col_b[3], nullptr);
        gemm(tc.col3b, 4, 256, 0u);                               // skip layer: lin2 part ...
        fill_feat();
        handoff();
        gemm(tc.col3a, 8, 256, 1u);                               // ... + feature part ...
        fill_cin();
        handoff();
        gemm(tc.col3a + (size_t)8 * 256 * UK, 2, 256, 1u);        // ... + [x | PE | n] part (:113-115)
        relu_epilogue(256, tc.col_b[3], true);
        handoff();
        load_params(tc.col_b[4], nullptr);
        gemm(tc.col4, 8, 256, 0u);
        relu_epilogue(256, tc.col_b[4], false);                   // lin5 (256 -> 3) folded into the epilogue
        handoff();
        if (tid < UM && sl >= 0) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
                w.smp_rgb[3 * (size_t)sl + j] = sigmoid_(part[0][tid][j] + part[1][tid][j] + __ldg(tc.col_b[5] + j));
        }
        cta_sync_compute();
    }
    tc_fence_before();
    cta_sync_compute();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace arah
