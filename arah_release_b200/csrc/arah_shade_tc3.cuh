// arah_shade_tc3.cuh — k_shade_tc3: shading (SDF forward + reverse-mode gradient + colour MLP) on tcgen05, TF32 operands:
// activations in tensor memory (.ts MMA form), pre-swizzled weight chunks through a 5-slot TMA ring fed by a producer warp,
// chunk-granular overlap of MMA and epilogue.
//
// Roles (320 threads): warps 0-7 epilogue/compute, warp 8 TMA producer, warp 9 MMA issuer.
// TMEM is split into two 256-column regions R0/R1.  For GEMM g the activations A live in one region and the accumulators
// D in the other; the epilogue converts D to the next layer's A IN PLACE (batch by batch: tcgen05.ld 32 columns ->
// FiLM-sine / x cos-factor / ReLU -> tcgen05.st to the same columns) and arrives on a per-chunk "ready" mbarrier; the MMA
// warp issues chunk c of the next GEMM as soon as ready[c] and the weight slot are there, writing the next D into the
// region the previous A has vacated.  So layer l's epilogue and layer l+1's MMAs run concurrently, chunk by chunk; the
// static per-tile program (18 GEMM segments, chunk order 0,4,1,5,.. = the order the two column-halves finish) is shared
// by producer and issuer.
#pragma once
#include <cuda_bf16.h>

#include "arah_kernels.cuh"
#include "arah_tc2.cuh"

namespace arah {

struct ShadeTC {
    const float* sdf_Wt0;      // [3][256]
    const float* sdf_W0;       // [256][3]
    const float* sdf_F;        // [6][256]  30 f
    const float* sdf_G;        // [6][256]  30 (f b + phi)
    const float* sdf_fwd[5];   // layers 1..5, swizzled chunks of B[n=out][k=in]
    const float* sdf_bwd[5];   // layers 1..5, swizzled chunks of B[n=in][k=out]
    const float* sdf_w6;       // [256]
    const float* sdf_b6;       // [1] device scalar
    const float* col0;         // 10 chunks, N=256: k = [feat 256 | x,PE,n 33 | pad]
    const float* col1;         // 8 chunks
    const float* col2;         // 8 chunks, N=128
    const float* col3b;        // 4 chunks (lin2 output part of the skip layer)
    const float* col3a;        // 10 chunks (network-input part)
    const float* col4;         // 8 chunks
    const float* col_W5;       // [3][256]
    const float* col_b[6];
};

constexpr int TC_SCRATCH_FLOATS = 6 * UM * 256 / 2 + UM * 256;     // bf16 cos factors + fp32 feature, in float units
constexpr int TC3_THREADS = 320;
constexpr int TC3_NSLOTS = 5;
constexpr int TC3_NSEG = 18;      // 5 forward + 5 reverse + lin0 (2) + lin1 + lin2 + lin3 (3) + lin4

struct Seg {
    const float* w;      // weight chunk images
    uint16_t N;          // output columns (256 / 128)
    uint8_t wbase;       // first weight chunk of this segment inside w
    uint8_t nchunks;
    uint8_t a_reg, d_reg;// TMEM region of A / D
    uint8_t acc;         // accumulate onto D from the first MMA
    uint8_t order;       // 0: 0,4,1,5,2,6,3,7   1: 0,2,1,3   2: 0,1
};
__device__ __forceinline__ int seg_chunk(int order, int i) {
    if (order == 0) return (i >> 1) + ((i & 1) << 2);
    if (order == 1) return (i >> 1) + ((i & 1) << 1);
    return i;
}

// per-CTA parameter block: 3584 floats of first / last layer weights (+ 512 scratch) and, staged once instead of re-read from
// L2 before every layer (a load + two CTA barriers per layer: 13 % of the SDF-only pass), the FiLM factors F, G of the six SDF
// layers (2 x 1536) and the five colour biases (1280)
constexpr int TC3_PRM_FLOATS = 3584 + 3072 + 1280;
__host__ __device__ constexpr size_t shade_tc3_smem_bytes() {
    // ring | cin[128][36] | params[3584] | xs[128][4] | part[2][128][4] | prog | barriers
    return (size_t)(TC3_NSLOTS * RING_SLOT_FLOATS + UM * 36 + TC3_PRM_FLOATS + UM * 4 + 2 * UM * 4) * 4 + TC3_NSEG * sizeof(Seg) + 512 + 1024;
}

// SDF_ONLY = true: the forward pass alone (same GEMM program prefix, same epilogue arithmetic, hence bit-identical smp_sdf) for
// the exact alpha cull (k_alpha_cull): no cos factors, no scratch traffic, no reverse pass, no colour MLP.
template <bool SDF_ONLY>
__global__ void __launch_bounds__(TC3_THREADS, 1) k_shade_tc3(FrameParams fp, ShadeTC tc, Work w) {
    extern __shared__ uint8_t raw_smem[];
    const int n = w.counters[w.shade_ctr];
    constexpr int NSEG = SDF_ONLY ? 5 : TC3_NSEG;
    if ((int)blockIdx.x * UM >= n) return;
    const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;
    float* sm = reinterpret_cast<float*>(raw_smem + (base - smem_u32(raw_smem)));
    float* ring = sm;
    float (*cin)[36] = reinterpret_cast<float (*)[36]>(ring + TC3_NSLOTS * RING_SLOT_FLOATS);
    float* prm = reinterpret_cast<float*>(cin) + UM * 36;      // [3584] per-tile constant columns, see P_* offsets
    float (*xs)[4] = reinterpret_cast<float (*)[4]>(prm + TC3_PRM_FLOATS);
    float (*part)[UM][4] = reinterpret_cast<float (*)[UM][4]>(reinterpret_cast<float*>(xs) + UM * 4);
    Seg* prog = reinterpret_cast<Seg*>(reinterpret_cast<float*>(part) + 2 * UM * 4);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(prog) + ((TC3_NSEG * sizeof(Seg) + 15) / 16) * 16);
    uint64_t* full = bars;                       // [5]
    uint64_t* empty = bars + TC3_NSLOTS;         // [5]
    uint64_t* ready = bars + 2 * TC3_NSLOTS;     // [8] A chunk c written by all 128 rows (4 warp arrivals)
    uint64_t* done_bar = ready + 8;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(done_bar + 1);
    // parameter block: everything the epilogues read per column, staged once per CTA
    constexpr int P_W0T = 0, P_W6 = 1280, P_W0 = 1536, P_W5 = 2304, P_LF = 3584, P_LG = P_LF + 1536, P_CB = P_LG + 1536;   // P_CB: col_b[0..4] at 0,256,512,640,896
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < TC3_NSLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 8; ++i) mbar_init(&ready[i], 4);
        mbar_init(done_bar, 1);
        mbar_fence_init();
        int s = 0;
        auto add = [&](const float* wp, int N, int wbase, int nch, int a, int d, int acc, int order) {
            Seg& g = prog[s++]; g.w = wp; g.N = (uint16_t)N; g.wbase = (uint8_t)wbase; g.nchunks = (uint8_t)nch;
            g.a_reg = (uint8_t)a; g.d_reg = (uint8_t)d; g.acc = (uint8_t)acc; g.order = (uint8_t)order;
        };
        for (int l = 1; l <= 5; ++l) add(tc.sdf_fwd[l - 1], 256, 0, 8, (l - 1) & 1, l & 1, 0, 0);          // A: R0,R1,R0,R1,R0
        if (!SDF_ONLY) {
            for (int l = 5; l >= 1; --l) add(tc.sdf_bwd[l - 1], 256, 0, 8, l & 1, (l - 1) & 1, 0, 0);      // A: R1,R0,R1,R0,R1
            add(tc.col0, 256, 0, 8, 1, 0, 0, 0);
            add(tc.col0, 256, 8, 2, 1, 0, 1, 2);
            add(tc.col1, 256, 0, 8, 0, 1, 0, 0);
            add(tc.col2, 128, 0, 8, 1, 0, 0, 0);
            add(tc.col3b, 256, 0, 4, 0, 1, 0, 1);
            add(tc.col3a, 256, 0, 8, 0, 1, 1, 0);
            add(tc.col3a, 256, 8, 2, 0, 1, 1, 2);
            add(tc.col4, 256, 0, 8, 1, 0, 0, 0);
        }
    }
    if (warp == 0) tmem_alloc(tslot, 512);
    for (int i = tid; i < 768; i += TC3_THREADS) { prm[P_W0T + i] = __ldg(tc.sdf_Wt0 + i); prm[P_W0 + i] = __ldg(tc.sdf_W0 + i); prm[P_W5 + i] = __ldg(tc.col_W5 + i); }
    for (int i = tid; i < 256; i += TC3_THREADS) prm[P_W6 + i] = __ldg(tc.sdf_w6 + i);
    for (int i = tid; i < 1536; i += TC3_THREADS) { prm[P_LF + i] = __ldg(tc.sdf_F + i); prm[P_LG + i] = __ldg(tc.sdf_G + i); }
    if (!SDF_ONLY) {
        const int cb_off[5] = {0, 256, 512, 640, 896}, cb_n[5] = {256, 256, 128, 256, 256};
        for (int l = 0; l < 5; ++l)
            for (int i = tid; i < cb_n[l]; i += TC3_THREADS) prm[P_CB + cb_off[l] + i] = __ldg(tc.col_b[l] + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;

    if (warp == 8) {                 // ===== TMA producer =====
        if (lane == 0) {
            uint32_t slot = 0, use = 0;
            for (int tile = blockIdx.x; tile * UM < n; tile += gridDim.x) {
                for (int s = 0; s < NSEG; ++s) {
                    const Seg g = prog[s];
                    const uint32_t bytes = (uint32_t)g.N * UK * 4;
                    for (int i = 0; i < g.nchunks; ++i) {
                        const int c = seg_chunk(g.order, i);
                        if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1u);
                        mbar_expect_tx(&full[slot], bytes);
                        bulk_g2s(ring + slot * RING_SLOT_FLOATS, reinterpret_cast<const char*>(g.w) + (size_t)(g.wbase + c) * bytes, bytes, &full[slot]);
                        if (++slot == TC3_NSLOTS) { slot = 0; ++use; }
                    }
                }
            }
        }
        return;
    }
    if (warp == 9) {                 // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t slot = 0, use = 0, rpar = 0;
            for (int tile = blockIdx.x; tile * UM < n; tile += gridDim.x) {
                for (int s = 0; s < NSEG; ++s) {
                    const Seg g = prog[s];
                    const uint32_t idesc = umma_idesc_tf32(UM, g.N);
                    const uint32_t ta = tbase + 256u * g.a_reg, td = tbase + 256u * g.d_reg;
                    for (int i = 0; i < g.nchunks; ++i) {
                        const int c = seg_chunk(g.order, i);
                        mbar_wait(&ready[c], (rpar >> c) & 1u);
                        rpar ^= (1u << c);
                        mbar_wait(&full[slot], use & 1u);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(ring + slot * RING_SLOT_FLOATS);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_tf32_ts(td, ta + (uint32_t)(c * UK + k * 8), umma_smem_desc_sw128(b_addr + k * 32), idesc, (i > 0 || k > 0) ? 1u : (uint32_t)g.acc);
                        umma_commit(&empty[slot]);
                        if (++slot == TC3_NSLOTS) { slot = 0; ++use; }
                    }
                    umma_commit(done_bar);
                }
            }
        }
        return;
    }
    // ===== compute / epilogue warps =====
    const int q = warp & 3, half = warp >> 2;
    const int r = 32 * q + lane;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    uint32_t done_par = 0;
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
    // write one 32-column batch of A (TMEM region `reg`, chunk `chunk`) and publish it to the MMA warp
    auto a_put = [&](int reg, int chunk, float (&v)[32]) {
        a_tmem_store(trow + 256u * reg + 32u * chunk, v);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[chunk]);
    };
    auto wait_done = [&]() {
        mbar_wait(done_bar, done_par);
        done_par ^= 1u;
        __syncwarp();
        tc_fence_after();
    };
    const float* lp0 = prm + P_LF;                                // per-column parameters of the current layer (F | bias)
    const float* lp1 = prm + P_LG;                                //                                             (G)

    PhaseClk pc; pc.start((tid == 32 && w.phase_clk) ? w.phase_clk + 8 : nullptr);
    for (int tile = blockIdx.x; tile * UM < n; tile += gridDim.x) {
        int sl = -1;
        if (tid < UM) {
            const int i = tile * UM + tid;
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) { sl = w.shade_list[i]; xn[0] = w.smp_xn[3 * (size_t)sl]; xn[1] = w.smp_xn[3 * (size_t)sl + 1]; xn[2] = w.smp_xn[3 * (size_t)sl + 2]; }
            xs[tid][0] = xn[0]; xs[tid][1] = xn[1]; xs[tid][2] = xn[2]; xs[tid][3] = 0.f;
        }
        // ================= SDF forward =================
        lp0 = prm + P_LF; lp1 = prm + P_LG;
        cta_sync_compute();                                       // xs visible
        {   // layer 0 (K = 3) on the FP32 pipe -> A1 in R0
            const float x = xs[r][0], y = xs[r][1], z = xs[r][2];
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int col0 = 128 * half + 32 * b;
                float h[32], c[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int cc = col0 + i;
                    const float a = fmaf(prm[P_W0T + 512 + cc], z, fmaf(prm[P_W0T + 256 + cc], y, prm[P_W0T + cc] * x));
                    float s_, c_;
                    __sincosf(fmaf(a, lp0[cc], lp1[cc]), &s_, &c_);
                    h[i] = s_; c[i] = c_ * lp0[cc];
                }
                a_put(0, col0 / 32, h);
                if (!SDF_ONLY) cf_put(0, b, c);
            }
        }
        pc.mark(0);                                               // tile setup + layer 0
        for (int l = 1; l < 6; ++l) {
            lp0 = prm + P_LF + l * 256; lp1 = prm + P_LG + l * 256;
            wait_done();                                          // GEMM l complete: D in R[l&1]
            pc.mark(1);                                           // waiting for forward GEMMs
            const int dreg = l & 1;
            float dot = 0.f;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int col0 = 128 * half + 32 * b;
                float v[32], c[32];
                tmem_ld32(trow + 256u * dreg + (uint32_t)col0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float s_, c_;
                    __sincosf(fmaf(v[i], lp0[col0 + i], lp1[col0 + i]), &s_, &c_);
                    v[i] = s_; c[i] = c_ * lp0[col0 + i];
                }
                if (!SDF_ONLY) cf_put(l, b, c);
                if (l < 5) a_put(dreg, col0 / 32, v);             // in place: D(l) -> A(l+1)
                else {
                    if (!SDF_ONLY) feat_put(b, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) dot = fmaf(v[i], prm[P_W6 + col0 + i], dot);
                    if (!SDF_ONLY) {
                        float g[32];                              // g_a5 = w6 * cf5, in place in R1 (A of the first reverse GEMM)
#pragma unroll
                        for (int i = 0; i < 32; ++i) g[i] = c[i] * prm[P_W6 + col0 + i];
                        a_put(1, col0 / 32, g);
                    }
                }
            }
            if (l == 5) part[half][r][0] = dot;
            pc.mark(2);                                           // forward epilogues
        }
        cta_sync_compute();
        if (tid < UM && sl >= 0 && (SDF_ONLY || !w.shade_keep_sdf)) w.smp_sdf[sl] = sdf_to_metres(part[0][tid][0] + part[1][tid][0] + __ldg(tc.sdf_b6), fp.cmin, fp.cmax);
        if (SDF_ONLY) { cta_sync_compute(); pc.mark(3); continue; }      // (part / xs are rewritten by the next tile)
        // ================= reverse pass =================
        float g3[3] = {0.f, 0.f, 0.f};
        for (int l = 5; l >= 1; --l) {
            wait_done();                                          // g_h(l-1) = g_a(l) @ W_l in R[(l-1)&1]
            pc.mark(3);                                           // waiting for reverse GEMMs
            const int dreg = (l - 1) & 1;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int col0 = 128 * half + 32 * b;
                float v[32], c[32];
                cf_get(l - 1, b, c);
                tmem_ld32(trow + 256u * dreg + (uint32_t)col0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= c[i];
                if (l > 1) a_put(dreg, col0 / 32, v);
                else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float* w0 = prm + P_W0 + (col0 + i) * 3;
                        g3[0] = fmaf(v[i], w0[0], g3[0]); g3[1] = fmaf(v[i], w0[1], g3[1]); g3[2] = fmaf(v[i], w0[2], g3[2]);
                    }
                }
            }
            pc.mark(4);                                           // reverse epilogues
        }
        // ---- feature part of colour lin0: A <- feat in R1 (free since reverse GEMM l=1 completed).  Its accumulators go to
        // R0, which the other warps may still be reading (reverse epilogue l=1) -> everyone must be out of R0 first.
        cta_sync_compute();
#pragma unroll 1
        for (int b = 0; b < 4; ++b) { float v[32]; feat_get(b, v); a_put(1, (128 * half + 32 * b) / 32, v); }
        part[half][r][0] = g3[0]; part[half][r][1] = g3[1]; part[half][r][2] = g3[2];
        cta_sync_compute();
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
                for (int j = 0; j < 3; ++j) c[k++] = sin_cw(v[j] * fr);
#pragma unroll
                for (int j = 0; j < 3; ++j) c[k++] = sin_cw(v[j] * fr + 1.57079632679489662f);      // cos; |arg| <= 8, TF32 consumer
            }
            c[30] = nrm[0]; c[31] = nrm[1]; c[32] = nrm[2]; c[33] = 0.f; c[34] = 0.f; c[35] = 0.f;
        }
        lp0 = prm + P_CB;
        cta_sync_compute();                                       // publishes cin
        auto fill_cin = [&](int reg) {
            if (half == 0) {
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const int k = 32 * c + i; v[i] = (k < 33) ? cin[r][k] : 0.f; }
                    a_put(reg, c, v);
                }
            }
        };
        auto relu_epilogue = [&](int N, int dreg, bool store) {
            const int per = N / 2;
            float acc3[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
            for (int b = 0; b < per / 32; ++b) {
                const int col0 = per * half + 32 * b;
                float v[32];
                tmem_ld32(trow + 256u * dreg + (uint32_t)col0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + lp0[col0 + i], 0.f);
                if (store) a_put(dreg, col0 / 32, v);
                else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        acc3[0] = fmaf(v[i], prm[P_W5 + col0 + i], acc3[0]);
                        acc3[1] = fmaf(v[i], prm[P_W5 + 256 + col0 + i], acc3[1]);
                        acc3[2] = fmaf(v[i], prm[P_W5 + 512 + col0 + i], acc3[2]);
                    }
                }
            }
            if (!store) { part[half][r][0] = acc3[0]; part[half][r][1] = acc3[1]; part[half][r][2] = acc3[2]; }
        };
        pc.mark(5);                                               // feature refill + colour inputs
        // ================= colour MLP =================
        wait_done();                                              // lin0, feature part done (A = R1 chunks free again)
        fill_cin(1);
        wait_done();                                              // lin0 complete, D in R0
        relu_epilogue(256, 0, true);
        lp0 = prm + P_CB + 256;
        wait_done();                                              // lin1, D in R1
        relu_epilogue(256, 1, true);
        lp0 = prm + P_CB + 512;
        wait_done();                                              // lin2 (N = 128), D in R0[0..127]
        relu_epilogue(128, 0, true);                              // -> A chunks 0..3 of R0
        lp0 = prm + P_CB + 640;
        wait_done();                                              // lin3, lin2-output part done -> R0 may be overwritten
#pragma unroll 1
        for (int b = 0; b < 4; ++b) { float v[32]; feat_get(b, v); a_put(0, (128 * half + 32 * b) / 32, v); }
        wait_done();                                              // lin3, feature part done
        fill_cin(0);
        wait_done();                                              // lin3 complete, D in R1
        relu_epilogue(256, 1, true);
        lp0 = prm + P_CB + 896;
        wait_done();                                              // lin4, D in R0
        relu_epilogue(256, 0, false);                             // lin5 (256 -> 3) folded into the epilogue
        cta_sync_compute();
        if (tid < UM && sl >= 0) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
                w.smp_rgb[3 * (size_t)sl + j] = sigmoid_(part[0][tid][j] + part[1][tid][j] + __ldg(tc.col_b[5] + j));
        }
        cta_sync_compute();
        pc.mark(6);                                               // colour MLP (waits + epilogues)
    }
    tc_fence_before();
    cta_sync_compute();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace arah
