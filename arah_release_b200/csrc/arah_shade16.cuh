// arah_shade16.cuh — k_shade16: gradient + colour pass of the samples that survive the exact alpha cull (SDF forward, reverse-mode
// gradient, colour MLP; renderer/implicit_differentiable_renderer.py:311-361) on tcgen05 kind::f16.
//
// Same program as k_shade_tc3 (round 1: TF32 operands) — 18 GEMM segments per 128-sample tile, activations in tensor memory,
// in-place D -> A epilogues, weight chunks through a 5-slot TMA ring fed by a producer warp, MMA-issuer warp — with fp16 operands:
// fp16 has TF32's 11-bit significand, issues at twice the rate and halves the weight stream, and k_shade_tc3 was bound by exactly
// that stream (9.2 TB/s L2 -> SM, profiles/r02_render_kernels_ncu.md).  No weight scaling is needed for a single-pass product:
// values below fp16's normal range keep an absolute error <= 2^-25, far below the 2^-11 relative operand rounding of everything
// else (the forward SDF layers reuse the pre-scaled hi images of the root-finding engine; their 1/s is folded into F).
// Layout of the A operand in a 256-column region (two K values per column):
//   map 0 (written by a 256-wide epilogue / layer 0 / the feature refill): K-chunk c (64 values) at columns 128 (c >> 1) + 32 (c & 1),
//         chunk 4 (the 33 colour inputs x | PE(view) | normal, zero padded) at columns 64..95;
//   map 1 (written by the 128-wide epilogue of lin2): K-chunk c at columns 64 c.
// In place: a thread reads 32 accumulator columns and overwrites 16 columns that lie inside what it has already read.
// Per-CTA scratch (global, L2 resident): the cos factors of layers 0..4 and the feature vector, both as fp16 (393 KB per CTA;
// round 1 kept six layers as bf16 plus an fp32 feature, 524 KB).  A line is discarded from L2 after its last read
// (discard.global.L2), so the dirty scratch lines are dropped instead of being written back to HBM.
#pragma once
#include <cuda_bf16.h>

#include "arah_f16x3.cuh"
#include "arah_tc2.cuh"
#include "arah_work.cuh"

namespace arah {

struct Shade16 {
    const float* sdf_Wt0;      // [3][256]
    const float* sdf_W0;       // [256][3]
    const float* sdf_F;        // [6][256]  30 f          (the kernel folds 1 / s_l of the scaled forward images in)
    const float* sdf_G;        // [6][256]  30 (f b + phi)
    const float* sdf_scale;    // [5][2]    (s, 1 / s) of the forward images, layers 1..5
    const __half* sdf_fwd;     // layers 1..5: 4 chunks x 256 x 64 (the hi images of arah_sdf16.cuh, scaled by s_l)
    const __half* sdf_bwd;     // layers 1..5: 4 chunks x 256 x 64, B[n = in][k = out] (W^T), unscaled
    const float* sdf_w6;       // [256]
    const float* sdf_b6;       // [1]
    const __half* col0;        // 5 chunks, N = 256: k = [feat 256 | x, PE, n 33 | pad]
    const __half* col1;        // 4 chunks
    const __half* col2;        // 4 chunks, N = 128
    const __half* col3b;       // 2 chunks (lin2-output part of the skip layer)
    const __half* col3a;       // 5 chunks (network-input part)
    const __half* col4;        // 4 chunks
    const float* col_W5;       // [3][256]
    const float* col_b[6];
};
constexpr size_t SHADE16_BWD_BYTES = 5 * 131072, SHADE16_COL_BYTES = (size_t)(5 + 4 + 2 + 5 + 4) * 32768 + 4 * 16384;

constexpr int SH16_THREADS = 320;
constexpr int SH16_NSLOTS = 5;
constexpr int SH16_NSEG = 18;
constexpr int SH16_PRM_FLOATS = 3584 + 3072 + 1280;
constexpr int SH16_SCRATCH_FLOATS = 96 * 256 * 4;     // 80 rows of cos factors (5 layers x 4 batches x 4) + 16 feature rows, 256 uint4 each

struct Seg16 {
    const __half* w;     // weight chunk images
    uint16_t N;          // output columns (256 / 128)
    uint8_t wbase;       // first weight chunk of this segment inside w
    uint8_t nchunks;
    uint8_t a_reg, d_reg;// TMEM region of A / D
    uint8_t acc;         // accumulate onto D from the first MMA
    uint8_t amap;        // 0 / 1: operand column map (see above); 2: the single colour-input chunk at columns 64..95
};
__device__ __forceinline__ uint32_t sh16_acol(int amap, int c) {
    return amap == 0 ? (uint32_t)(128 * (c >> 1) + 32 * (c & 1)) : (amap == 1 ? (uint32_t)(64 * c) : 64u);
}

__host__ __device__ constexpr size_t shade16_smem_bytes() {
    // ring | cin[128][36] | params | xs[128][4] | part[2][128][4] | prog | barriers
    return (size_t)(SH16_NSLOTS * 32768) + (size_t)(UM * 36 + SH16_PRM_FLOATS + UM * 4 + 2 * UM * 4) * 4 + SH16_NSEG * sizeof(Seg16) + 512 + 1024;
}

__global__ void __launch_bounds__(SH16_THREADS, 1) k_shade16(FrameParams fp, Shade16 tc, Work w) {
    extern __shared__ uint8_t raw_smem[];
    const int n = w.counters[w.shade_ctr];
    if ((int)blockIdx.x * UM >= n) return;
    const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;
    uint8_t* sm = raw_smem + (base - smem_u32(raw_smem));
    uint8_t* ring = sm;
    float (*cin)[36] = reinterpret_cast<float (*)[36]>(ring + SH16_NSLOTS * 32768);
    float* prm = reinterpret_cast<float*>(cin) + UM * 36;
    float (*xs)[4] = reinterpret_cast<float (*)[4]>(prm + SH16_PRM_FLOATS);
    float (*part)[UM][4] = reinterpret_cast<float (*)[UM][4]>(reinterpret_cast<float*>(xs) + UM * 4);
    Seg16* prog = reinterpret_cast<Seg16*>(reinterpret_cast<float*>(part) + 2 * UM * 4);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(prog) + ((SH16_NSEG * sizeof(Seg16) + 15) / 16) * 16);
    uint64_t* full = bars;                       // [5]
    uint64_t* empty = bars + SH16_NSLOTS;        // [5]
    uint64_t* ready = bars + 2 * SH16_NSLOTS;    // [5] A chunk c written by the 4 warps that own it
    uint64_t* done_bar = ready + 5;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(done_bar + 1);
    constexpr int P_W0T = 0, P_W6 = 1280, P_W0 = 1536, P_W5 = 2304, P_LF = 3584, P_LG = P_LF + 1536, P_CB = P_LG + 1536;   // P_CB: col_b[0..4] at 0,256,512,640,896
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < SH16_NSLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 5; ++i) mbar_init(&ready[i], 4);
        mbar_init(done_bar, 1);
        mbar_fence_init();
        int s = 0;
        auto add = [&](const __half* wp, int N, int wbase, int nch, int a, int d, int acc, int amap) {
            Seg16& g = prog[s++]; g.w = wp; g.N = (uint16_t)N; g.wbase = (uint8_t)wbase; g.nchunks = (uint8_t)nch;
            g.a_reg = (uint8_t)a; g.d_reg = (uint8_t)d; g.acc = (uint8_t)acc; g.amap = (uint8_t)amap;
        };
        for (int l = 1; l <= 5; ++l) add(tc.sdf_fwd + (size_t)(l - 1) * 65536, 256, 0, 4, (l - 1) & 1, l & 1, 0, 0);      // A: R0,R1,R0,R1,R0
        for (int l = 5; l >= 1; --l) add(tc.sdf_bwd + (size_t)(l - 1) * 65536, 256, 0, 4, l & 1, (l - 1) & 1, 0, 0);      // A: R1,R0,R1,R0,R1
        add(tc.col0, 256, 0, 4, 1, 0, 0, 0);          // lin0, feature part
        add(tc.col0, 256, 4, 1, 1, 0, 1, 2);          // lin0, colour-input chunk
        add(tc.col1, 256, 0, 4, 0, 1, 0, 0);
        add(tc.col2, 128, 0, 4, 1, 0, 0, 0);
        add(tc.col3b, 256, 0, 2, 0, 1, 0, 1);         // lin3, lin2-output part (K = 128)
        add(tc.col3a, 256, 0, 4, 0, 1, 1, 0);         // lin3, feature part
        add(tc.col3a, 256, 4, 1, 0, 1, 1, 2);         // lin3, colour-input chunk
        add(tc.col4, 256, 0, 4, 1, 0, 0, 0);
    }
    if (warp == 0) tmem_alloc(tslot, 512);
    for (int i = tid; i < 768; i += SH16_THREADS) { prm[P_W0T + i] = __ldg(tc.sdf_Wt0 + i); prm[P_W0 + i] = __ldg(tc.sdf_W0 + i); prm[P_W5 + i] = __ldg(tc.col_W5 + i); }
    for (int i = tid; i < 256; i += SH16_THREADS) prm[P_W6 + i] = __ldg(tc.sdf_w6 + i);
    for (int i = tid; i < 1536; i += SH16_THREADS) {
        // F multiplies the accumulator of the SCALED forward image: fold 1 / s_l in (exact power of two); cos factors use the unscaled F
        prm[P_LF + i] = __ldg(tc.sdf_F + i);
        prm[P_LG + i] = __ldg(tc.sdf_G + i);
    }
    {
        const int cb_off[5] = {0, 256, 512, 640, 896}, cb_n[5] = {256, 256, 128, 256, 256};
        for (int l = 0; l < 5; ++l)
            for (int i = tid; i < cb_n[l]; i += SH16_THREADS) prm[P_CB + cb_off[l] + i] = __ldg(tc.col_b[l] + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;

    if (warp == 8) {                 // ===== TMA producer =====
        if (lane == 0) {
            uint32_t slot = 0, use = 0;
            for (int tile = blockIdx.x; tile * UM < n; tile += gridDim.x) {
                for (int s = 0; s < SH16_NSEG; ++s) {
                    const Seg16 g = prog[s];
                    const uint32_t bytes = (uint32_t)g.N * HK * 2;
                    for (int c = 0; c < g.nchunks; ++c) {
                        if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1u);
                        mbar_expect_tx(&full[slot], bytes);
                        bulk_g2s(ring + slot * 32768, reinterpret_cast<const char*>(g.w) + (size_t)(g.wbase + c) * bytes, bytes, &full[slot]);
                        if (++slot == SH16_NSLOTS) { slot = 0; ++use; }
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
                for (int s = 0; s < SH16_NSEG; ++s) {
                    const Seg16 g = prog[s];
                    const uint32_t idesc = umma_idesc_f16(UM, g.N);
                    const uint32_t ta = tbase + 256u * g.a_reg, td = tbase + 256u * g.d_reg;
                    for (int c = 0; c < g.nchunks; ++c) {
                        const int rc = (g.amap == 2) ? 4 : c;                 // which ready barrier guards this operand chunk
                        mbar_wait(&ready[rc], (rpar >> rc) & 1u);
                        rpar ^= (1u << rc);
                        mbar_wait(&full[slot], use & 1u);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(ring + slot * 32768), a_col = ta + sh16_acol(g.amap, c);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ts(td, a_col + 8u * k, umma_smem_desc_sw128(b_addr + 32u * k), idesc, (c > 0 || k > 0) ? 1u : (uint32_t)g.acc);
                        umma_commit(&empty[slot]);
                        if (++slot == SH16_NSLOTS) { slot = 0; ++use; }
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
    uint4* scr = reinterpret_cast<uint4*>(w.scratch + (size_t)blockIdx.x * SH16_SCRATCH_FLOATS) + tid;
    // scratch rows: 32 values of one thread as 4 uint4 of packed halfs; row (l, b, i) of the cos factors at (l * 4 + b) * 4 + i,
    // feature row (b, i) at 80 + 4 b + i; `last`: drop the four 128-byte lines this warp has just read from L2
    auto pack16 = [&](const float (&v)[32], uint32_t (&p)[16]) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            p[i] = *reinterpret_cast<const uint32_t*>(&hh);
        }
    };
    auto row_put = [&](int row0, const uint32_t (&p)[16]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) scr[(size_t)(row0 + i) * 256] = make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
    };
    auto row_get = [&](int row0, uint32_t (&p)[16]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 u = scr[(size_t)(row0 + i) * 256];
            p[4 * i] = u.x; p[4 * i + 1] = u.y; p[4 * i + 2] = u.z; p[4 * i + 3] = u.w;
        }
    };
    auto row_discard = [&](int row0) {                          // call after the values have been consumed (warp-converged)
        __syncwarp();
        if ((lane & 7) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("discard.global.L2 [%0], 128;" ::"l"(scr + (size_t)(row0 + i) * 256) : "memory");
        }
    };
    auto cf_put = [&](int l, int b, const float (&v)[32]) { uint32_t p[16]; pack16(v, p); row_put((l * 4 + b) * 4, p); };
    auto cf_get = [&](int l, int b, float (&v)[32]) {
        uint32_t p[16];
        row_get((l * 4 + b) * 4, p);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&p[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    };
    // 32 activations -> 16 packed operand columns at `col` of region `reg`
    auto a_store = [&](int reg, uint32_t col, const float (&v)[32]) {
        uint32_t p[16];
        pack16(v, p);
        tmem_st16(trow + 256u * reg + col, p);
    };
    // this warp's share of operand chunk `chunk` is complete
    auto a_publish = [&](int chunk) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[chunk]);
    };
    // batch b (32 of this thread's 128 columns) of a 256-wide layer: columns 128 half + 32 b -> chunk 2 half + (b >> 1), map 0
    auto a_put256 = [&](int reg, int b, const float (&v)[32]) {
        a_store(reg, (uint32_t)(128 * half + 16 * b), v);
        if (b & 1) a_publish(2 * half + (b >> 1));
    };
    // feature vector (packed halfs in scratch) -> operand chunks 0..3 of region `reg`
    auto feat_refill = [&](int reg, bool last) {
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            uint32_t p[16];
            row_get(80 + 4 * b, p);
            tmem_st16(trow + 256u * reg + (uint32_t)(128 * half + 16 * b), p);
            if (b & 1) a_publish(2 * half + (b >> 1));
        }
        if (last) { for (int b = 0; b < 4; ++b) row_discard(80 + 4 * b); }      // (a_publish has waited for the stores)
    };
    auto wait_done = [&]() {
        mbar_wait(done_bar, done_par);
        done_par ^= 1u;
        __syncwarp();
        tc_fence_after();
    };
    const float* lp0 = prm + P_LF;
    const float* lp1 = prm + P_LG;

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
                a_put256(0, b, h);
                cf_put(0, b, c);
            }
        }
        pc.mark(0);
        for (int l = 1; l < 6; ++l) {
            lp0 = prm + P_LF + l * 256; lp1 = prm + P_LG + l * 256;
            const float inv = __ldg(tc.sdf_scale + 2 * (l - 1) + 1);     // the forward image of layer l is scaled by s_l
            wait_done();                                          // GEMM l complete: D in R[l&1]
            pc.mark(1);
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
                    __sincosf(fmaf(v[i] * inv, lp0[col0 + i], lp1[col0 + i]), &s_, &c_);
                    v[i] = s_; c[i] = c_ * lp0[col0 + i];
                }
                if (l < 5) { cf_put(l, b, c); a_put256(dreg, b, v); }      // in place: D(l) -> A(l+1)
                else {
                    { uint32_t p[16]; pack16(v, p); row_put(80 + 4 * b, p); }     // the feature vector of the colour network
#pragma unroll
                    for (int i = 0; i < 32; ++i) dot = fmaf(v[i], prm[P_W6 + col0 + i], dot);
                    float g[32];                                  // g_a5 = w6 * cf5, in place in R1 (A of the first reverse GEMM)
#pragma unroll
                    for (int i = 0; i < 32; ++i) g[i] = c[i] * prm[P_W6 + col0 + i];
                    a_put256(1, b, g);
                }
            }
            if (l == 5) part[half][r][0] = dot;
            pc.mark(2);
        }
        cta_sync_compute();
        if (tid < UM && sl >= 0 && !w.shade_keep_sdf) w.smp_sdf[sl] = sdf_to_metres(part[0][tid][0] + part[1][tid][0] + __ldg(tc.sdf_b6), fp.cmin, fp.cmax);
        // ================= reverse pass =================
        float g3[3] = {0.f, 0.f, 0.f};
        for (int l = 5; l >= 1; --l) {
            wait_done();                                          // g_h(l-1) = g_a(l) @ W_l in R[(l-1)&1]
            pc.mark(3);
            const int dreg = (l - 1) & 1;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
                const int col0 = 128 * half + 32 * b;
                float v[32], c[32];
                cf_get(l - 1, b, c);
                tmem_ld32(trow + 256u * dreg + (uint32_t)col0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= c[i];
                if (l > 1) a_put256(dreg, b, v);
                else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float* w0 = prm + P_W0 + (col0 + i) * 3;
                        g3[0] = fmaf(v[i], w0[0], g3[0]); g3[1] = fmaf(v[i], w0[1], g3[1]); g3[2] = fmaf(v[i], w0[2], g3[2]);
                    }
                }
            }
            // last use of the cos factors of layer l - 1: every lane's values have been consumed once the operand stores that
            // depend on them have completed (a_publish of batch 3 waits for them); layer 0's are dropped further down
            if (l > 1) { for (int b = 0; b < 4; ++b) row_discard(((l - 1) * 4 + b) * 4); }
            pc.mark(4);
        }
        // ---- feature part of colour lin0: A <- feat in R1 (free since reverse GEMM l=1 completed).  Its accumulators go to
        // R0, which the other warps may still be reading (reverse epilogue l=1) -> everyone must be out of R0 first.
        cta_sync_compute();
        feat_refill(1, false);
        part[half][r][0] = g3[0]; part[half][r][1] = g3[1]; part[half][r][2] = g3[2];
        for (int b = 0; b < 4; ++b) row_discard(4 * b);           // layer 0's cos factors (g3 above depends on all of them)
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
                for (int j = 0; j < 3; ++j) c[k++] = sin_cw(v[j] * fr + 1.57079632679489662f);      // cos; |arg| <= 8
            }
            c[30] = nrm[0]; c[31] = nrm[1]; c[32] = nrm[2]; c[33] = 0.f; c[34] = 0.f; c[35] = 0.f;
        }
        lp0 = prm + P_CB;
        cta_sync_compute();                                       // publishes cin
        // the 33 colour inputs + zero padding = operand chunk 4 (columns 64..95 of the region), written by the warps of half 0
        auto fill_cin = [&](int reg) {
            if (half == 0) {
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const int k = 32 * c + i; v[i] = (k < 33) ? cin[r][k] : 0.f; }
                    a_store(reg, 64u + 16u * c, v);
                }
                a_publish(4);
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
                if (store) {
                    if (N == 256) a_put256(dreg, b, v);
                    else {                                        // N = 128 (lin2): columns 64 half + 32 b -> chunk `half`, map 1
                        a_store(dreg, (uint32_t)(64 * half + 16 * b), v);
                        if (b & 1) a_publish(half);
                    }
                } else {
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
        pc.mark(5);
        // ================= colour MLP =================
        wait_done();                                              // lin0, feature part done
        fill_cin(1);
        wait_done();                                              // lin0 complete, D in R0
        relu_epilogue(256, 0, true);
        lp0 = prm + P_CB + 256;
        wait_done();                                              // lin1, D in R1
        relu_epilogue(256, 1, true);
        lp0 = prm + P_CB + 512;
        wait_done();                                              // lin2 (N = 128), D in R0[0..127]
        relu_epilogue(128, 0, true);                              // -> A chunks 0, 1 of R0 (map 1)
        lp0 = prm + P_CB + 640;
        wait_done();                                              // lin3, lin2-output part done -> R0 may be overwritten
        feat_refill(0, true);
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
        pc.mark(6);
    }
    tc_fence_before();
    cta_sync_compute();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace arah
