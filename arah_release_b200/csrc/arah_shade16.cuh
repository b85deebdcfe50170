// arah_shade16.cuh — k_shade16: gradient + colour pass of the samples that survive the exact alpha cull (SDF forward, reverse-mode
// gradient, colour MLP; renderer/implicit_differentiable_renderer.py:311-361) on tcgen05 kind::f16.
//
// 18 GEMM segments per 128-sample tile (5 forward + 5 reverse SDF layers, lin0 (2), lin1, lin2, lin3 (3), lin4), activations in
// tensor memory, weight chunks (64 K-values x N rows of fp16, pre-swizzled) through a 5-slot TMA ring fed by a producer warp, an
// MMA-issuer warp, and SIXTEEN epilogue warps laid out like the root-finding engine (arah_sdf16.cuh): four per TMEM lane quarter;
// warp (q, u) owns accumulator columns [64 j + 16 u, +16) of every K-chunk j and overwrites them IN PLACE with the 8 packed fp16
// columns of K-step u of the next operand's chunk j, then arrives on ready[j]: chunk j of the next GEMM can start after a quarter
// of the epilogue.  (Round 1's TF32 kernel and the first fp16 version had eight epilogue warps with 128 columns per thread.)
// fp16 has TF32's 11-bit significand, issues at twice the rate and halves the weight stream.  No weight scaling is needed for a
// single-pass product: values below fp16's normal range keep an absolute error <= 2^-25 (the forward SDF layers reuse the
// pre-scaled hi images of the root-finding engine; their 1 / s is applied to the accumulator).
// Operand columns of a 256-column region: K-step (j, k) of a 256- or 128-wide activation at [64 j + 16 k, +8); the 33 colour inputs
// x | PE(view) | normal (zero padded to 64 = one K-chunk) sit in the gaps of chunk 0: K-step k at [16 k + 8, +8).
// Per-CTA scratch (global, L2 resident): the cos factors of layers 0..4 and the feature vector as fp16, 393 KB per CTA; a line is
// discarded from L2 after its last read (discard.global.L2), so dirty scratch lines are dropped instead of written back to HBM.
#pragma once
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

constexpr int SH16_WARPS = 16;
constexpr int SH16_CTHREADS = 32 * SH16_WARPS;
constexpr int SH16_THREADS = SH16_CTHREADS + 64;          // + producer warp (16) + MMA warp (17)
constexpr int SH16_NSLOTS = 5;
constexpr int SH16_NSEG = 18;
constexpr int SH16_PRM_FLOATS = 3584 + 3072 + 1280;
constexpr int SH16_SCRATCH_FLOATS = 48 * SH16_CTHREADS * 4;     // 40 rows of cos factors (5 layers x 4 chunks x 2) + 8 feature rows, 512 uint4 each

struct Seg16 {
    const __half* w;     // weight chunk images
    uint16_t N;          // output columns (256 / 128)
    uint8_t wbase;       // first weight chunk of this segment inside w
    uint8_t nchunks;
    uint8_t a_reg, d_reg;// TMEM region of A / D
    uint8_t acc;         // accumulate onto D from the first MMA
    uint8_t cin;         // the single colour-input chunk (gap columns of chunk 0) instead of activation chunks
};

__host__ __device__ constexpr size_t shade16_smem_bytes() {
    // ring | cin[128][36] | params | xs[128][4] | part[4][128][4] | prog | barriers
    return (size_t)(SH16_NSLOTS * 32768) + (size_t)(UM * 36 + SH16_PRM_FLOATS + UM * 4 + 4 * UM * 4) * 4 + SH16_NSEG * sizeof(Seg16) + 512 + 1024;
}

__device__ __forceinline__ void sh16_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void sh16_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
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
    float (*part)[UM][4] = reinterpret_cast<float (*)[UM][4]>(reinterpret_cast<float*>(xs) + UM * 4);      // [4][128][4]
    Seg16* prog = reinterpret_cast<Seg16*>(reinterpret_cast<float*>(part) + 4 * UM * 4);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(prog) + ((SH16_NSEG * sizeof(Seg16) + 15) / 16) * 16);
    uint64_t* full = bars;                       // [5]
    uint64_t* empty = bars + SH16_NSLOTS;        // [5]
    uint64_t* ready = bars + 2 * SH16_NSLOTS;    // [5] operand chunk j (4: the colour-input chunk) written by all 16 warps
    uint64_t* done_bar = ready + 5;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(done_bar + 1);
    constexpr int P_W0T = 0, P_W6 = 1280, P_W0 = 1536, P_W5 = 2304, P_LF = 3584, P_LG = P_LF + 1536, P_CB = P_LG + 1536;   // P_CB: col_b[0..4] at 0,256,512,640,896
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < SH16_NSLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 5; ++i) mbar_init(&ready[i], SH16_WARPS);
        mbar_init(done_bar, 1);
        mbar_fence_init();
        int s = 0;
        auto add = [&](const __half* wp, int N, int wbase, int nch, int a, int d, int acc, int cinseg) {
            Seg16& g = prog[s++]; g.w = wp; g.N = (uint16_t)N; g.wbase = (uint8_t)wbase; g.nchunks = (uint8_t)nch;
            g.a_reg = (uint8_t)a; g.d_reg = (uint8_t)d; g.acc = (uint8_t)acc; g.cin = (uint8_t)cinseg;
        };
        for (int l = 1; l <= 5; ++l) add(tc.sdf_fwd + (size_t)(l - 1) * 65536, 256, 0, 4, (l - 1) & 1, l & 1, 0, 0);      // A: R0,R1,R0,R1,R0
        for (int l = 5; l >= 1; --l) add(tc.sdf_bwd + (size_t)(l - 1) * 65536, 256, 0, 4, l & 1, (l - 1) & 1, 0, 0);      // A: R1,R0,R1,R0,R1
        add(tc.col0, 256, 0, 4, 1, 0, 0, 0);          // lin0, feature part
        add(tc.col0, 256, 4, 1, 1, 0, 1, 1);          // lin0, colour-input chunk
        add(tc.col1, 256, 0, 4, 0, 1, 0, 0);
        add(tc.col2, 128, 0, 4, 1, 0, 0, 0);
        add(tc.col3b, 256, 0, 2, 0, 1, 0, 0);         // lin3, lin2-output part (K = 128: chunks 0, 1)
        add(tc.col3a, 256, 0, 4, 0, 1, 1, 0);         // lin3, feature part
        add(tc.col3a, 256, 4, 1, 0, 1, 1, 1);         // lin3, colour-input chunk
        add(tc.col4, 256, 0, 4, 1, 0, 0, 0);
    }
    if (warp == 17) tmem_alloc(tslot, 512);
    for (int i = tid; i < 768; i += SH16_THREADS) { prm[P_W0T + i] = __ldg(tc.sdf_Wt0 + i); prm[P_W0 + i] = __ldg(tc.sdf_W0 + i); prm[P_W5 + i] = __ldg(tc.col_W5 + i); }
    for (int i = tid; i < 256; i += SH16_THREADS) prm[P_W6 + i] = __ldg(tc.sdf_w6 + i);
    for (int i = tid; i < 1536; i += SH16_THREADS) { prm[P_LF + i] = __ldg(tc.sdf_F + i); prm[P_LG + i] = __ldg(tc.sdf_G + i); }
    {
        const int cb_off[5] = {0, 256, 512, 640, 896}, cb_n[5] = {256, 256, 128, 256, 256};
        for (int l = 0; l < 5; ++l)
            for (int i = tid; i < cb_n[l]; i += SH16_THREADS) prm[P_CB + cb_off[l] + i] = __ldg(tc.col_b[l] + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;

    if (warp == 16) {                // ===== TMA producer =====
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
    if (warp == 17) {                // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t slot = 0, use = 0, rpar = 0;
            for (int tile = blockIdx.x; tile * UM < n; tile += gridDim.x) {
                for (int s = 0; s < SH16_NSEG; ++s) {
                    const Seg16 g = prog[s];
                    const uint32_t idesc = umma_idesc_f16(UM, g.N);
                    const uint32_t ta = tbase + 256u * g.a_reg, td = tbase + 256u * g.d_reg;
                    for (int c = 0; c < g.nchunks; ++c) {
                        const int rc = g.cin ? 4 : c;                         // which ready barrier guards this operand chunk
                        mbar_wait(&ready[rc], (rpar >> rc) & 1u);
                        rpar ^= (1u << rc);
                        mbar_wait(&full[slot], use & 1u);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(ring + slot * 32768), a_col = ta + (g.cin ? 8u : 64u * (uint32_t)c);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ts(td, a_col + 16u * k, umma_smem_desc_sw128(b_addr + 32u * k), idesc, (c > 0 || k > 0) ? 1u : (uint32_t)g.acc);
                        umma_commit(&empty[slot]);
                        if (++slot == SH16_NSLOTS) { slot = 0; ++use; }
                    }
                    umma_commit(done_bar);
                }
            }
        }
        __syncwarp();
        asm volatile("bar.sync 2, 544;" ::: "memory");      // the epilogue warps are out of tensor memory
        tmem_dealloc(tbase, 512);
        return;
    }
    // ===== compute / epilogue warps: q = TMEM lane quarter, u = which 16 columns of every 64-column chunk =====
    const int q = warp & 3, u = warp >> 2;
    const int r = 32 * q + lane;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    uint32_t done_par = 0;
    auto sync_c = [&]() { asm volatile("bar.sync 1, 512;" ::: "memory"); };
    uint4* scr = reinterpret_cast<uint4*>(w.scratch + (size_t)blockIdx.x * SH16_SCRATCH_FLOATS) + tid;
    // scratch rows: 16 values of one thread as 2 uint4 of packed halfs; cos factors of (layer l, chunk j) at row (l * 4 + j) * 2,
    // the feature chunk j at row 40 + 2 j
    auto pack8 = [&](const float (&v)[16], uint32_t (&p)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            p[i] = *reinterpret_cast<const uint32_t*>(&hh);
        }
    };
    auto row_put = [&](int row0, const uint32_t (&p)[8]) {
        scr[(size_t)row0 * SH16_CTHREADS] = make_uint4(p[0], p[1], p[2], p[3]);
        scr[(size_t)(row0 + 1) * SH16_CTHREADS] = make_uint4(p[4], p[5], p[6], p[7]);
    };
    auto row_get = [&](int row0, uint32_t (&p)[8]) {
        const uint4 a = scr[(size_t)row0 * SH16_CTHREADS], b = scr[(size_t)(row0 + 1) * SH16_CTHREADS];
        p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = a.w; p[4] = b.x; p[5] = b.y; p[6] = b.z; p[7] = b.w;
    };
    auto row_discard = [&](int row0) {                          // call after the values have been consumed (warp-converged)
        __syncwarp();
        if ((lane & 7) == 0) {
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(scr + (size_t)row0 * SH16_CTHREADS) : "memory");
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(scr + (size_t)(row0 + 1) * SH16_CTHREADS) : "memory");
        }
    };
    auto unpack8 = [&](const uint32_t (&p)[8], float (&v)[16]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&p[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    };
    // 16 activations (K = 64 j + 16 u ..) -> the 8 packed columns of K-step u of chunk j in region `reg`
    auto a_store = [&](int reg, int j, const uint32_t (&p)[8]) { sh16_st8(trow + 256u * reg + (uint32_t)(64 * j + 16 * u), p); };
    auto a_publish = [&](int chunk) {                           // this warp's share of operand chunk `chunk` is complete
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[chunk]);
    };
    auto a_put = [&](int reg, int j, const float (&v)[16]) {
        uint32_t p[8];
        pack8(v, p);
        a_store(reg, j, p);
        a_publish(j);
    };
    // feature vector (packed halfs in scratch) -> operand chunks 0..3 of region `reg`
    auto feat_refill = [&](int reg, bool last) {
        uint32_t p[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) row_get(40 + 2 * j, p[j]);      // all eight loads in flight: one L2 latency instead of four
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a_store(reg, j, p[j]);
            a_publish(j);
        }
        if (last) { for (int j = 0; j < 4; ++j) row_discard(40 + 2 * j); }      // (a_publish has waited for the stores)
    };
    // the 33 colour inputs + zero padding = one K-chunk in the gap columns of chunk 0: warp (q, u) writes K-step u
    auto fill_cin = [&](int reg) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { const int k = 16 * u + i; v[i] = (k < 33) ? cin[r][k] : 0.f; }
        uint32_t p[8];
        pack8(v, p);
        sh16_st8(trow + 256u * reg + (uint32_t)(16 * u + 8), p);
        a_publish(4);
    };
    // 16 consecutive per-column parameters as four 128-bit shared-memory loads (LDS shares the MIO queue with MUFU; every carve-out
    // and every column offset 64 j + 16 u is 16-byte aligned)
    auto lds16 = [&](const float* p, float (&o)[16]) {
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const float4 t = *reinterpret_cast<const float4*>(p + 4 * i4);
            o[4 * i4] = t.x; o[4 * i4 + 1] = t.y; o[4 * i4 + 2] = t.z; o[4 * i4 + 3] = t.w;
        }
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
    // the sample of row `tid` of a tile, fetched one tile AHEAD (two dependent loads: list entry -> point)
    auto fetch = [&](int tile, int& sl, float (&xn)[3]) {
        sl = -1; xn[0] = 0.f; xn[1] = 0.f; xn[2] = 0.f;
        const int i = tile * UM + tid;
        if (i < n) { sl = w.shade_list[i]; xn[0] = w.smp_xn[3 * (size_t)sl]; xn[1] = w.smp_xn[3 * (size_t)sl + 1]; xn[2] = w.smp_xn[3 * (size_t)sl + 2]; }
    };
    int sl_next = -1;
    float xn_next[3] = {0.f, 0.f, 0.f};
    if (tid < UM) fetch(blockIdx.x, sl_next, xn_next);
    for (int tile = blockIdx.x; tile * UM < n; tile += gridDim.x) {
        int sl = -1;
        float4 Tq[3] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        float dq[3] = {0.f, 0.f, 0.f};
        if (tid < UM) {
            sl = sl_next;
            xs[tid][0] = xn_next[0]; xs[tid][1] = xn_next[1]; xs[tid][2] = xn_next[2]; xs[tid][3] = 0.f;
            if (sl >= 0) {                                        // what the colour inputs need, in flight during the SDF passes
                const float4* Tp = reinterpret_cast<const float4*>(w.smp_T + 12 * (size_t)sl);
                Tq[0] = Tp[0]; Tq[1] = Tp[1]; Tq[2] = Tp[2];
                const int ray = sl / w.S;
                dq[0] = w.ray_dirs[3 * ray]; dq[1] = w.ray_dirs[3 * ray + 1]; dq[2] = w.ray_dirs[3 * ray + 2];
            }
            fetch(tile + (int)gridDim.x, sl_next, xn_next);
        }
        // ================= SDF forward =================
        lp0 = prm + P_LF; lp1 = prm + P_LG;
        sync_c();                                                 // xs visible
        {   // layer 0 (K = 3) on the FP32 pipe -> A1 in R0
            const float x = xs[r][0], y = xs[r][1], z = xs[r][2];
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int col0 = 64 * j + 16 * u;
                float h[16], c[16], wx[16], wy[16], wz[16], F[16], G[16];
                lds16(prm + P_W0T + col0, wx); lds16(prm + P_W0T + 256 + col0, wy); lds16(prm + P_W0T + 512 + col0, wz);
                lds16(lp0 + col0, F); lds16(lp1 + col0, G);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a = fmaf(wz[i], z, fmaf(wy[i], y, wx[i] * x));
                    float s_, c_;
                    __sincosf(fmaf(a, F[i], G[i]), &s_, &c_);
                    h[i] = s_; c[i] = c_ * F[i];
                }
                a_put(0, j, h);
                { uint32_t p[8]; pack8(c, p); row_put(j * 2, p); }
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
            for (int j = 0; j < 4; ++j) {
                const int col0 = 64 * j + 16 * u;
                float v[16], c[16], F[16], G[16];
                sh16_ld16(trow + 256u * dreg + (uint32_t)col0, v);
                lds16(lp0 + col0, F); lds16(lp1 + col0, G);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float s_, c_;
                    __sincosf(fmaf(v[i] * inv, F[i], G[i]), &s_, &c_);
                    v[i] = s_; c[i] = c_ * F[i];
                }
                if (l < 5) {                                      // in place: D(l) -> A(l+1)
                    a_put(dreg, j, v);
                    uint32_t p[8]; pack8(c, p); row_put((l * 4 + j) * 2, p);
                } else {
                    { uint32_t p[8]; pack8(v, p); row_put(40 + 2 * j, p); }       // the feature vector of the colour network
                    float w6[16], g[16];                          // g_a5 = w6 * cf5, in place in R1 (A of the first reverse GEMM)
                    lds16(prm + P_W6 + col0, w6);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { dot = fmaf(v[i], w6[i], dot); g[i] = c[i] * w6[i]; }
                    a_put(1, j, g);
                }
            }
            if (l == 5) part[u][r][0] = dot;
            pc.mark(2);
        }
        sync_c();
        if (tid < UM && sl >= 0 && !w.shade_keep_sdf)
            w.smp_sdf[sl] = sdf_to_metres(((part[0][tid][0] + part[1][tid][0]) + (part[2][tid][0] + part[3][tid][0])) + __ldg(tc.sdf_b6), fp.cmin, fp.cmax);
        // ================= reverse pass =================
        float g3[3] = {0.f, 0.f, 0.f};
        for (int l = 5; l >= 1; --l) {
            // the cos factors of layer l - 1 (written by this thread in the forward pass) are fetched from L2 while the GEMM runs:
            // with the loads inside the chunk loop every chunk paid a full L2 latency (the whole reverse epilogue was latency-bound)
            uint32_t cfp[4][8];
#pragma unroll
            for (int j = 0; j < 4; ++j) row_get(((l - 1) * 4 + j) * 2, cfp[j]);
            wait_done();                                          // g_h(l-1) = g_a(l) @ W_l in R[(l-1)&1]
            pc.mark(3);
            const int dreg = (l - 1) & 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col0 = 64 * j + 16 * u;
                float v[16], c[16];
                unpack8(cfp[j], c);
                sh16_ld16(trow + 256u * dreg + (uint32_t)col0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] *= c[i];
                if (l > 1) a_put(dreg, j, v);
                else {
#pragma unroll
                    for (int t3 = 0; t3 < 3; ++t3) {              // W0 rows of 16 columns = 48 consecutive floats
                        float w0[16];
                        lds16(prm + P_W0 + col0 * 3 + 16 * t3, w0);
#pragma unroll
                        for (int e = 0; e < 16; ++e) { const int f = 16 * t3 + e; g3[f % 3] = fmaf(v[f / 3], w0[e], g3[f % 3]); }
                    }
                }
            }
            // last use of the cos factors of layer l - 1: every lane's values have been consumed once the operand stores that
            // depend on them have completed (a_publish waits for them); layer 0's are dropped further down
            if (l > 1) { for (int j = 0; j < 4; ++j) row_discard(((l - 1) * 4 + j) * 2); }
            pc.mark(4);
        }
        // ---- feature part of colour lin0: A <- feat in R1 (free since reverse GEMM l=1 completed).  Its accumulators go to
        // R0, which the other warps may still be reading (reverse epilogue l=1) -> everyone must be out of R0 first.
        part[u][r][0] = g3[0]; part[u][r][1] = g3[1]; part[u][r][2] = g3[2];
        for (int j = 0; j < 4; ++j) row_discard(j * 2);          // layer 0's cos factors (g3 above depends on all of them)
        sync_c();
        feat_refill(1, false);
        // ================= colour inputs =================
        if (tid < UM) {
            float v[3] = {0.f, 0.f, 0.f}, nrm[3] = {0.f, 0.f, 0.f};
            if (sl >= 0) {
                const float T[12] = {Tq[0].x, Tq[0].y, Tq[0].z, Tq[0].w, Tq[1].x, Tq[1].y, Tq[1].z, Tq[1].w, Tq[2].x, Tq[2].y, Tq[2].z, Tq[2].w};
                const float d[3] = {dq[0], dq[1], dq[2]};
                float g[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) g[k] = (part[0][tid][k] + part[1][tid][k]) + (part[2][tid][k] + part[3][tid][k]);
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
        sync_c();                                                 // publishes cin
        auto relu_epilogue = [&](int N, int dreg, bool store) {
            float acc3[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
            for (int j = 0; j < N / 64; ++j) {
                const int col0 = 64 * j + 16 * u;
                float v[16], bb[16];
                sh16_ld16(trow + 256u * dreg + (uint32_t)col0, v);
                lds16(lp0 + col0, bb);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i] + bb[i], 0.f);
                if (store) a_put(dreg, j, v);
                else {
#pragma unroll
                    for (int t3 = 0; t3 < 3; ++t3) {
                        float w5[16];
                        lds16(prm + P_W5 + 256 * t3 + col0, w5);
#pragma unroll
                        for (int i = 0; i < 16; ++i) acc3[t3] = fmaf(v[i], w5[i], acc3[t3]);
                    }
                }
            }
            if (!store) { part[u][r][0] = acc3[0]; part[u][r][1] = acc3[1]; part[u][r][2] = acc3[2]; }
        };
        pc.mark(5);
        // ================= colour MLP =================
        fill_cin(1);                                              // gap columns of R1: free, no need to wait for the feature part
        wait_done();                                              // lin0, feature part done
        wait_done();                                              // lin0 complete, D in R0
        relu_epilogue(256, 0, true);
        lp0 = prm + P_CB + 256;
        wait_done();                                              // lin1, D in R1
        relu_epilogue(256, 1, true);
        lp0 = prm + P_CB + 512;
        wait_done();                                              // lin2 (N = 128), D in R0[0..127]
        relu_epilogue(128, 0, true);                              // -> operand chunks 0, 1 of R0
        fill_cin(0);                                              // gap columns of chunk 0: this warp's own, already consumed accumulators
        lp0 = prm + P_CB + 640;
        wait_done();                                              // lin3, lin2-output part done -> R0's chunks may be overwritten
        feat_refill(0, true);
        wait_done();                                              // lin3, feature part done
        wait_done();                                              // lin3 complete, D in R1
        relu_epilogue(256, 1, true);
        lp0 = prm + P_CB + 896;
        wait_done();                                              // lin4, D in R0
        relu_epilogue(256, 0, false);                             // lin5 (256 -> 3) folded into the epilogue
        sync_c();
        if (tid < UM && sl >= 0) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
                w.smp_rgb[3 * (size_t)sl + j] = sigmoid_(((part[0][tid][j] + part[1][tid][j]) + (part[2][tid][j] + part[3][tid][j])) + __ldg(tc.col_b[5] + j));
        }
        sync_c();
        pc.mark(6);
    }
    tc_fence_before();
    asm volatile("bar.sync 2, 544;" ::: "memory");
}

}  // namespace arah
