// arah_corr_p.cuh — k_corr_persist: the per-sample correspondence search (search_canonical_corr,
// /root/reference/im2mesh/utils/root_finding_utils.py:267-362, with broyden.py:4-78) as ONE persistent kernel.
//
// Round 1 launched one kernel per Broyden iteration (51 launches); every launch streamed each active sample's 160-byte
// BroydenState through HBM and re-compacted the active list.  Here a CTA (one per SM) owns two resident tiles of 128 samples
// whose Broyden state never leaves shared memory: every row runs its own iteration count (<= 50, broyden.py:47) and a row that
// finishes is re-filled at once from a device-wide queue of on-samples (blocks of 128 claimed with one atomicAdd), so the tiles
// stay full until the queue is empty.  Per-row arithmetic is the per-point restatement of broyden.py that k_corr_step /
// k_corr_tc5 ran (broyden_begin / broyden_update / corr_finalize): a row's result does not depend on its tile mates.
//
// Skinning MLP (metaavatar/models/decoder.py:201-233; 3 -> 128 -> 128 -> 128 -> 128 -> 25, Softplus(beta = 100)):
//   layer 0 (K = 3) on the FP32 pipe; layers 1..4 on tcgen05 kind::f16 in split precision (arah_f16x3.cuh), three products per
//   K-step: X_lo.B_hi + X_hi.B_lo + X_hi.B_hi, fp32 accumulation in TMEM.
// TMEM (512 columns): tile T at 256 T: X_hi [0,64) | X_lo [64,128) (two K values per column) | D [128,256).
// Shared memory: the hi weight images of all four layers stay RESIDENT (104 KB, loaded once per CTA); only the lo images
// stream from L2, one layer (32 KB) at a time through a 2-slot ring, and a streamed layer serves BOTH tiles.
// Warps: 0-7 tile A, 8-15 tile B (q = warp & 3 -> TMEM lanes 32 q.., h = (warp >> 2) & 1 -> column half), warp 16 = one elected
// thread that issues every TMA copy and every MMA.  While one tile's MMAs run, the other tile's 8 warps run their epilogue
// (bias + softplus + hi/lo split, MUFU-bound), so the two tiles alternate on the tensor pipe and on the MUFU pipe.
#pragma once
#include "arah_f16x3.cuh"
#include "arah_work.cuh"

namespace arah {

struct SkinF16 {
    const float* Wt0;              // [3][128] fp32
    const float* b[5];             // biases (b[4] padded to 32)
    const __half* hi;              // hi images: layers 1..3 (2 chunks x 128 x 64 each = 32 KB) then layer 4 (2 x 32 x 64 = 8 KB)
    const __half* lo;              // lo images, same layout
    const float* scale;            // [4][2]: (s, 1 / s) of layers 1..4
};

constexpr int CP_THREADS = 544;                       // 2 x 8 tile warps + the TMA / MMA warp
constexpr int CP_HI_BYTES = 3 * 32768 + 8192;         // resident hi images
constexpr int CP_SLOT_BYTES = 32768;                  // one layer's lo image
constexpr int CP_STATE_WORDS = 40;                    // per row, SoA: word f of row r at st[f * 128 + r]
// state words
enum { CS_X = 0, CS_J = 3, CS_GX = 12, CS_DX = 15, CS_BX = 18, CS_BT = 21, CS_BN = 33, CS_TG = 34, CS_OWNER = 37, CS_IT = 38, CS_EV = 39 };
constexpr int CP_IDLE = -2, CP_FRESH = -1;            // CS_IT: no work / first evaluation pending / index of the pending iteration

__host__ __device__ constexpr size_t corr_persist_smem_bytes() {
    return (size_t)CP_HI_BYTES + 2 * CP_SLOT_BYTES + 2 * CP_STATE_WORDS * UM * 4 + 2 * UM * 4 * 4 + (24 * 16 + 3 * 128 + 5 * 128 + 8) * 4 + 256 + 1024;
}

__device__ __forceinline__ void named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ bool named_sync_or(int id, int nthreads, bool pred) {
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"((uint32_t)pred), "r"(id), "r"(nthreads) : "memory");
    return r != 0;
}

// ---- per-point phase helpers: the same formulas as arah_math.cuh (hierarchical_softmax, blend_T) with the MUFU forms of exp and
// 1 / x (relative error ~2^-22, the noise floor of the split-precision MLP that produced the logits) and 128-bit loads of the bone
// transforms; this phase sits on the critical path of a tile's iteration.
__device__ __forceinline__ float fast_ex2(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sigmoid(float x) { return fast_rcp(1.0f + fast_ex2(x * -1.4426950408889634f)); }
__device__ __forceinline__ void fast_softmax3(const float* x, float* y) {
    const float m = fmaxf(x[0], fmaxf(x[1], x[2]));
    const float e0 = fast_ex2((x[0] - m) * 1.4426950408889634f), e1 = fast_ex2((x[1] - m) * 1.4426950408889634f),
                e2 = fast_ex2((x[2] - m) * 1.4426950408889634f);
    const float inv = fast_rcp(e0 + e1 + e2);
    y[0] = e0 * inv; y[1] = e1 * inv; y[2] = e2 * inv;
}
// x: 25 logits already multiplied by 20 -> p: 24 weights (utils/utils.py:138-181)
__device__ __forceinline__ void hierarchical_softmax_fast(const float* x, float* p) {
    float sm[3];
    fast_softmax3(x + 1, sm);
    const float s0 = fast_sigmoid(x[0]);
    p[0] = 1.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) p[1 + k] = p[0] * s0 * sm[k];
    p[0] = p[0] * (1.0f - s0);
#define ARAH_SPLITF(c, q, g) { const float s_ = fast_sigmoid(x[g]); p[c] = p[q] * s_; p[q] = p[q] * (1.0f - s_); }
    ARAH_SPLITF(4, 1, 4) ARAH_SPLITF(5, 2, 5) ARAH_SPLITF(6, 3, 6)
    ARAH_SPLITF(7, 4, 7) ARAH_SPLITF(8, 5, 8) ARAH_SPLITF(9, 6, 9)
    ARAH_SPLITF(10, 7, 10) ARAH_SPLITF(11, 8, 11)
    fast_softmax3(x + 12, sm);
    {
        const float s24 = fast_sigmoid(x[24]), p9 = p[9];
#pragma unroll
        for (int k = 0; k < 3; ++k) p[12 + k] = p9 * s24 * sm[k];
        p[9] = p9 * (1.0f - s24);
    }
    ARAH_SPLITF(15, 12, 15)
    ARAH_SPLITF(16, 13, 16) ARAH_SPLITF(17, 14, 17)
    ARAH_SPLITF(18, 16, 18) ARAH_SPLITF(19, 17, 19)
    ARAH_SPLITF(20, 18, 20) ARAH_SPLITF(21, 19, 21)
    ARAH_SPLITF(22, 20, 22) ARAH_SPLITF(23, 21, 23)
#undef ARAH_SPLITF
}
// T12 = sum_j w_j B_j[:3,:] with B in shared memory ([24][16] floats, 16-byte aligned): same accumulation order as blend_T
__device__ __forceinline__ void blend_T_smem(const float* w, const float* sB, float* T12) {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float4* B = reinterpret_cast<const float4*>(sB + j * 16);
        const float4 b0 = B[0], b1 = B[1], b2 = B[2];
        const float wj = w[j];
        a0.x += wj * b0.x; a0.y += wj * b0.y; a0.z += wj * b0.z; a0.w += wj * b0.w;
        a1.x += wj * b1.x; a1.y += wj * b1.y; a1.z += wj * b1.z; a1.w += wj * b1.w;
        a2.x += wj * b2.x; a2.y += wj * b2.y; a2.z += wj * b2.z; a2.w += wj * b2.w;
    }
    T12[0] = a0.x; T12[1] = a0.y; T12[2] = a0.z; T12[3] = a0.w;
    T12[4] = a1.x; T12[5] = a1.y; T12[6] = a1.z; T12[7] = a1.w;
    T12[8] = a2.x; T12[9] = a2.y; T12[10] = a2.z; T12[11] = a2.w;
}

__global__ void __launch_bounds__(CP_THREADS, 1) k_corr_persist(FrameParams fp, SkinF16 sk, Work w) {
    extern __shared__ uint8_t raw_smem[];
    const int n_on = w.counters[C_ON];
    if (n_on <= 0) return;
    const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;
    uint8_t* sm = raw_smem + (base - smem_u32(raw_smem));
    uint8_t* sHi = sm;                                                   // resident hi images (1024-aligned)
    uint8_t* sRing = sHi + CP_HI_BYTES;                                  // 2 x 32 KB lo slots (1024-aligned)
    float* sState = reinterpret_cast<float*>(sRing + 2 * CP_SLOT_BYTES); // [2][40][128]
    float* sXs = sState + 2 * CP_STATE_WORDS * UM;                       // [2][128][4] normalised query points
    float* sB = sXs + 2 * UM * 4;                                        // bone transforms [24][16]
    float* sW0 = sB + 24 * 16;                                           // layer-0 weights [3][128]
    float* sb = sW0 + 3 * 128;                                           // biases 4 x 128, then 32
    float* sInv = sb + 5 * 128;                                          // 1 / scale of layers 1..4
    uint64_t* bars = reinterpret_cast<uint64_t*>(sInv + 8);
    uint64_t* full = bars;              // [2] lo slot landed
    uint64_t* empty = bars + 2;         // [2] the MMAs that read the slot completed
    uint64_t* ready = bars + 4;         // [2] tile T: operand X written by its 8 warps (and its D consumed)
    uint64_t* done = bars + 6;          // [2] tile T: the job's accumulators are complete
    uint64_t* wres = bars + 8;          // resident images landed
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 9);
    volatile int* tile_dead = reinterpret_cast<volatile int*>(tslot + 1);        // [2]
    int* qstate = const_cast<int*>(tile_dead) + 2;                       // per tile [12]: blk_base[8], pos, nclaimed, exhausted, -
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&ready[i], 8); mbar_init(&done[i], 1); }
        mbar_init(wres, 1);
        mbar_fence_init();
        tile_dead[0] = 0; tile_dead[1] = 0;
        for (int i = 0; i < 24; ++i) qstate[i] = 0;
    }
    if (warp == 16) tmem_alloc(tslot, 512);
    for (int i = tid; i < 24 * 16; i += CP_THREADS) sB[i] = __ldg(fp.bone_T + i);
    for (int i = tid; i < 3 * 128; i += CP_THREADS) sW0[i] = __ldg(sk.Wt0 + i);
    for (int i = tid; i < 4 * 128; i += CP_THREADS) sb[i] = __ldg(sk.b[i >> 7] + (i & 127));
    if (tid < 32) sb[512 + tid] = __ldg(sk.b[4] + tid);
    if (tid < 4) sInv[tid] = __ldg(sk.scale + 2 * tid + 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;

    if (warp == 16) {                                   // ===== TMA + MMA thread =====
        if (lane == 0) {
            const char* glo = reinterpret_cast<const char*>(sk.lo);
            auto lo_bytes = [](int s) -> uint32_t { return s < 3 ? 32768u : 8192u; };
            auto fetch = [&](uint32_t g) {             // stage g (layer 1 + g % 4) -> slot g & 1
                const int s = (int)(g & 3u), slot = (int)(g & 1u);
                mbar_expect_tx(&full[slot], lo_bytes(s));
                bulk_g2s(sRing + slot * CP_SLOT_BYTES, glo + (size_t)s * 32768, lo_bytes(s), &full[slot]);
            };
            mbar_expect_tx(wres, (uint32_t)CP_HI_BYTES);
            for (int s = 0; s < 4; ++s) bulk_g2s(sHi + s * 32768, reinterpret_cast<const char*>(sk.hi) + (size_t)s * 32768, lo_bytes(s), wres);
            fetch(0);
            uint32_t fetched = 1;                       // stages fetched so far
            mbar_wait(wres, 0);
            bool alive[2] = {true, true};
            uint32_t rpar[2] = {0u, 0u};
            uint32_t g = 0;                             // stage counter
            while (alive[0] || alive[1]) {
                for (int s = 0; s < 4; ++s, ++g) {
                    const int slot = (int)(g & 1u);
                    // keep one stage in flight: stage g + 1 goes into the other slot once the MMAs of stage g - 1 have released it
                    if (fetched == g + 1) {
                        if (g >= 1) mbar_wait(&empty[slot ^ 1], ((g - 1) >> 1) & 1u);
                        fetch(g + 1);
                        ++fetched;
                    }
                    const int N = (s < 3) ? 128 : 32;
                    const uint32_t idesc = umma_idesc_f16(UM, N), img = (uint32_t)N * HK * 2;
                    const uint32_t bh0 = smem_u32(sHi + s * 32768), bl0 = smem_u32(sRing + slot * CP_SLOT_BYTES);
                    bool waited = false, any = false;
                    for (int t = 0; t < 2; ++t) {
                        if (!alive[t]) continue;
                        mbar_wait(&ready[t], rpar[t]);
                        rpar[t] ^= 1u;
                        if (s == 0 && tile_dead[t]) { alive[t] = false; continue; }
                        if (!waited) { mbar_wait(&full[slot], (g >> 1) & 1u); waited = true; }
                        tc_fence_after();
                        const uint32_t tb = tbase + 256u * t;
#pragma unroll
                        for (int kc = 0; kc < 2; ++kc)
                            umma_f16x3_chunk(tb + 128u, tb + 32u * kc, tb + 64u + 32u * kc, bh0 + kc * img, bl0 + kc * img, idesc, kc == 0);
                        umma_commit(&done[t]);
                        any = true;
                    }
                    if (!any) break;                    // both tiles retired (only possible at s == 0)
                    umma_commit(&empty[slot]);
                }
            }
            // drain: every fetched stage must have landed before the CTA may exit; stages [g, fetched) were never consumed
            for (uint32_t f = g; f < fetched; ++f) mbar_wait(&full[f & 1u], (f >> 1) & 1u);
        }
        __syncwarp();
        named_sync(5, CP_THREADS);
        tmem_dealloc(tbase, 512);
        return;
    }

    // ===== tile engines =====
    const int T = warp >> 3;                                            // tile of this warp
    const int q = warp & 3, h = (warp >> 2) & 1, r = 32 * q + lane;     // TMEM lane quarter, column half, row
    const int bar_tile = 1 + T;
    float* st = sState + T * CP_STATE_WORDS * UM;
    float (*xs)[4] = reinterpret_cast<float (*)[4]>(sXs + T * UM * 4);
    int* qs = qstate + 12 * T;                                          // blk_base[8], pos, nclaimed, exhausted
    const uint32_t tb = tbase + 256u * T + ((uint32_t)(32 * q) << 16);  // this thread's TMEM row, tile base column
    uint32_t done_par = 0;
    int evals = 0;
    PhaseClk pc; pc.start((tid == 0) ? w.phase_clk : nullptr);
    auto wait_done = [&]() { mbar_wait(&done[T], done_par); done_par ^= 1u; __syncwarp(); tc_fence_after(); };
    auto publish = [&]() {                              // this warp's part of X is in TMEM (and its reads of D are complete)
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[T]);
    };
    // softplus(acc * inv + bias) of 32 accumulator columns -> 16 + 16 packed operand columns
    auto emit = [&](float (&v)[32], int xcol) {
        uint32_t hi[16], lo[16];
        split_pack_f16(v, hi, lo);
        tmem_st16(tb + (uint32_t)xcol, hi);
        tmem_st16(tb + 64u + (uint32_t)xcol, lo);
    };
    auto layer0 = [&]() {
        const float x = xs[r][0], y = xs[r][1], z = xs[r][2];
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
            const int col0 = 64 * h + 32 * b;
            float v[32];
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {                              // 128-bit shared-memory loads (all carve-outs are 16-byte aligned)
                const int cc = col0 + 4 * i4;
                const float4 wx = *reinterpret_cast<const float4*>(sW0 + cc), wy = *reinterpret_cast<const float4*>(sW0 + 128 + cc),
                             wz = *reinterpret_cast<const float4*>(sW0 + 256 + cc), bb = *reinterpret_cast<const float4*>(sb + cc);
                v[4 * i4 + 0] = softplus100_fast(fmaf(wz.x, z, fmaf(wy.x, y, wx.x * x)) + bb.x);
                v[4 * i4 + 1] = softplus100_fast(fmaf(wz.y, z, fmaf(wy.y, y, wx.y * x)) + bb.y);
                v[4 * i4 + 2] = softplus100_fast(fmaf(wz.z, z, fmaf(wy.z, y, wx.z * x)) + bb.z);
                v[4 * i4 + 3] = softplus100_fast(fmaf(wz.w, z, fmaf(wy.w, y, wx.w * x)) + bb.w);
            }
            emit(v, col0 / 2);
        }
        publish();
    };
    // Work queue: blocks of 128 on-samples are claimed with one atomicAdd on a device-wide cursor by the tile's helper warp
    // (tile-local warp 4, idle during the per-point phase) so that at least three full blocks always lie ahead of the position
    // the row warps consume from; the seeds of a claimed block are prefetched into L2.  Claims made during a phase are
    // published by the tile barrier that ends it.  blk ring: 8 slots, at most 6 blocks are alive at a time.
    auto claim_ahead = [&]() {                          // all 32 lanes of the helper warp
        int nclaimed = qs[9];
        const int pos = *reinterpret_cast<volatile int*>(&qs[8]);
        while (nclaimed * UM - pos < 5 * UM) {
            int b = -1;
            if (lane == 0 && !qs[10]) {
                b = atomicAdd(&w.counters[C_CORR_CURSOR], UM);
                if (b >= n_on) { b = -1; qs[10] = 1; }
            }
            b = __shfl_sync(0xffffffffu, b, 0);
            if (b >= 0) {
                const char* p0 = reinterpret_cast<const char*>(w.corr_seed + b);
                for (int l = lane; l < (int)(UM * sizeof(CorrSeed) / 128); l += 32)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + (size_t)l * 128));
            }
            if (lane == 0) qs[nclaimed & 7] = b;
            ++nclaimed;
        }
        if (lane == 0) qs[9] = nclaimed;
        __syncwarp();
    };
    // ---- start: every row is idle and asks for a sample
    if (h == 0) st[CS_IT * UM + r] = __int_as_float(CP_IDLE);
    if (h == 1 && q == 0) claim_ahead();
    named_sync(bar_tile, 2 * UM);
    bool first_round = true;
    while (true) {
        // =================== per-point phase (the 128 threads with h == 0 own one row each) ===================
        // (Measured: letting both warps of a lane quarter take 16 rows each — two half-filled warps to hide the latency of this
        // dependent chain — doubles the issued instructions next to the other tile's epilogue warps and costs 2 ms.  Reverted.)
        if (h == 0) {
            int it = __float_as_int(st[CS_IT * UM + r]);
            bool need = first_round;                                    // row wants a new sample
            float lgv[32];
            if (!first_round) {                                         // warp-uniform: tcgen05.ld is a warp-collective instruction
                tmem_ld32(tb + 128u, lgv);
                tc_fence_before();
            }
            if (!first_round && it != CP_IDLE) {
                float T12[12], g[3];
                BroydenState<3> s;
                {
                    float lg[25], wj[NJ], xb[3];
                    const float inv4 = sInv[3];
#pragma unroll
                    for (int k = 0; k < 25; ++k) lg[k] = fmaf(lgv[k], inv4, sb[512 + k]) * 20.0f;
                    hierarchical_softmax_fast(lg, wj);
                    blend_T_smem(wj, sB, T12);
#pragma unroll
                    for (int k = 0; k < 3; ++k) s.x[k] = st[(CS_X + k) * UM + r];
                    apply_T(T12, s.x, xb);
#pragma unroll
                    for (int k = 0; k < 3; ++k) { s.tgt[k] = st[(CS_TG + k) * UM + r]; g[k] = xb[k] - s.tgt[k]; }
                }
                bool active;
                if (it == CP_FRESH) {
                    // initial evaluation: J^-1 from the blended transform at x0 (root_finding_utils.py:327-328), g(x0), first update
                    float A3[9], Ai[9], Tinit[12];
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                        for (int c = 0; c < 3; ++c) A3[rr * 3 + c] = T12[rr * 4 + c];
                    invert3(A3, Ai);
#pragma unroll
                    for (int e = 0; e < 12; ++e) Tinit[e] = st[(CS_BT + e) * UM + r];
                    const float x0[3] = {s.x[0], s.x[1], s.x[2]};
                    broyden_begin<3>(s, x0, g, Ai, Tinit);
                    s.g_evals = 2;                                       // the reference evaluates g twice here
                    active = true;                                       // every point takes at least one step (broyden.py:45)
                    it = 0;
#pragma unroll
                    for (int k = 0; k < 9; ++k) st[(CS_J + k) * UM + r] = s.Jinv[k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { st[(CS_GX + k) * UM + r] = s.gx[k]; st[(CS_BX + k) * UM + r] = s.best_x[k]; }
                    st[CS_BN * UM + r] = s.best_n;
                } else {
                    float dx[3];
#pragma unroll
                    for (int k = 0; k < 9; ++k) s.Jinv[k] = st[(CS_J + k) * UM + r];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { s.gx[k] = st[(CS_GX + k) * UM + r]; dx[k] = st[(CS_DX + k) * UM + r]; s.best_x[k] = st[(CS_BX + k) * UM + r]; }
#pragma unroll
                    for (int e = 0; e < 12; ++e) s.best_T[e] = st[(CS_BT + e) * UM + r];
                    s.best_n = st[CS_BN * UM + r];
                    s.g_evals = __float_as_int(st[CS_EV * UM + r]);
                    active = broyden_update<3>(s, dx, g, T12);
                    if (it + 1 >= BROYDEN_ITERS) active = false;
                    ++it;
                    if (active) {
#pragma unroll
                        for (int k = 0; k < 9; ++k) st[(CS_J + k) * UM + r] = s.Jinv[k];
#pragma unroll
                        for (int k = 0; k < 3; ++k) { st[(CS_GX + k) * UM + r] = s.gx[k]; st[(CS_BX + k) * UM + r] = s.best_x[k]; }
#pragma unroll
                        for (int e = 0; e < 12; ++e) st[(CS_BT + e) * UM + r] = s.best_T[e];
                        st[CS_BN * UM + r] = s.best_n;
                    }
                }
                if (active) {
                    // next query: x + update (broyden.py:50-51); the applied step is kept for the rank-1 update
                    float xq[3], xn[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) { xq[k] = s.x[k] + s.upd[k]; st[(CS_X + k) * UM + r] = xq[k]; st[(CS_DX + k) * UM + r] = s.upd[k]; }
                    normalize3(fp, xq, xn);
                    xs[r][0] = xn[0]; xs[r][1] = xn[1]; xs[r][2] = xn[2];
                    st[CS_IT * UM + r] = __int_as_float(it);
                    st[CS_EV * UM + r] = __int_as_float(s.g_evals);
                } else {
                    s.owner = __float_as_int(st[CS_OWNER * UM + r]);
                    corr_finalize(fp, w, s);
                    evals += s.g_evals;
                    need = true;
                }
            }
            // ---- re-fill the rows that finished
            {
                const unsigned m = __ballot_sync(0xffffffffu, need);
                int p0 = 0;
                if (m && lane == (__ffs(m) - 1)) p0 = atomicAdd(&qs[8], __popc(m));
                p0 = __shfl_sync(0xffffffffu, p0, m ? (__ffs(m) - 1) : 0);
                if (need) {
                    const int p = p0 + __popc(m & ((1u << lane) - 1u));
                    const int bb = qs[(p >> 7) & 7];
                    const int idx = bb + (p & (UM - 1));
                    if (bb >= 0 && idx < n_on) {
                        const float4* sp = reinterpret_cast<const float4*>(w.corr_seed + idx);
                        const float4 a = __ldg(sp), t0 = __ldg(sp + 1), t1 = __ldg(sp + 2), t2 = __ldg(sp + 3), c = __ldg(sp + 4);
                        const float x0[3] = {a.x, a.y, a.z};
                        float xn[3];
                        st[(CS_X + 0) * UM + r] = a.x; st[(CS_X + 1) * UM + r] = a.y; st[(CS_X + 2) * UM + r] = a.z;
                        st[CS_OWNER * UM + r] = a.w;
                        st[(CS_BT + 0) * UM + r] = t0.x; st[(CS_BT + 1) * UM + r] = t0.y; st[(CS_BT + 2) * UM + r] = t0.z; st[(CS_BT + 3) * UM + r] = t0.w;
                        st[(CS_BT + 4) * UM + r] = t1.x; st[(CS_BT + 5) * UM + r] = t1.y; st[(CS_BT + 6) * UM + r] = t1.z; st[(CS_BT + 7) * UM + r] = t1.w;
                        st[(CS_BT + 8) * UM + r] = t2.x; st[(CS_BT + 9) * UM + r] = t2.y; st[(CS_BT + 10) * UM + r] = t2.z; st[(CS_BT + 11) * UM + r] = t2.w;
                        st[(CS_TG + 0) * UM + r] = c.x; st[(CS_TG + 1) * UM + r] = c.y; st[(CS_TG + 2) * UM + r] = c.z;
                        st[CS_IT * UM + r] = __int_as_float(CP_FRESH);
                        normalize3(fp, x0, xn);
                        xs[r][0] = xn[0]; xs[r][1] = xn[1]; xs[r][2] = xn[2];
                        need = false;
                    } else {
                        st[CS_IT * UM + r] = __int_as_float(CP_IDLE);
                        xs[r][0] = 0.f; xs[r][1] = 0.f; xs[r][2] = 0.f;
                    }
                }
            }
        }
        else if (q == 0) claim_ahead();                                  // helper warp: keep the block queue ahead
        first_round = false;
        pc.mark(5);
        // a tile lives while any of its rows has work (also publishes xs to the h == 1 warps)
        const bool row_live = (h == 0) && (__float_as_int(st[CS_IT * UM + r]) != CP_IDLE);
        const bool live = named_sync_or(bar_tile, 2 * UM, row_live);
        if (!live) {
            if (tid == T * 2 * UM) tile_dead[T] = 1;
            __threadfence_block();
            named_sync(bar_tile, 2 * UM);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[T]);                       // lets the MMA thread observe tile_dead
            break;
        }
        // =================== skinning MLP ===================
        layer0();
        pc.mark(1);
#pragma unroll 1
        for (int l = 1; l < 4; ++l) {
            wait_done();                                                 // accumulators of layer l (pre-activations, scaled by s_l)
            pc.mark(2);
            const float inv = sInv[l - 1];
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
                const int col0 = 64 * h + 32 * b;
                float v[32];
                tmem_ld32(tb + 128u + (uint32_t)col0, v);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 bb = *reinterpret_cast<const float4*>(sb + 128 * l + col0 + 4 * i4);
                    v[4 * i4 + 0] = softplus100_fast(fmaf(v[4 * i4 + 0], inv, bb.x));
                    v[4 * i4 + 1] = softplus100_fast(fmaf(v[4 * i4 + 1], inv, bb.y));
                    v[4 * i4 + 2] = softplus100_fast(fmaf(v[4 * i4 + 2], inv, bb.z));
                    v[4 * i4 + 3] = softplus100_fast(fmaf(v[4 * i4 + 3], inv, bb.w));
                }
                emit(v, col0 / 2);
            }
            publish();
            pc.mark(3);
        }
        wait_done();                                                     // logits of this round are in D[0, 32)
        pc.mark(4);
    }
    warp_stat_add(evals, &w.counters[C_STAT_CORR_EVALS]);
    named_sync(5, CP_THREADS);
}

// on-samples whose search converged -> shade_list (order: blocks in arrival order, ascending inside a block), C_SHADE = their number
__global__ void __launch_bounds__(1024) k_shade_compact(Work w) {
    __shared__ int wsum[32];
    __shared__ int blk_base;
    const int n = w.counters[C_ON];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i0 = blockIdx.x * 1024; i0 < n; i0 += gridDim.x * 1024) {
        const int i = i0 + tid;
        int sl = 0;
        bool ok = false;
        if (i < n) { sl = w.on_list[i]; ok = w.smp_conv[sl] != 0; }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) wsum[wid] = __popc(m);
        __syncthreads();
        if (wid == 0) {
            int v = wsum[lane], incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            wsum[lane] = incl - v;
            if (lane == 31) blk_base = incl ? atomicAdd(&w.counters[C_SHADE], incl) : 0;
        }
        __syncthreads();
        if (ok) w.shade_list[blk_base + wsum[wid] + __popc(m & ((1u << lane) - 1u))] = sl;
        __syncthreads();
    }
}

}  // namespace arah
