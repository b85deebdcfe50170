// arah_sdf3x.cuh — the 256-wide FiLM-SIREN SDF on tcgen05 in split precision (3xTF32 ~ fp32) for the ROOT-FINDING
// kernels of round 1 that are still in use (k_trace_tc3: sphere tracing when the vertex index does not fit next to the persistent
// kernel's weight ring; k_iso_init_tc3: Jacobian initialisation of the joint search), engine v3 roles (8 epilogue warps, TMA producer warp,
// MMA-issuer warp, per-chunk ready barriers).
//
// A 256-wide layer needs A_hi (256 cols) + A_lo (256 cols) + D (256 cols) = 768 TMEM columns, TMEM has 512.  So:
//   A_hi : TMEM, ping-pongs with D between the two 256-column regions (in-place D -> A_hi),
//   A_lo : shared memory, 8 K-chunks x 16 KB, SWIZZLE_128B K-major (the SS form of tcgen05.mma reads it),
//   B    : per K-chunk two 32 KB images [B_hi], [B_lo] through a 3-slot ring,
//   D   += A_lo(smem).B_hi + A_hi(tmem).B_hi        (when B_hi(c) has landed)
//   D   += A_hi(tmem).B_lo                          (when B_lo(c) has landed).
// Shared memory is exactly full (128 KB + 96 KB + 3 KB), so the dynamic segment must start 1024-byte aligned (checked).
// The activation math is the fp32 FFMA kernels' own formula sin(30 (f (acc + b) + phi)) with a 1-2 ulp Cody-Waite sine
// (sin_cw): root finding keeps resolving 1e-5 m.
#pragma once
#include "arah_tc2.cuh"
#include "arah_work.cuh"

namespace arah {

constexpr int TC3_THREADS = 320;      // 8 epilogue warps + TMA producer warp + MMA-issuer warp
// order in which the K-chunks of a 256- (order 0) / 128-wide (order 1) operand become ready: the two column halves finish alternately
__device__ __forceinline__ int seg_chunk(int order, int i) {
    if (order == 0) return (i >> 1) + ((i & 1) << 2);
    if (order == 1) return (i >> 1) + ((i & 1) << 1);
    return i;
}

struct SkinTC {
    const float* Wt0;      // [3][128]
    const float* b[5];     // biases (b[4] padded to 32)
    const float* hid[3];   // layers 1..3: 4 chunks of [hi | lo] images, N = 128
    const float* out;      // layer 4: 4 chunks of [hi | lo] images, N = 32 (25 padded)
};

constexpr int LGS = 33;     // row stride of the logits staging tile: odd, so that row-per-lane accesses hit 32 different banks

struct SdfTC {
    const float* Wt0;        // [3][256]
    const float* b[6];
    const float* freq;       // [6][256]
    const float* phase;      // [6][256]
    const float* hid[5];     // layers 1..5: 8 chunks x [hi 32 KB | lo 32 KB]
    const float* w6;
    const float* b6;         // [1] device scalar
};

constexpr int S3_NSLOTS = 3;
constexpr int S3_ALO_FLOATS = 8 * A_CHUNK_FLOATS;              // 128 KB
constexpr int S3_RING_FLOATS = S3_NSLOTS * RING_SLOT_FLOATS;   // 96 KB

struct S3Bars { uint64_t* full; uint64_t* empty; uint64_t* ready; uint64_t* done; };

// 32 consecutive per-column parameters as 8 LDG.128 (the L1 is almost entirely carved out as shared memory in these kernels)
__device__ __forceinline__ void ldg32(const float* __restrict__ p, float (&v)[32]) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float4 t = __ldg(q + j); v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w; }
}

// ---- producer: one tile's worth of SDF weight items (5 layers x 8 chunks x {hi, lo}) ----------------------------------
__device__ __forceinline__ void s3_produce_item(float* ring, const S3Bars& bar, uint32_t& slot, uint32_t& use, const void* src, uint32_t bytes) {
    if (use > 0) mbar_wait(&bar.empty[slot], (use - 1) & 1u);
    mbar_expect_tx(&bar.full[slot], bytes);
    bulk_g2s(ring + slot * RING_SLOT_FLOATS, src, bytes, &bar.full[slot]);
    if (++slot == S3_NSLOTS) { slot = 0; ++use; }
}
__device__ __forceinline__ void s3_produce_sdf(float* ring, const S3Bars& bar, uint32_t& slot, uint32_t& use, const SdfTC& sd) {
    for (int l = 0; l < 5; ++l)
        for (int i = 0; i < 8; ++i) {
            const int c = seg_chunk(0, i);
            const char* p = reinterpret_cast<const char*>(sd.hid[l]) + (size_t)c * 65536;
            s3_produce_item(ring, bar, slot, use, p, 32768u);
            s3_produce_item(ring, bar, slot, use, p + 32768, 32768u);
        }
}
// ---- MMA issuer: one tile's SDF layers -----------------------------------------------------------------------------------
__device__ __forceinline__ void s3_mma_sdf(float* ring, const float* A_lo, const S3Bars& bar, uint32_t& slot, uint32_t& use, uint32_t& rpar, uint32_t tbase) {
    const uint32_t idesc = umma_idesc_tf32(UM, 256);
    for (int L = 1; L <= 5; ++L) {
        const uint32_t ta = tbase + 256u * ((L - 1) & 1), td = tbase + 256u * (L & 1);
        for (int i = 0; i < 8; ++i) {
            const int c = seg_chunk(0, i);
            mbar_wait(&bar.ready[c], (rpar >> c) & 1u);
            rpar ^= (1u << c);
            mbar_wait(&bar.full[slot], use & 1u);                 // B_hi(c)
            tc_fence_after();
            uint32_t b_addr = smem_u32(ring + slot * RING_SLOT_FLOATS);
            const uint32_t al = smem_u32(A_lo + c * A_CHUNK_FLOATS);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                umma_tf32(td, umma_smem_desc_sw128(al + k * 32), umma_smem_desc_sw128(b_addr + k * 32), idesc, (i > 0 || k > 0) ? 1u : 0u);   // A_lo . B_hi
                umma_tf32_ts(td, ta + (uint32_t)(c * UK + k * 8), umma_smem_desc_sw128(b_addr + k * 32), idesc, 1u);                            // A_hi . B_hi
            }
            umma_commit(&bar.empty[slot]);
            if (++slot == S3_NSLOTS) { slot = 0; ++use; }
            mbar_wait(&bar.full[slot], use & 1u);                 // B_lo(c)
            tc_fence_after();
            b_addr = smem_u32(ring + slot * RING_SLOT_FLOATS);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_tf32_ts(td, ta + (uint32_t)(c * UK + k * 8), umma_smem_desc_sw128(b_addr + k * 32), idesc, 1u);                            // A_hi . B_lo
            umma_commit(&bar.empty[slot]);
            if (++slot == S3_NSLOTS) { slot = 0; ++use; }
        }
        umma_commit(bar.done);
    }
}
// ---- compute warps: SDF of the 128 rows whose normalised points sit in xs[r][0..2]; returns this thread's partial of
//      w6 . h5 over its 128 columns (caller adds the two halves and b6).  Must be called by all 8 compute warps.
__device__ __forceinline__ float s3_compute_sdf(const SdfTC& sd, const float* xs3, float* A_lo, const S3Bars& bar, uint32_t& done_par, uint32_t tbase, PhaseClk* pc = nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, half = warp >> 2, r = 32 * q + lane;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    auto put = [&](int reg, int chunk, const float (&v)[32]) {
        float hi[32], lo[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { hi[i] = tf32_rn(v[i]); lo[i] = v[i] - hi[i]; }
        tmem_st32(trow + 256u * reg + 32u * chunk, hi);
        a_store_chunk(A_lo, r, chunk, lo);                         // rounds lo to TF32
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
                h[i] = sin_cw(30.0f * (pf[i] * (a + pb[i]) + pp[i]));
            }
            put(0, col0 / 32, h);
        }
    }
    if (pc) pc->mark(1);
    float dot = 0.f;
    for (int L = 1; L <= 5; ++L) {
        mbar_wait(bar.done, done_par);
        done_par ^= 1u;
        __syncwarp();
        tc_fence_after();
        if (pc) pc->mark(2);
        const int dreg = L & 1;
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            const int col0 = 128 * half + 32 * b;
            float v[32], pf[32], pb[32], pp[32];
            ldg32(sd.freq + L * 256 + col0, pf); ldg32(sd.b[L] + col0, pb); ldg32(sd.phase + L * 256 + col0, pp);
            tmem_ld32(trow + 256u * dreg + (uint32_t)col0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = sin_cw(30.0f * (pf[i] * (v[i] + pb[i]) + pp[i]));
            if (L < 5) put(dreg, col0 / 32, v);
            else {
                ldg32(sd.w6 + col0, pf);
#pragma unroll
                for (int i = 0; i < 32; ++i) dot = fmaf(v[i], pf[i], dot);
            }
        }
        if (pc) pc->mark(3);
    }
    return dot;
}

// =====================================================================================================================
// k_trace_tc3: one sphere-tracing step (ray_tracing.py:198-241) for the active rays, 128 rays per tile.
__host__ __device__ constexpr size_t trace_tc3_smem_bytes() { return (size_t)(S3_ALO_FLOATS + S3_RING_FLOATS + UM * 3 + 2 * UM) * 4 + 160; }

__global__ void __launch_bounds__(TC3_THREADS, 1) k_trace_tc3(FrameParams fp, SdfTC sd, Work w, int iter) {
    extern __shared__ __align__(1024) uint8_t raw_smem[];
    const int n = w.counters[C_TRACE + iter];
    if ((int)blockIdx.x * UM >= n) return;
    if (smem_u32(raw_smem) & 1023u) __trap();                       // SWIZZLE_128B tiles need the segment 1024-aligned
    float* A_lo = reinterpret_cast<float*>(raw_smem);
    float* ring = A_lo + S3_ALO_FLOATS;
    float* xs3 = ring + S3_RING_FLOATS;                             // [128][3]
    float (*part)[UM] = reinterpret_cast<float (*)[UM]>(xs3 + UM * 3);   // [2][128]
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
    const int ntiles = (n + UM - 1) / UM;
    if (warp == 8) {
        if (lane == 0) { uint32_t slot = 0, use = 0; for (int t = blockIdx.x; t < ntiles; t += gridDim.x) s3_produce_sdf(ring, bar, slot, use, sd); }
        return;
    }
    if (warp == 9) {
        if (lane == 0) { uint32_t slot = 0, use = 0, rpar = 0; for (int t = blockIdx.x; t < ntiles; t += gridDim.x) s3_mma_sdf(ring, A_lo, bar, slot, use, rpar, tbase); }
        return;
    }
    const int half = warp >> 2, r = 32 * (warp & 3) + lane;
    uint32_t done_par = 0;
    const int* list = (iter & 1) ? w.listB : w.listA;
    int* next = (iter & 1) ? w.listA : w.listB;
    PhaseClk pc; pc.start((tid == 32 && w.phase_clk) ? w.phase_clk + 16 : nullptr);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int ray = -1;
        if (tid < UM) {
            const int i = tile * UM + tid;
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) { ray = list[i]; const RayCur& c = w.ray_cur[ray]; xn[0] = c.xn[0]; xn[1] = c.xn[1]; xn[2] = c.xn[2]; }
            xs3[3 * tid] = xn[0]; xs3[3 * tid + 1] = xn[1]; xs3[3 * tid + 2] = xn[2];
        }
        cta_sync_compute();
        pc.mark(0);
        const float dot = s3_compute_sdf(sd, xs3, A_lo, bar, done_par, tbase, &pc);
        part[half][r] = dot;
        cta_sync_compute();
        if (tid < UM) {                                             // marching logic, identical to k_trace_iter
            bool still = false;
            if (ray >= 0) {
                const float sdf = sdf_to_metres(part[0][tid] + part[1][tid] + __ldg(sd.b6), fp.cmin, fp.cmax);
                float t = w.ray_t[ray];
                const float far_ = w.near_far[2 * ray + 1];
                const float sm = fminf(fmaxf(sdf, -0.1f), 0.1f);
                bool diverge = false;
                if (fabsf(sm) > CVG_THRESH && fabsf(sdf) < 1e6f) { t = t + sm; diverge = t >= far_; w.ray_t[ray] = t; }
                still = !(fabsf(sdf) <= CVG_THRESH || diverge);
                w.ray_flags[ray] = (still ? 1 : 0) | (diverge ? 2 : 0);
            }
            if (iter + 1 < TRACE_ITERS) warp_append(still, ray, next, &w.counters[C_TRACE + iter + 1]);
            warp_stat_add(ray >= 0 ? 1 : 0, &w.counters[C_STAT_TRACE_EVALS]);
        }
        cta_sync_compute();
        pc.mark(4);
    }
    tc_fence_before();
    cta_sync_compute();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// ---- skinning MLP (3xTF32) on the same 3-slot ring / barrier set, for the joint-search kernel -----------------------------
// TMEM: X hi [0,128) | lo [128,256), accumulators ping-pong Da = [256,384) / Db = [384,512) (as k_corr_tc3).
__device__ __forceinline__ void s3_produce_skin(float* ring, const S3Bars& bar, uint32_t& slot, uint32_t& use, const SkinTC& sk) {
    for (int s = 0; s < 4; ++s) {
        const char* wsrc = reinterpret_cast<const char*>((s < 3) ? sk.hid[s] : sk.out);
        const uint32_t bytes = (s < 3) ? 32768u : 8192u;
        for (int i = 0; i < 4; ++i) s3_produce_item(ring, bar, slot, use, wsrc + (size_t)seg_chunk(1, i) * bytes, bytes);
    }
}
__device__ __forceinline__ void s3_mma_skin(float* ring, const S3Bars& bar, uint32_t& slot, uint32_t& use, uint32_t& rpar, uint32_t tbase) {
    for (int s = 0; s < 4; ++s) {
        const int N = (s < 3) ? 128 : 32;
        const uint32_t idesc = umma_idesc_tf32(UM, N);
        const uint32_t td = tbase + ((s & 1) ? 384u : 256u);
        for (int i = 0; i < 4; ++i) {
            const int c = seg_chunk(1, i);
            mbar_wait(&bar.ready[c], (rpar >> c) & 1u);
            rpar ^= (1u << c);
            mbar_wait(&bar.full[slot], use & 1u);
            tc_fence_after();
            const uint32_t bh = smem_u32(ring + slot * RING_SLOT_FLOATS), bl = bh + (uint32_t)N * UK * 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t col = (uint32_t)(c * UK + k * 8), ko = k * 32;
                umma_tf32_ts(td, tbase + 128u + col, umma_smem_desc_sw128(bh + ko), idesc, (i > 0 || k > 0) ? 1u : 0u);
                umma_tf32_ts(td, tbase + col, umma_smem_desc_sw128(bl + ko), idesc, 1u);
                umma_tf32_ts(td, tbase + col, umma_smem_desc_sw128(bh + ko), idesc, 1u);
            }
            umma_commit(&bar.empty[slot]);
            if (++slot == S3_NSLOTS) { slot = 0; ++use; }
        }
        umma_commit(bar.done);
    }
}

}  // namespace arah
