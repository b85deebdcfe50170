// arah_sdf16.cuh — the 256-wide FiLM-SIREN SDF (hyperlayers.py:385-415, siren_modules.py:31-37) on tcgen05 kind::f16 in split
// precision (arah_f16x3.cuh: three fp16 products ~ one fp32 product), the engine of the persistent root-finding kernels
// (sphere tracing, joint search) and of the canonical SDF lattice.
//
// Everything lives in tensor memory now.  TMEM = two 256-column regions R0 / R1.  The activations of layer L sit in one region
// as four K-chunks of 64 values: chunk j, K-step k -> columns [64 j + 16 k, +8) hi | [64 j + 16 k + 8, +8) lo (two K values per column);
// the accumulators of layer L go to the other region.  The epilogue turns D into the next layer's operand IN PLACE: sixteen
// compute warps, four per TMEM lane quarter; warp (q, u) reads accumulator columns [64 j + 16 u, +16) of chunk j and overwrites
// them with 8 hi + 8 lo operand columns (K-step u of chunk j), then arrives on the chunk's `ready` barrier; the MMA warp starts chunk j of the next layer as soon as ready[j] and the weight images are
// there and writes the new accumulators into the region the previous operand vacated: layer L's epilogue and layer L + 1's MMAs
// overlap chunk by chunk, exactly as in round 1's engine — but at twice the MMA rate, with half the weight bytes, and without
// the 128 KB shared-memory copy of A_lo, which frees the room for the 1-NN vertex index in the tracing kernel.
// Weights: per layer and K-chunk one hi and one lo image (256 rows x 64 K x 2 B = 32 KB each), streamed through a 3-slot ring by
// a producer warp.  D (+)= X_lo.B_hi + X_hi.B_hi when the hi image has landed, += X_hi.B_lo when the lo image has.
#pragma once
#include "arah_f16x3.cuh"
#include "arah_work.cuh"

namespace arah {

struct SdfF16 {
    const float* Wt0;        // [3][256]
    const float* b[6];
    const float* freq;       // [6][256]
    const float* phase;      // [6][256]
    const __half* hi;        // layers 1..5: 4 chunks x 256 x 64 halfs (128 KB per layer)
    const __half* lo;
    const float* scale;      // [5][2]: (s, 1 / s)
    const float* w6;
    const float* b6;         // [1] device scalar
};
constexpr size_t SDF_F16_IMAGE_BYTES = 5 * 131072;

constexpr int S16_NSLOTS = 3;
constexpr int S16_SLOT_BYTES = 32768;
constexpr int S16_WARPS = 16;            // compute warps: q = warp & 3 -> TMEM lane quarter, u = warp >> 2 -> 16 columns of every K-chunk
constexpr int S16_CTHREADS = 32 * S16_WARPS;
constexpr int S16_THREADS = S16_CTHREADS + 64;   // + producer warp (16) + MMA warp (17)

struct S16Ctl {                          // shared-memory control block
    uint64_t full[S16_NSLOTS], empty[S16_NSLOTS];
    uint64_t ready[4];                   // SDF operand chunk j written (one arrival per compute warp)
    uint64_t done;                       // a layer's accumulators are complete
    uint64_t go;                         // compute -> MMA warp: decision about the next evaluation is in cont[]
    uint64_t ready_sk;                   // skinning operand written by all compute warps (joint search only)
    uint32_t tslot;
    volatile int stop;                   // compute -> producer: no further evaluation
    volatile int cont[2];
};

__device__ __forceinline__ void s16_ctl_init(S16Ctl* c) {      // one thread
    for (int i = 0; i < S16_NSLOTS; ++i) { mbar_init(&c->full[i], 1); mbar_init(&c->empty[i], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&c->ready[i], S16_WARPS);
    mbar_init(&c->done, 1); mbar_init(&c->go, 1); mbar_init(&c->ready_sk, S16_WARPS);
    mbar_fence_init();
    c->stop = 0; c->cont[0] = 0; c->cont[1] = 0;
}
__device__ __forceinline__ int s16_chunk(int i) { return i; }                                  // every warp finishes chunk 0 first, then 1, 2, 3
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}

__device__ __forceinline__ void s16_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }      // the 16 compute warps
__device__ __forceinline__ bool s16_sync_or(bool pred) {                         // same barrier + OR of a predicate
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, 1, 512, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"((uint32_t)pred) : "memory");
    return r != 0;
}
__device__ __forceinline__ void s16_sync_exit() { asm volatile("bar.sync 2, 544;" ::: "memory"); } // compute warps + MMA warp: TMEM is idle

// ---- producer (one thread): ring position + what is still in flight -------------------------------------------------------
struct S16Prod {
    uint32_t slot = 0, use = 0, issued = 0;
};
// stage one weight image; with `spec` the wait for a free slot gives up when the compute warps have declared the end
__device__ __forceinline__ bool s16_put(uint8_t* ring, S16Ctl* c, S16Prod& p, const void* src, uint32_t bytes, bool spec) {
    if (p.use > 0) {
        if (spec) { while (!mbar_try(&c->empty[p.slot], (p.use - 1) & 1u)) if (c->stop) return false; }
        else mbar_wait(&c->empty[p.slot], (p.use - 1) & 1u);
    }
    if (spec && c->stop) return false;
    mbar_expect_tx(&c->full[p.slot], bytes);
    bulk_g2s(ring + p.slot * S16_SLOT_BYTES, src, bytes, &c->full[p.slot]);
    ++p.issued;
    if (++p.slot == S16_NSLOTS) { p.slot = 0; ++p.use; }
    return true;
}
__device__ __forceinline__ bool s16_produce_sdf(uint8_t* ring, S16Ctl* c, S16Prod& p, const SdfF16& sd, bool spec) {
    for (int L = 0; L < 5; ++L)
        for (int i = 0; i < 4; ++i) {
            const int ch = s16_chunk(i);
            const size_t off = (size_t)L * 131072 + (size_t)ch * 32768;
            if (!s16_put(ring, c, p, reinterpret_cast<const char*>(sd.hi) + off, 32768u, spec)) return false;
            if (!s16_put(ring, c, p, reinterpret_cast<const char*>(sd.lo) + off, 32768u, spec)) return false;
        }
    return true;
}
// before the producer thread may return: every copy it issued has landed (the last S16_NSLOTS at most can be in flight)
__device__ __forceinline__ void s16_drain(S16Ctl* c, const S16Prod& p) {
    const uint32_t n = p.issued < (uint32_t)S16_NSLOTS ? p.issued : (uint32_t)S16_NSLOTS;
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t idx = p.issued - 1 - k;
        mbar_wait(&c->full[idx % S16_NSLOTS], (idx / S16_NSLOTS) & 1u);
    }
}

// ---- MMA issuer (one thread) -------------------------------------------------------------------------------------------------
struct S16Mma { uint32_t slot = 0, use = 0, rpar = 0; };
__device__ __forceinline__ void s16_mma_sdf(uint8_t* ring, S16Ctl* c, S16Mma& m, uint32_t tbase) {
    const uint32_t idesc = umma_idesc_f16(UM, 256);
    for (int L = 1; L <= 5; ++L) {
        const uint32_t ta = tbase + 256u * ((L - 1) & 1), td = tbase + 256u * (L & 1);
        for (int i = 0; i < 4; ++i) {
            const int ch = s16_chunk(i);
            mbar_wait(&c->ready[ch], (m.rpar >> ch) & 1u);
            m.rpar ^= (1u << ch);
            const uint32_t xh = ta + 64u * ch, xl = xh + 8u;                     // K-step k: hi at xh + 16 k, lo 8 columns behind
            mbar_wait(&c->full[m.slot], m.use & 1u);                           // B_hi(ch)
            tc_fence_after();
            uint32_t b = smem_u32(ring + m.slot * S16_SLOT_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                umma_f16_ts(td, xl + 16u * k, umma_smem_desc_sw128(b + 32u * k), idesc, (i > 0 || k > 0) ? 1u : 0u);   // X_lo . B_hi
                umma_f16_ts(td, xh + 16u * k, umma_smem_desc_sw128(b + 32u * k), idesc, 1u);                           // X_hi . B_hi
            }
            umma_commit(&c->empty[m.slot]);
            if (++m.slot == S16_NSLOTS) { m.slot = 0; ++m.use; }
            mbar_wait(&c->full[m.slot], m.use & 1u);                           // B_lo(ch)
            tc_fence_after();
            b = smem_u32(ring + m.slot * S16_SLOT_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ts(td, xh + 16u * k, umma_smem_desc_sw128(b + 32u * k), idesc, 1u);   // X_hi . B_lo
            umma_commit(&c->empty[m.slot]);
            if (++m.slot == S16_NSLOTS) { m.slot = 0; ++m.use; }
        }
        umma_commit(&c->done);
    }
}

// ---- compute warps --------------------------------------------------------------------------------------------------------------
// 16 consecutive per-column parameters as 4 LDG.128
__device__ __forceinline__ void s16_ldg16(const float* __restrict__ p, float (&v)[16]) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float4 t = __ldg(q + j); v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w; }
}
__device__ __forceinline__ void s16_tmem_ld16(uint32_t taddr, float (&v)[16]) {
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
// SDF of the row of this thread (TMEM lane 32 q + lane) at the normalised point (x, y, z): all 16 compute warps call it; warp
// (q, u) works on columns [64 j + 16 u, 64 j + 16 u + 16) of every K-chunk j, so chunk j of the next layer is complete after a
// quarter of the epilogue.  In place: the 16 accumulator columns just read become 8 hi + 8 lo operand columns.
// Returns this thread's partial of w6 . h5 over its 64 columns (caller adds the four u-parts and b6).
// inv5: 1 / scale of layers 1..5 (shared or global memory).
__device__ __forceinline__ float s16_compute_sdf(const SdfF16& sd, float x, float y, float z, S16Ctl* c, uint32_t& done_par, uint32_t tbase,
                                                 const float* inv5, PhaseClk* pc = nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, u = (warp >> 2) & 3;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    auto put = [&](int reg, int j, const float (&v)[16]) {
        uint32_t p[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
            p[i] = *reinterpret_cast<const uint32_t*>(&h);
            p[8 + i] = *reinterpret_cast<const uint32_t*>(&l);
        }
        tmem_st16(trow + 256u * reg + 64u * j + 16u * u, p);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&c->ready[j]);
    };
    {   // layer 0 (K = 3) on the FP32 pipe
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int col0 = 64 * j + 16 * u;
            float v[16], w0[16], w1[16], w2[16], pf[16], pb[16], pp[16];
            s16_ldg16(sd.Wt0 + col0, w0); s16_ldg16(sd.Wt0 + 256 + col0, w1); s16_ldg16(sd.Wt0 + 512 + col0, w2);
            s16_ldg16(sd.freq + col0, pf); s16_ldg16(sd.b[0] + col0, pb); s16_ldg16(sd.phase + col0, pp);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float a = fmaf(w2[i], z, fmaf(w1[i], y, w0[i] * x));
                v[i] = sin_cw(30.0f * (pf[i] * (a + pb[i]) + pp[i]));
            }
            put(0, j, v);
        }
    }
    if (pc) pc->mark(1);
    float dot = 0.f;
#pragma unroll 1
    for (int L = 1; L <= 5; ++L) {
        mbar_wait(&c->done, done_par);
        done_par ^= 1u;
        __syncwarp();
        tc_fence_after();
        if (pc) pc->mark(2);
        const int dreg = L & 1;
        const float inv = inv5[L - 1];
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int col0 = 64 * j + 16 * u;
            float v[16], pf[16], pb[16], pp[16];
            s16_ldg16(sd.freq + L * 256 + col0, pf); s16_ldg16(sd.b[L] + col0, pb); s16_ldg16(sd.phase + L * 256 + col0, pp);
            s16_tmem_ld16(trow + 256u * dreg + (uint32_t)col0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = sin_cw(30.0f * (pf[i] * (v[i] * inv + pb[i]) + pp[i]));
            if (L < 5) put(dreg, j, v);
            else {
                s16_ldg16(sd.w6 + col0, pf);
#pragma unroll
                for (int i = 0; i < 16; ++i) dot = fmaf(v[i], pf[i], dot);
            }
        }
        if (L == 5) tc_fence_before();
        if (pc) pc->mark(3);
    }
    return dot;
}

}  // namespace arah
