// arah_sdf16.cuh — the 256-wide FiLM-SIREN SDF (hyperlayers.py:385-415, siren_modules.py:31-37) on tcgen05 kind::f16 in split
// precision (arah_f16x3.cuh: three fp16 products ~ one fp32 product), the engine of the persistent root-finding kernels
// (sphere tracing, joint search) and of the canonical SDF lattice.
//
// Everything lives in tensor memory now.  TMEM = two 256-column regions R0 / R1.  The activations of layer L sit in one region
// as four K-chunks of 64 values: chunk j -> columns [64 j, 64 j + 32) hi | [64 j + 32, 64 j + 64) lo (two K values per column);
// the accumulators of layer L go to the other region.  The epilogue turns D into the next layer's operand IN PLACE, chunk by
// chunk (a thread reads the 64 accumulator columns of a chunk, then overwrites them with 32 hi + 32 lo columns) and arrives on
// the chunk's `ready` barrier; the MMA warp starts chunk j of the next layer as soon as ready[j] and the weight images are
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
constexpr int S16_THREADS = 320;         // 8 compute warps, producer warp, MMA warp

struct S16Ctl {                          // shared-memory control block
    uint64_t full[S16_NSLOTS], empty[S16_NSLOTS];
    uint64_t ready[4];                   // SDF operand chunk j written (4 warp arrivals)
    uint64_t done;                       // a layer's accumulators are complete
    uint64_t go;                         // compute -> MMA warp: decision about the next evaluation is in cont[]
    uint64_t ready_sk;                   // skinning operand written by all 8 warps (joint search only)
    uint32_t tslot;
    volatile int stop;                   // compute -> producer: no further evaluation
    volatile int cont[2];
};

__device__ __forceinline__ void s16_ctl_init(S16Ctl* c) {      // one thread
    for (int i = 0; i < S16_NSLOTS; ++i) { mbar_init(&c->full[i], 1); mbar_init(&c->empty[i], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&c->ready[i], 4);
    mbar_init(&c->done, 1); mbar_init(&c->go, 1); mbar_init(&c->ready_sk, 8);
    mbar_fence_init();
    c->stop = 0; c->cont[0] = 0; c->cont[1] = 0;
}
__device__ __forceinline__ int s16_chunk(int i) { return (i >> 1) + ((i & 1) << 1); }      // 0, 2, 1, 3: the order the two column halves finish
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}

__device__ __forceinline__ bool cta_or_compute(bool pred) {                      // barrier 1 over the 256 compute threads + OR
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, 1, 256, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"((uint32_t)pred) : "memory");
    return r != 0;
}

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
            const uint32_t xh = ta + 64u * ch, xl = xh + 32u;
            mbar_wait(&c->full[m.slot], m.use & 1u);                           // B_hi(ch)
            tc_fence_after();
            uint32_t b = smem_u32(ring + m.slot * S16_SLOT_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                umma_f16_ts(td, xl + 8u * k, umma_smem_desc_sw128(b + 32u * k), idesc, (i > 0 || k > 0) ? 1u : 0u);    // X_lo . B_hi
                umma_f16_ts(td, xh + 8u * k, umma_smem_desc_sw128(b + 32u * k), idesc, 1u);                            // X_hi . B_hi
            }
            umma_commit(&c->empty[m.slot]);
            if (++m.slot == S16_NSLOTS) { m.slot = 0; ++m.use; }
            mbar_wait(&c->full[m.slot], m.use & 1u);                           // B_lo(ch)
            tc_fence_after();
            b = smem_u32(ring + m.slot * S16_SLOT_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ts(td, xh + 8u * k, umma_smem_desc_sw128(b + 32u * k), idesc, 1u);    // X_hi . B_lo
            umma_commit(&c->empty[m.slot]);
            if (++m.slot == S16_NSLOTS) { m.slot = 0; ++m.use; }
        }
        umma_commit(&c->done);
    }
}

// ---- compute warps --------------------------------------------------------------------------------------------------------------
// 32 consecutive per-column parameters as 8 LDG.128
__device__ __forceinline__ void s16_ldg32(const float* __restrict__ p, float (&v)[32]) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float4 t = __ldg(q + j); v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w; }
}
// SDF of the row of this thread (TMEM lane 32 q + lane) at the normalised point (x, y, z): all 8 compute warps call it; a thread
// works on the 128 columns of its half h.  Returns this thread's partial of w6 . h5 (caller adds the two halves and b6).
// inv5: 1 / scale of layers 1..5 (shared or global memory).
__device__ __forceinline__ float s16_compute_sdf(const SdfF16& sd, float x, float y, float z, S16Ctl* c, uint32_t& done_par, uint32_t tbase,
                                                 const float* inv5, PhaseClk* pc = nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, h = (warp >> 2) & 1;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    // operand chunk j of region `reg` from the activations v0 (K = 64 j ..) and v1 (K = 64 j + 32 ..)
    auto put = [&](int reg, int j, const uint32_t (&hi0)[16], const uint32_t (&lo0)[16], const float (&v1)[32]) {
        uint32_t hi1[16], lo1[16];
        split_pack_f16(v1, hi1, lo1);
        const uint32_t a = trow + 256u * reg + 64u * j;
        tmem_st16(a, hi0); tmem_st16(a + 16u, hi1); tmem_st16(a + 32u, lo0); tmem_st16(a + 48u, lo1);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&c->ready[j]);
    };
    {   // layer 0 (K = 3) on the FP32 pipe
#pragma unroll 1
        for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * h + jj;
            uint32_t hi0[16], lo0[16];
            float v[32];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int col0 = 64 * j + 32 * hf;
                float w0[32], w1[32], w2[32], pf[32], pb[32], pp[32];
                s16_ldg32(sd.Wt0 + col0, w0); s16_ldg32(sd.Wt0 + 256 + col0, w1); s16_ldg32(sd.Wt0 + 512 + col0, w2);
                s16_ldg32(sd.freq + col0, pf); s16_ldg32(sd.b[0] + col0, pb); s16_ldg32(sd.phase + col0, pp);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float a = fmaf(w2[i], z, fmaf(w1[i], y, w0[i] * x));
                    v[i] = sin_cw(30.0f * (pf[i] * (a + pb[i]) + pp[i]));
                }
                if (hf == 0) split_pack_f16(v, hi0, lo0);
            }
            put(0, j, hi0, lo0, v);
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
        for (int jj = 0; jj < 2; ++jj) {
            const int j = 2 * h + jj;
            uint32_t hi0[16], lo0[16];
            float v[32];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int col0 = 64 * j + 32 * hf;
                float pf[32], pb[32], pp[32];
                s16_ldg32(sd.freq + L * 256 + col0, pf); s16_ldg32(sd.b[L] + col0, pb); s16_ldg32(sd.phase + L * 256 + col0, pp);
                tmem_ld32(trow + 256u * dreg + (uint32_t)col0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = sin_cw(30.0f * (pf[i] * (v[i] * inv + pb[i]) + pp[i]));
                if (L < 5) { if (hf == 0) split_pack_f16(v, hi0, lo0); }
                else {
                    s16_ldg32(sd.w6 + col0, pf);
#pragma unroll
                    for (int i = 0; i < 32; ++i) dot = fmaf(v[i], pf[i], dot);
                }
            }
            if (L < 5) put(dreg, j, hi0, lo0, v);          // in place: both halves of the chunk's accumulators have been read
        }
        if (L == 5) tc_fence_before();
        if (pc) pc->mark(3);
    }
    return dot;
}

}  // namespace arah
