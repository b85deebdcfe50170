// arah_umma.cuh — tcgen05 (5th-gen tensor core) building blocks for the shading tiles, sm_100a only.
//
// One CTA computes D[128 x N] (+)= A[128 x K] . B[N x K]^T with
//   A : activations, fp32 read as TF32, K-major, SWIZZLE_128B canonical layout in shared memory
//       (K-chunks of 32 floats = one 128-byte swizzle row; chunk c at A + c*16 KB; row r at (r/8)*1024 + (r%8)*128;
//        16-byte unit j of the row stored at unit j ^ (r%8)),
//   B : weights [N][K] (the reference's own [out][in] layout for forward layers), pre-swizzled per 32-wide K-chunk in
//       global memory by k_pack_umma so that one chunk (N*128 bytes) is a contiguous image of its shared-memory tile:
//       a single 1-D TMA bulk copy stages it — no tensor map needed,
//   D : fp32 accumulators in TMEM (lane = row, column = output feature), read back with tcgen05.ld (32x32b).
// The instruction descriptor / shared-memory descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp
// (UMMA::InstrDescriptor, UMMA::SmemDescriptor) from the vendored CUTLASS tree; the PTX is written here by hand.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "arah_tile.cuh"

namespace arah {

constexpr int UM = 128;                 // rows per UMMA tile (UMMA_M, cta_group::1)
constexpr int UK = 32;                  // floats per K-chunk (128-byte swizzle row)
constexpr int A_CHUNK_FLOATS = UM * UK; // 16 KB
constexpr int UMMA_K_TF32 = 8;          // K per tcgen05.mma kind::tf32

// ---- TMEM allocation (one warp) --------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes (st.shared of the A tile) visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- descriptors -----------------------------------------------------------------------------------------------
// UMMA::InstrDescriptor: c_format F32 (1) @4, a/b_format TF32 (2) @7/@10, a/b K-major (0) @15/@16, N>>3 @17, M>>4 @24
__device__ __forceinline__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// UMMA::SmemDescriptor, K-major SWIZZLE_128B: start>>4 @0, LBO(=1, unused for swizzled K-major) @16, SBO = 1024>>4 @32,
// version 1 @46, layout_type SWIZZLE_128B (2) @61
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane (warp w may only touch lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- A tile addressing -----------------------------------------------------------------------------------------
// float offset of (row r, 16-byte unit j of K-chunk c) inside the A buffer
__device__ __forceinline__ int a_unit_off(int r, int c, int j) {
    return c * A_CHUNK_FLOATS + (r >> 3) * 256 + (r & 7) * 32 + ((j ^ (r & 7)) << 2);
}
// fp32 -> TF32 with round-to-nearest (the tensor core would otherwise TRUNCATE the low 13 mantissa bits: 2x the
// error and a systematic bias towards zero that accumulates coherently over K)
// cvt.rna.tf32.f32 (round to nearest, ties away from zero) lowers to 4 SASS instructions on sm_100a (|x| >= inf test, add,
// mask, select); every epilogue runs it once or twice per activation.  For finite values — all we ever feed it: bounded
// activations and weights — the add-and-mask below yields the identical bit pattern in 2 instructions.
__device__ __forceinline__ float tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// store 32 consecutive K values (one chunk) of row r
__device__ __forceinline__ void a_store_chunk(float* A, int r, int c, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(A + a_unit_off(r, c, j)) =
            make_float4(tf32_rn(v[4 * j]), tf32_rn(v[4 * j + 1]), tf32_rn(v[4 * j + 2]), tf32_rn(v[4 * j + 3]));
}

// ---- weight ring + MMA issue (called by ONE thread) --------------------------------------------------------------
struct URing {
    float* buf;           // [2][N_max*32] floats (32 KB per slot for N = 256)
    uint64_t* full;       // [2] TMA landed
    uint64_t* empty;      // [2] MMAs that read the slot completed
    uint32_t fill_cnt;    // chunks issued so far (slot = cnt & 1, use index = cnt >> 1)
    uint32_t mma_cnt;     // chunks consumed so far
};
constexpr int RING_SLOT_FLOATS = 256 * UK;     // 8192 floats = 32 KB

// One layer: D[UM x N] (+)= A[:, 0 .. 32*nchunks) . B^T with B chunk c at Wsw + c*N*32 floats (pre-swizzled).
// `first_accumulate` = 0 overwrites D with the first MMA.  After the last MMA commits `done_bar` (if not null).
__device__ __forceinline__ void umma_layer_issue(URing& rg, const float* A_smem, const float* __restrict__ Wsw, int nchunks, int N,
                                                 uint32_t tmem_d, uint32_t first_accumulate, uint64_t* done_bar) {
    const uint32_t idesc = umma_idesc_tf32(UM, N);
    const uint32_t chunk_bytes = (uint32_t)N * UK * 4;
    auto load = [&](int c) {
        const uint32_t s = rg.fill_cnt & 1u, use = rg.fill_cnt >> 1;
        if (use > 0) mbar_wait(&rg.empty[s], (use - 1) & 1u);          // MMAs of the previous tenant are done
        mbar_expect_tx(&rg.full[s], chunk_bytes);
        bulk_g2s(rg.buf + s * RING_SLOT_FLOATS, Wsw + (size_t)c * N * UK, chunk_bytes, &rg.full[s]);
        rg.fill_cnt++;
    };
    load(0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) load(c + 1);
        const uint32_t s = rg.mma_cnt & 1u, use = rg.mma_cnt >> 1;
        mbar_wait(&rg.full[s], use & 1u);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(A_smem + c * A_CHUNK_FLOATS);
        const uint32_t b_addr = smem_u32(rg.buf + s * RING_SLOT_FLOATS);
#pragma unroll
        for (int k = 0; k < UK / UMMA_K_TF32; ++k) {
            const uint64_t ad = umma_smem_desc_sw128(a_addr + k * UMMA_K_TF32 * 4);
            const uint64_t bd = umma_smem_desc_sw128(b_addr + k * UMMA_K_TF32 * 4);
            umma_tf32(tmem_d, ad, bd, idesc, (c > 0 || k > 0) ? 1u : first_accumulate);
        }
        umma_commit(&rg.empty[s]);
        rg.mma_cnt++;
    }
    if (done_bar) umma_commit(done_bar);
}

}  // namespace arah

namespace arah {

// ---- split-precision (3xTF32) layer: D (+)= (A_hi + A_lo) . (B_hi + B_lo)^T without the lo.lo term ------------------
// Used where residuals must resolve 1e-5 m (root finding): hi = RN_tf32(x), lo = RN_tf32(x - hi); the three products
// carry ~21 mantissa bits of each operand (error ~2^-22 |a||b| per term, comparable to fp32 accumulation noise).
// Ring slot layout per K-chunk: [B_hi image (N*32 floats) | B_lo image (N*32 floats)], staged by ONE bulk copy.
__device__ __forceinline__ void umma_layer_issue_x3(URing& rg, const float* A_hi, const float* A_lo, const float* __restrict__ Wsw,
                                                    int nchunks, int N, uint32_t tmem_d, uint64_t* done_bar) {
    const uint32_t idesc = umma_idesc_tf32(UM, N);
    const uint32_t chunk_bytes = (uint32_t)N * UK * 4 * 2;
    auto load = [&](int c) {
        const uint32_t s = rg.fill_cnt & 1u, use = rg.fill_cnt >> 1;
        if (use > 0) mbar_wait(&rg.empty[s], (use - 1) & 1u);
        mbar_expect_tx(&rg.full[s], chunk_bytes);
        bulk_g2s(rg.buf + s * RING_SLOT_FLOATS, Wsw + (size_t)c * N * UK * 2, chunk_bytes, &rg.full[s]);
        rg.fill_cnt++;
    };
    load(0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) load(c + 1);
        const uint32_t s = rg.mma_cnt & 1u, use = rg.mma_cnt >> 1;
        mbar_wait(&rg.full[s], use & 1u);
        tc_fence_after();
        const uint32_t ah = smem_u32(A_hi + c * A_CHUNK_FLOATS), al = smem_u32(A_lo + c * A_CHUNK_FLOATS);
        const uint32_t bh = smem_u32(rg.buf + s * RING_SLOT_FLOATS), bl = bh + (uint32_t)N * UK * 4;
#pragma unroll
        for (int k = 0; k < UK / UMMA_K_TF32; ++k) {
            const uint32_t ko = k * UMMA_K_TF32 * 4;
            umma_tf32(tmem_d, umma_smem_desc_sw128(al + ko), umma_smem_desc_sw128(bh + ko), idesc, (c > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tmem_d, umma_smem_desc_sw128(ah + ko), umma_smem_desc_sw128(bl + ko), idesc, 1u);
            umma_tf32(tmem_d, umma_smem_desc_sw128(ah + ko), umma_smem_desc_sw128(bh + ko), idesc, 1u);
        }
        umma_commit(&rg.empty[s]);
        rg.mma_cnt++;
    }
    if (done_bar) umma_commit(done_bar);
}
// store one K-chunk of row r split into hi/lo tiles
__device__ __forceinline__ void a_store_chunk_split(float* A_hi, float* A_lo, int r, int c, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { h[e] = tf32_rn(v[4 * j + e]); l[e] = tf32_rn(v[4 * j + e] - h[e]); }
        const int off = a_unit_off(r, c, j);
        *reinterpret_cast<float4*>(A_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(A_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
}
// softplus(beta=100) with MUFU exp/log: max(x,0) + log(1 + exp(-|100 x|)) / 100; abs error < 1e-8 thanks to the 1/100
// log(1 + exp(-|100 x|)) / 100 + max(x, 0).  Same two MUFU results as __expf / __logf (identical roundings of the argument and
// of the result), but through the .ftz forms: the non-ftz intrinsics carry denormal fix-up code (3-4 extra instructions per
// call) that can never trigger here (exp's denormal results vanish in 1 + e, log's argument lies in [1, 2]).
// The two scale factors on either side are merged (100 log2 e; ln 2 / 100): 6 instructions instead of 8; against the two-step
// scaling the result moves by < 1e-8 absolute (the log term is <= 0.0069), two orders below the 3xTF32 product error.
__device__ __forceinline__ float softplus100_fast(float x) {
    float e, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-fabsf(x) * 144.26950408889634f));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
    return fmaf(l, 0.006931471805599453f, fmaxf(x, 0.0f));
}
// (Round 2 tried moving the logarithm of every second element to the FP32 pipe, e P(e) with a degree-6 polynomial, to relieve the
// MUFU unit in k_corr_persist: one MUFU less but four FP32 instructions more per element made the kernel 1.3 ms SLOWER — the
// epilogues are as much issue-bound as MUFU-bound.  Reverted.  Also tried: a warp-voted shortcut — for |x| >= 0.1675 the function
// returns max(x, 0) bit-exactly (1 + e rounds to 1), and two thirds of the (warp, hidden unit) pairs of the synthetic frame qualify —
// but a vote + uniform branch per element serialises the 32 independent MUFU chains the straight-line code overlaps: 31.9 -> 40.3 ms.)

}  // namespace arah
