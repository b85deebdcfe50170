// arah_tc2.cuh — second-generation tensor-core tile engine (sm_100a):
//   * activations (MMA operand A) live in TMEM, not shared memory: the epilogue thread that owns row r writes its 32-column
//     batches with tcgen05.st and the next layer reads them with the `.ts` form of tcgen05.mma (A from TMEM, B from smem);
//   * all of shared memory becomes a deep weight ring (6 x 32 KB) fed by a dedicated TMA producer warp that runs ahead
//     through the tile's static chunk schedule (across layer boundaries), so L2 latency is hidden and the MMA issuer only
//     ever waits for bandwidth;
//   * compute warps synchronise on named barrier 1 (256 threads); the producer warp never joins.
#pragma once
#include "arah_umma.cuh"

namespace arah {

constexpr int TC_NSLOTS = 6;
constexpr int TC_THREADS = 288;             // 8 compute warps + 1 producer warp

__device__ __forceinline__ void cta_sync_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// write 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// round to TF32 and store one 32-column batch of A (row = this thread's lane)
__device__ __forceinline__ void a_tmem_store(uint32_t taddr, float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = tf32_rn(v[i]);
    tmem_st32(taddr, v);
}
// hi/lo split store (3xTF32): hi -> taddr_hi, lo -> taddr_lo
__device__ __forceinline__ void a_tmem_store_split(uint32_t taddr_hi, uint32_t taddr_lo, const float (&v)[32]) {
    float h[32], l[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { h[i] = tf32_rn(v[i]); l[i] = tf32_rn(v[i] - h[i]); }
    tmem_st32(taddr_hi, h);
    tmem_st32(taddr_lo, l);
}

// sin_cw / sincos_cw (Cody-Waite sine / cosine used by the tensor-core epilogues) live in arah_math.cuh so that the host tests can
// compile them (tests/test_host_math.py).

struct RingPos {
    uint32_t slot, use;
    __device__ __forceinline__ void next() { if (++slot == TC_NSLOTS) { slot = 0; ++use; } }
};
struct TCRing {
    float* buf;          // [TC_NSLOTS][RING_SLOT_FLOATS]
    uint64_t* full;      // [TC_NSLOTS]
    uint64_t* empty;     // [TC_NSLOTS]
};
__device__ __forceinline__ void tcring_init(const TCRing& rg) {      // one thread
    for (int i = 0; i < TC_NSLOTS; ++i) { mbar_init(&rg.full[i], 1); mbar_init(&rg.empty[i], 1); }
}
// producer (one lane of the producer warp): stream `nchunks` chunks of `bytes` each starting at src
__device__ __forceinline__ void tcring_produce(const TCRing& rg, RingPos& p, const float* __restrict__ src, int nchunks, uint32_t bytes) {
    for (int c = 0; c < nchunks; ++c) {
        if (p.use > 0) mbar_wait(&rg.empty[p.slot], (p.use - 1) & 1u);
        mbar_expect_tx(&rg.full[p.slot], bytes);
        bulk_g2s(rg.buf + p.slot * RING_SLOT_FLOATS, reinterpret_cast<const char*>(src) + (size_t)c * bytes, bytes, &rg.full[p.slot]);
        p.next();
    }
}
// consumer (the MMA-issuing thread): D (+)= A_tmem[:, 0..32*nchunks) . B^T, single precision pass
__device__ __forceinline__ void tcring_mma_layer(const TCRing& rg, RingPos& p, uint32_t tmem_a, int nchunks, int N, uint32_t tmem_d,
                                                 uint32_t first_accumulate, uint64_t* done_bar) {
    const uint32_t idesc = umma_idesc_tf32(UM, N);
    for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&rg.full[p.slot], p.use & 1u);
        tc_fence_after();
        const uint32_t b_addr = smem_u32(rg.buf + p.slot * RING_SLOT_FLOATS);
#pragma unroll
        for (int k = 0; k < UK / UMMA_K_TF32; ++k)
            umma_tf32_ts(tmem_d, tmem_a + (uint32_t)(c * UK + k * UMMA_K_TF32), umma_smem_desc_sw128(b_addr + k * UMMA_K_TF32 * 4), idesc,
                         (c > 0 || k > 0) ? 1u : first_accumulate);
        umma_commit(&rg.empty[p.slot]);
        p.next();
    }
    if (done_bar) umma_commit(done_bar);
}
// 3xTF32 consumer: chunk image = [B_hi | B_lo]; D = A_lo.B_hi + A_hi.B_lo + A_hi.B_hi
__device__ __forceinline__ void tcring_mma_layer_x3(const TCRing& rg, RingPos& p, uint32_t tmem_a_hi, uint32_t tmem_a_lo, int nchunks, int N,
                                                    uint32_t tmem_d, uint64_t* done_bar) {
    const uint32_t idesc = umma_idesc_tf32(UM, N);
    for (int c = 0; c < nchunks; ++c) {
        mbar_wait(&rg.full[p.slot], p.use & 1u);
        tc_fence_after();
        const uint32_t bh = smem_u32(rg.buf + p.slot * RING_SLOT_FLOATS), bl = bh + (uint32_t)N * UK * 4;
#pragma unroll
        for (int k = 0; k < UK / UMMA_K_TF32; ++k) {
            const uint32_t col = (uint32_t)(c * UK + k * UMMA_K_TF32), ko = k * UMMA_K_TF32 * 4;
            umma_tf32_ts(tmem_d, tmem_a_lo + col, umma_smem_desc_sw128(bh + ko), idesc, (c > 0 || k > 0) ? 1u : 0u);
            umma_tf32_ts(tmem_d, tmem_a_hi + col, umma_smem_desc_sw128(bl + ko), idesc, 1u);
            umma_tf32_ts(tmem_d, tmem_a_hi + col, umma_smem_desc_sw128(bh + ko), idesc, 1u);
        }
        umma_commit(&rg.empty[p.slot]);
        p.next();
    }
    if (done_bar) umma_commit(done_bar);
}

}  // namespace arah
