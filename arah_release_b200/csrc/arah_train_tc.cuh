// arah_train_tc.cuh — tcgen05 (5th-gen tensor core) GEMM of the training engine, sm_100a only.
//
// Same contract as k_gemm in arah_train_cuda.cuh:  C[i][j] (+)= bias[j] + sum_k A[i sa_i + k sa_k] B[k sb_k + j sb_j]
// with arbitrary strides, so one kernel serves forward (A k-contiguous, B k-contiguous = W[out][in]), backward-data
// (B j-contiguous) and weight-gradient (A i-contiguous, B j-contiguous, K = number of points, split-K + atomics) calls.
//
// One CTA owns a 128 x BN output tile (BN in {32, 64, 128, 256} >= N-extent, zero padded):
//   * both operands are staged by all 8 warps from global memory (coalesced along whichever index is contiguous) into the
//     K-major SWIZZLE_128B canonical layout of arah_umma.cuh (rows of 32 fp32 = one 128-byte swizzle row, 16-byte units
//     XOR-ed with row % 8; conflict-free 16-byte stores), rounded to TF32 with round-to-nearest on the way
//     (cvt.rna.tf32.f32; the tensor core would truncate),
//   * two stages: while the elected thread's four tcgen05.mma (M128 x BN x K8, kind::tf32, operands from shared memory,
//     fp32 accumulators in TMEM) of K-chunk c run, all warps stage chunk c+1; tcgen05.commit -> mbarrier frees a stage,
//   * epilogue: tcgen05.ld 32 columns per warp pass, bias, then store / read-modify-write / atomicAdd (split-K).
// Precision: fp32 accumulation in TMEM; operands either 3xTF32 (X3: hi/lo split, three MMAs per K-step, fp32-class — the
// default) or single-pass TF32.  BASELINE configs[2] asks for bf16-class MLP arithmetic with gradient cosine >= 0.999 against
// the fp32 reference: single-pass TF32 (3 more mantissa bits than bf16) measured only 0.956 on the first SIREN layer's weight
// gradient (the x30 sine arguments amplify operand rounding), 3xTF32 matches fp32 (tests/test_gpu_train.py).
#pragma once
#include "arah_umma.cuh"

namespace arah {
namespace train {

constexpr int TC_THREADS_GEMM = 256;
__host__ __device__ constexpr int gemm_tc_smem_bytes(int BN, bool x3) { return 2 * (UM * UK + BN * UK) * 4 * (x3 ? 2 : 1) + 1024 + 64; }

// ---- operand staging, split in two halves so that the global loads of K-chunk c+1 are in flight (in registers) while the
// tensor core works on chunk c:  fetch_operand (global -> registers)  ...  commit_operand (registers -> TF32 -> swizzled smem)
// Tile = ROWS rows x 32 k-values starting at (row0, k0); zero outside [.., nrows) x [.., kend).
// K_CONTIG: 8 consecutive lanes cover the 32 k-values of one row (one 128-byte line; one LDG.128 each when VEC);
// otherwise consecutive lanes take consecutive rows for a fixed 16-byte unit (coalesced along the row index).
template <int ROWS, bool K_CONTIG, bool VEC>
__device__ __forceinline__ void fetch_operand(float4 (&reg)[ROWS * 8 / TC_THREADS_GEMM], const float* __restrict__ src, long s_row, long s_k,
                                              int row0, int nrows, int k0, int kend, int tid) {
#pragma unroll
    for (int it = 0; it < ROWS * 8 / TC_THREADS_GEMM; ++it) {
        const int u = it * TC_THREADS_GEMM + tid;
        const int r = K_CONTIG ? (u >> 3) : (u % ROWS);
        const int j = K_CONTIG ? (u & 7) : (u / ROWS);
        const int gr = row0 + r, gk = k0 + 4 * j;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < nrows && gk < kend) {
            const float* p = src + (long)gr * s_row + (long)gk * s_k;
            if (VEC && gk + 3 < kend) v = __ldg(reinterpret_cast<const float4*>(p));
            else {
                v.x = __ldg(p);
                if (gk + 1 < kend) v.y = __ldg(p + s_k);
                if (gk + 2 < kend) v.z = __ldg(p + 2 * s_k);
                if (gk + 3 < kend) v.w = __ldg(p + 3 * s_k);
            }
        }
        reg[it] = v;
    }
}
// X3: split precision — dst holds hi = RN_tf32(x), dst_lo holds RN_tf32(x - hi) (3xTF32, see arah_umma.cuh)
template <int ROWS, bool K_CONTIG, bool X3>
__device__ __forceinline__ void commit_operand(float* __restrict__ dst, float* __restrict__ dst_lo, const float4 (&reg)[ROWS * 8 / TC_THREADS_GEMM], int tid) {
#pragma unroll
    for (int it = 0; it < ROWS * 8 / TC_THREADS_GEMM; ++it) {
        const int u = it * TC_THREADS_GEMM + tid;
        const int r = K_CONTIG ? (u >> 3) : (u % ROWS);
        const int j = K_CONTIG ? (u & 7) : (u / ROWS);
        const float4 v = reg[it];
        const float4 hi = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
        const int off = a_unit_off(r, 0, j);
        *reinterpret_cast<float4*>(dst + off) = hi;
        if (X3) *reinterpret_cast<float4*>(dst_lo + off) = make_float4(tf32_rn(v.x - hi.x), tf32_rn(v.y - hi.y), tf32_rn(v.z - hi.z), tf32_rn(v.w - hi.w));
    }
}

// VA / VB: the k-contiguous operand may be read with 16-byte loads (base and row stride 16-byte aligned);
// VC: C rows may be written with 16-byte stores (ldc % 4 == 0, base aligned)
template <int BN, bool A_KC, bool B_KC, bool X3, bool VA, bool VB>
__global__ void __launch_bounds__(TC_THREADS_GEMM) k_gemm_tc(int M, int N, int K, const float* __restrict__ A, long sa_i, long sa_k,
                                                             const float* __restrict__ B, long sb_k, long sb_j, float* __restrict__ C, int ldc,
                                                             const float* __restrict__ bias, int accumulate, int kchunk, int use_atomic, int vec_c) {
    extern __shared__ uint8_t raw_smem[];
    const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;
    float* sm = reinterpret_cast<float*>(raw_smem + (base - smem_u32(raw_smem)));
    constexpr int STAGE = (UM + BN) * UK * (X3 ? 2 : 1);      // floats per stage: A_hi [A_lo] B_hi [B_lo]
    float* As[2] = {sm, sm + STAGE};
    float* Al[2] = {sm + UM * UK, sm + STAGE + UM * UK};                        // only X3
    float* Bs[2] = {sm + (X3 ? 2 : 1) * UM * UK, sm + STAGE + (X3 ? 2 : 1) * UM * UK};
    float* Bl[2] = {Bs[0] + BN * UK, Bs[1] + BN * UK};                          // only X3
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + 2 * STAGE);                // free[2], done
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i0 = blockIdx.y * UM, j0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
    const int nchunks = (kend - kbeg + UK - 1) / UK;
    float4 ra[UM * 8 / TC_THREADS_GEMM], rb[BN * 8 / TC_THREADS_GEMM];
    if (nchunks > 0) {       // first chunk's loads overlap the barrier / TMEM set-up below
        fetch_operand<UM, A_KC, VA>(ra, A, sa_i, sa_k, i0, M, kbeg, kend, tid);
        fetch_operand<BN, B_KC, VB>(rb, B, sb_j, sb_k, j0, N, kbeg, kend, tid);
    }
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tslot, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    const uint32_t idesc = umma_idesc_tf32(UM, BN);
    for (int c = 0; c < nchunks; ++c) {
        const int s = c & 1;
        if (c >= 2) mbar_wait(&bars[s], ((c >> 1) - 1) & 1);        // the MMAs that read this stage two chunks ago are done
        commit_operand<UM, A_KC, X3>(As[s], Al[s], ra, tid);
        commit_operand<BN, B_KC, X3>(Bs[s], Bl[s], rb, tid);
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(As[s]), b_addr = smem_u32(Bs[s]);
            const uint32_t al_addr = smem_u32(Al[s]), bl_addr = smem_u32(Bl[s]);
#pragma unroll
            for (int k = 0; k < UK / UMMA_K_TF32; ++k) {
                const uint32_t ko = k * UMMA_K_TF32 * 4;
                if (X3) {       // (a_hi + a_lo)(b_hi + b_lo) without lo.lo, small terms first
                    umma_tf32(tbase, umma_smem_desc_sw128(al_addr + ko), umma_smem_desc_sw128(b_addr + ko), idesc, (c > 0 || k > 0) ? 1u : 0u);
                    umma_tf32(tbase, umma_smem_desc_sw128(a_addr + ko), umma_smem_desc_sw128(bl_addr + ko), idesc, 1u);
                    umma_tf32(tbase, umma_smem_desc_sw128(a_addr + ko), umma_smem_desc_sw128(b_addr + ko), idesc, 1u);
                } else umma_tf32(tbase, umma_smem_desc_sw128(a_addr + ko), umma_smem_desc_sw128(b_addr + ko), idesc, (c > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&bars[s]);
            if (c == nchunks - 1) umma_commit(&bars[2]);
        }
        if (c + 1 < nchunks) {   // next chunk: global -> registers while the tensor core runs
            const int k0 = kbeg + (c + 1) * UK;
            fetch_operand<UM, A_KC, VA>(ra, A, sa_i, sa_k, i0, M, k0, kend, tid);
            fetch_operand<BN, B_KC, VB>(rb, B, sb_j, sb_k, j0, N, k0, kend, tid);
        }
    }
    if (nchunks > 0) {
        mbar_wait(&bars[2], 0);
        tc_fence_after();
        const int q = warp & 3;
        const int gi = i0 + 32 * q + lane;
        for (int b = warp >> 2; b < BN / 32; b += 2) {
            float v[32];
            tmem_ld32(tbase + ((uint32_t)(32 * q) << 16) + (uint32_t)(32 * b), v);
            if (gi < M) {
                float* dst = C + (long)gi * ldc;
                const int gj0 = j0 + 32 * b;
                const bool zfirst = blockIdx.z == 0;
                if (vec_c && !use_atomic && gj0 + 32 <= N) {       // this thread owns 128 contiguous bytes of its row
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        float4 x = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                        if (bias && zfirst) { const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + gj0 + e)); x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w; }
                        float4* d4 = reinterpret_cast<float4*>(dst + gj0 + e);
                        if (accumulate) { const float4 o = *d4; x.x += o.x; x.y += o.y; x.z += o.z; x.w += o.w; }
                        *d4 = x;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int gj = gj0 + e;
                        if (gj < N) {
                            float x = v[e];
                            if (bias && zfirst) x += bias[gj];
                            if (use_atomic) atomicAdd(dst + gj, x);
                            else dst[gj] = accumulate ? (dst[gj] + x) : x;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, BN);
}

}  // namespace train
}  // namespace arah
