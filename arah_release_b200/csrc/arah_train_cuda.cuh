// arah_train_cuda.cuh — CUDA backend of the training engine (arah_train.h): strided SIMT fp32 GEMM with split-K,
// generic element-wise and column-reduction launchers.  fp32 FFMA on purpose: the training path is checked against the
// reference's fp32 autograd gradients to ~1e-5 relative; a tcgen05 3xTF32 GEMM can replace k_gemm behind the same call.
#pragma once
#include <cuda_runtime.h>
#include "arah_train.h"
#include "arah_train_tc.cuh"

namespace arah {
namespace train {

template <class F>
__global__ void __launch_bounds__(256) k_for_each(size_t n, F f) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) f(i);
}

// out[r][j] += sum_m f(m, j)[r];  f is called exactly once per (m, j) (it may also write per-element results)
template <int NR, class F>
__global__ void __launch_bounds__(256) k_col_reduce(int M, int N, int bx, int rows, F f, float* o0, float* o1, float* o2) {
    __shared__ float sh[NR][256];
    const int by = 256 / bx;
    const int tx = threadIdx.x % bx, ty = threadIdx.x / bx;
    const int m0 = blockIdx.x * rows, m1 = min(M, m0 + rows);
    float* outs[3] = {o0, o1, o2};
    for (int jb = 0; jb < N; jb += bx) {          // every thread runs the same number of rounds (barriers inside)
        const int j = jb + tx;
        const bool active = j < N;
        float acc[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) acc[r] = 0.0f;
        if (active)
            for (int m = m0 + ty; m < m1; m += by) {
                float red[NR];
                f(m, j, red);
#pragma unroll
                for (int r = 0; r < NR; ++r) acc[r] += red[r];
            }
        if (by > 1) {
            __syncthreads();
#pragma unroll
            for (int r = 0; r < NR; ++r) sh[r][threadIdx.x] = acc[r];
            __syncthreads();
            if (ty == 0) {
#pragma unroll
                for (int r = 0; r < NR; ++r) { float s = 0.0f; for (int y = 0; y < by; ++y) s += sh[r][y * bx + tx]; acc[r] = s; }
            }
        }
        if (ty == 0 && active) {
#pragma unroll
            for (int r = 0; r < NR; ++r) if (outs[r]) atomicAdd(outs[r] + j, acc[r]);
        }
    }
}

// C[i][j] (+)= bias[j] + sum_k A[i sa_i + k sa_k] B[k sb_k + j sb_j];  128 x BN x 16 tiles, 8 x (BN/16) per thread.
// A_KC: A is k-contiguous (sa_k == 1) else i-contiguous; B_JC: B is j-contiguous (sb_j == 1) else k-contiguous — only the
// thread -> element mapping of the tile loads depends on it (coalescing); any strides are legal.
template <int BN, bool A_KC, bool B_JC>
__global__ void __launch_bounds__(256) k_gemm(int M, int N, int K, const float* __restrict__ A, long sa_i, long sa_k,
                                              const float* __restrict__ B, long sb_k, long sb_j, float* __restrict__ C, int ldc,
                                              const float* __restrict__ bias, int accumulate, int kchunk, int use_atomic) {
    constexpr int BM = 128, BKK = 16, TM = 8, TN = BN / 16;
    __shared__ __align__(16) float As[BKK][BM + 4];
    __shared__ __align__(16) float Bs[BKK][BN + 4];
    const int t = threadIdx.x, tx = t % 16, ty = t / 16;
    const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
    float acc[TM][TN];
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) acc[r][c] = 0.0f;
    for (int k0 = kbeg; k0 < kend; k0 += BKK) {
#pragma unroll
        for (int e = 0; e < BM * BKK / 256; ++e) {
            const int idx = e * 256 + t;
            const int i = A_KC ? idx / BKK : idx % BM, k = A_KC ? idx % BKK : idx / BM;
            const int gi = i0 + i, gk = k0 + k;
            As[k][i] = (gi < M && gk < kend) ? A[(long)gi * sa_i + (long)gk * sa_k] : 0.0f;
        }
#pragma unroll
        for (int e = 0; e < BKK * BN / 256; ++e) {
            const int idx = e * 256 + t;
            const int j = B_JC ? idx % BN : idx / BKK, k = B_JC ? idx / BN : idx % BKK;
            const int gj = j0 + j, gk = k0 + k;
            Bs[k][j] = (gj < N && gk < kend) ? B[(long)gk * sb_k + (long)gj * sb_j] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BKK; ++k) {
            float a[TM], b[TN];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int c = 0; c < TN; ++c) b[c] = Bs[k][tx * TN + c];
#pragma unroll
            for (int r = 0; r < TM; ++r)
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < TM; ++r) {
        const int gi = i0 + ty * TM + r;
        if (gi >= M) continue;
#pragma unroll
        for (int c = 0; c < TN; ++c) {
            const int gj = j0 + tx * TN + c;
            if (gj >= N) continue;
            float v = acc[r][c];
            if (bias && blockIdx.z == 0) v += bias[gj];
            float* dst = C + (long)gi * ldc + gj;
            if (use_atomic) atomicAdd(dst, v);
            else *dst = accumulate ? (*dst + v) : v;
        }
    }
}

struct CudaBK {
    typedef cudaStream_t Stream;
    static float* alloc(size_t floats) {
        void* p = nullptr;
        if (cudaMalloc(&p, (floats ? floats : 1) * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return static_cast<float*>(p);
    }
    static void free(float* p) { if (p) cudaFree(p); }
    static void zero(float* p, size_t n, Stream st) { if (n) cudaMemsetAsync(p, 0, n * sizeof(float), st); }
    template <class F>
    static void for_each(size_t n, F f, Stream st) {
        if (n == 0) return;
        k_for_each<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, f);
        ++launches();
    }
    template <int NR, class F>
    static void col_reduce(int M, int N, F f, float* const* outs, Stream st) {
        static_assert(NR >= 1 && NR <= 3, "up to three sums per pass");
        if (M == 0 || N == 0) return;
        int bx = 32;
        while (bx < N && bx < 256) bx <<= 1;
        // 32 rows per row-lane: enough blocks (M/32 for 256-wide matrices) to keep the memory system busy; the price is one atomic
        // per column and block
        const int rows = 32 * (256 / bx);
        k_col_reduce<NR, F><<<(unsigned)((M + rows - 1) / rows), 256, 0, st>>>(M, N, bx, rows, f, outs[0], NR > 1 ? outs[1] : nullptr, NR > 2 ? outs[2] : nullptr);
        ++launches();
    }
    // 0 = tcgen05 3xTF32 (default, fp32-class), 1 = fp32 SIMT FFMA, 2 = tcgen05 single-pass TF32
    static int& precision() { static int p = 0; return p; }
    static void gemm_tc(int M, int N, int K, const float* A, long sa_i, long sa_k, const float* B, long sb_k, long sb_j, float* C, int ldc,
                        const float* bias, bool accumulate, Stream st) {
        const bool akc = (sa_k == 1), bkc = (sb_k == 1);
        const int BN = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
        dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + 127) / 128), 1);
        int splits = 1;
        const long tiles = (long)grid.x * grid.y;
        if (accumulate && K >= 2048 && tiles < 296) {
            splits = (int)((296 + tiles - 1) / tiles);
            const int maxs = (K + 511) / 512;
            if (splits > maxs) splits = maxs;
            if (splits < 1) splits = 1;
        }
        int kchunk = (K + splits - 1) / splits;
        kchunk = (kchunk + 31) / 32 * 32;
        splits = (K + kchunk - 1) / kchunk;
        grid.z = (unsigned)splits;
        const int ua = splits > 1 ? 1 : 0, acc = accumulate ? 1 : 0;
        const bool x3 = precision() == 0;
        const int smem = gemm_tc_smem_bytes(BN, x3);
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
        const bool va = akc && (sa_i % 4 == 0) && al16(A), vb = bkc && (sb_j % 4 == 0) && al16(B);
        const int vc = (ldc % 4 == 0 && al16(C) && (!bias || al16(bias))) ? 1 : 0;
#define ARAH_TC_CASE(BN_, AK, BK_, X3_, VA_, VB_) do { \
            static bool attr_set = false; \
            if (!attr_set) { cudaFuncSetAttribute(k_gemm_tc<BN_, AK, BK_, X3_, VA_, VB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_tc_smem_bytes(BN_, X3_)); attr_set = true; } \
            k_gemm_tc<BN_, AK, BK_, X3_, VA_, VB_><<<grid, TC_THREADS_GEMM, smem, st>>>(M, N, K, A, sa_i, sa_k, B, sb_k, sb_j, C, ldc, bias, acc, kchunk, ua, vc); } while (0)
        // vector loads only exist for k-contiguous operands: (AK, VA) in {(1,1), (1,0), (0,0)}
#define ARAH_TC_B(BN_, X3_, AK, VA_) do { \
            if (bkc && vb) ARAH_TC_CASE(BN_, AK, true, X3_, VA_, true); else if (bkc) ARAH_TC_CASE(BN_, AK, true, X3_, VA_, false); \
            else ARAH_TC_CASE(BN_, AK, false, X3_, VA_, false); } while (0)
#define ARAH_TC_AB(BN_, X3_) do { \
            if (akc && va) ARAH_TC_B(BN_, X3_, true, true); else if (akc) ARAH_TC_B(BN_, X3_, true, false); else ARAH_TC_B(BN_, X3_, false, false); } while (0)
#define ARAH_TC_BN(BN_) do { if (x3) ARAH_TC_AB(BN_, true); else ARAH_TC_AB(BN_, false); } while (0)
        if (BN == 32) ARAH_TC_BN(32); else if (BN == 64) ARAH_TC_BN(64); else if (BN == 128) ARAH_TC_BN(128); else ARAH_TC_BN(256);
#undef ARAH_TC_BN
#undef ARAH_TC_AB
#undef ARAH_TC_B
#undef ARAH_TC_CASE
        ++launches();
    }
    static void gemm(int M, int N, int K, const float* A, long sa_i, long sa_k, const float* B, long sb_k, long sb_j, float* C, int ldc,
                     const float* bias, bool accumulate, Stream st) {
        if (M == 0 || N == 0) return;
        if (precision() != 1) {
            if (N > 256) {      // wider than one tile: column blocks of 256
                for (int j = 0; j < N; j += 256)
                    gemm_tc(M, (N - j) < 256 ? (N - j) : 256, K, A, sa_i, sa_k, B + (long)j * sb_j, sb_k, sb_j, C + j, ldc, bias ? bias + j : nullptr, accumulate, st);
            } else gemm_tc(M, N, K, A, sa_i, sa_k, B, sb_k, sb_j, C, ldc, bias, accumulate, st);
            return;
        }
        const bool akc = (sa_k == 1), bjc = (sb_j == 1);
        const int BN = (N <= 32) ? 32 : 128;
        dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + 127) / 128), 1);
        // split-K for weight-gradient shapes (few output tiles, K = number of points): partial sums meet in C with atomics
        int splits = 1;
        const long tiles = (long)grid.x * grid.y;
        if (accumulate && K >= 2048 && tiles < 296) {
            splits = (int)((592 + tiles - 1) / tiles);
            const int maxs = (K + 255) / 256;
            if (splits > maxs) splits = maxs;
            if (splits < 1) splits = 1;
        }
        int kchunk = (K + splits - 1) / splits;
        kchunk = (kchunk + 15) / 16 * 16;
        splits = (K + kchunk - 1) / kchunk;
        grid.z = (unsigned)splits;
        const int ua = splits > 1 ? 1 : 0, acc = accumulate ? 1 : 0;
#define ARAH_GEMM_CASE(BN_, AK, BJ) k_gemm<BN_, AK, BJ><<<grid, 256, 0, st>>>(M, N, K, A, sa_i, sa_k, B, sb_k, sb_j, C, ldc, bias, acc, kchunk, ua)
        if (BN == 32) {
            if (akc && bjc) ARAH_GEMM_CASE(32, true, true); else if (akc) ARAH_GEMM_CASE(32, true, false);
            else if (bjc) ARAH_GEMM_CASE(32, false, true); else ARAH_GEMM_CASE(32, false, false);
        } else {
            if (akc && bjc) ARAH_GEMM_CASE(128, true, true); else if (akc) ARAH_GEMM_CASE(128, true, false);
            else if (bjc) ARAH_GEMM_CASE(128, false, true); else ARAH_GEMM_CASE(128, false, false);
        }
#undef ARAH_GEMM_CASE
        ++launches();
    }
    static long& launches() { static thread_local long n = 0; return n; }
};

}  // namespace train
}  // namespace arah
