// arah_sdf_fwd16.cuh — k_sdf_fwd16: SDF value of every converged sample (the pass in front of the exact alpha cull,
// renderer/implicit_differentiable_renderer.py:311-313,359 restated for the whole shade list) on tcgen05 kind::f16, single pass.
//
// The value that compositing turns into sigma / alpha needs operand precision of ~11 bits (round 1 used TF32 here); fp16 has the
// same significand, issues at twice the TF32 rate and halves the weight stream (the kernel round 1 ran for this pass moved
// 156 GB of weight images per frame from L2, 8.2 TB/s).  Sixteen epilogue warps instead of eight: a lane quarter of tensor
// memory is shared by four warps, each of which converts 16 accumulator columns of EVERY K-chunk (tcgen05.ld x16 -> FiLM sine
// -> 8 packed operand columns, in place), so chunk j of the next layer is complete after a quarter of the epilogue and the MMA
// warp runs chunk by chunk behind it.
// TMEM: regions R0 / R1 of 256 columns; layer l reads its operand from R[(l-1)&1] (K-step (j, k) = 8 columns at 64 j + 16 k),
// accumulates into R[l&1]; sin(fma(acc, F, G)) with F = 30 f / s_l, G = 30 (f b + phi) staged in shared memory.
// Weights: the hi images of arah_sdf16.cuh (pre-scaled by s_l), 32 KB per K-chunk, 5-slot ring, producer warp.
#pragma once
#include "arah_sdf16.cuh"

namespace arah {

constexpr int F16_THREADS = 576;            // 16 epilogue warps + producer + MMA issuer
constexpr int F16_NSLOTS = 5;

__host__ __device__ constexpr size_t sdf_fwd16_smem_bytes() {
    // ring | F, G [6][256] each | W0t [3][256] | w6 [256] | xs [128][4] | part [4][128] | barriers
    return (size_t)F16_NSLOTS * 32768 + (size_t)(2 * 6 * 256 + 3 * 256 + 256 + UM * 4 + 4 * UM) * 4 + 256;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// Lattice mode (arah_grid16.cuh, the coarse pass of the banded lattice): the points are the first `n` lattice points of an N^3
// lattice and the RAW network output goes to out[i].
struct Fwd16Grid { int N; float voxel; int n; float* out; };

// g.out == nullptr: write w.smp_sdf[slot] (metres) for the slots of w.shade_list[0 .. counters[C_SHADE]).
__global__ void __launch_bounds__(F16_THREADS, 1) k_sdf_fwd16(FrameParams fp, SdfF16 sd, Work w, Fwd16Grid g) {
    extern __shared__ __align__(1024) uint8_t raw_smem[];
    const int n = g.out ? g.n : w.counters[C_SHADE];
    const int ntiles = (n + UM - 1) / UM;
    if ((int)blockIdx.x >= ntiles) return;
    if (smem_u32(raw_smem) & 1023u) __trap();
    uint8_t* ring = raw_smem;
    float* sF = reinterpret_cast<float*>(ring + F16_NSLOTS * 32768);       // [6][256]
    float* sG = sF + 6 * 256;
    float* sW0 = sG + 6 * 256;                                              // [3][256]
    float* sW6 = sW0 + 3 * 256;
    float (*xs)[4] = reinterpret_cast<float (*)[4]>(sW6 + 256);
    float (*part)[UM] = reinterpret_cast<float (*)[UM]>(reinterpret_cast<float*>(xs) + UM * 4);   // [4][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(part) + 4 * UM);
    uint64_t* full = bars;                      // [5]
    uint64_t* empty = bars + F16_NSLOTS;        // [5]
    uint64_t* ready = bars + 2 * F16_NSLOTS;    // [4] operand chunk j written by all 16 warps
    uint64_t* done = ready + 4;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < F16_NSLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 4; ++i) mbar_init(&ready[i], 16);
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 17) tmem_alloc(tslot, 512);
    for (int i = tid; i < 6 * 256; i += F16_THREADS) {
        const int l = i >> 8;
        const float f = __ldg(sd.freq + i), inv = (l == 0) ? 1.0f : __ldg(sd.scale + 2 * (l - 1) + 1);
        sF[i] = 30.0f * f * inv;
        sG[i] = 30.0f * (f * __ldg(sd.b[l] + (i & 255)) + __ldg(sd.phase + i));
    }
    for (int i = tid; i < 3 * 256; i += F16_THREADS) sW0[i] = __ldg(sd.Wt0 + i);
    for (int i = tid; i < 256; i += F16_THREADS) sW6[i] = __ldg(sd.w6 + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;

    if (warp == 16) {                                   // ===== TMA producer =====
        if (lane == 0) {
            uint32_t slot = 0, use = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int L = 0; L < 5; ++L)
                    for (int j = 0; j < 4; ++j) {
                        if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1u);
                        mbar_expect_tx(&full[slot], 32768u);
                        bulk_g2s(ring + slot * 32768, reinterpret_cast<const char*>(sd.hi) + (size_t)L * 131072 + (size_t)j * 32768, 32768u, &full[slot]);
                        if (++slot == F16_NSLOTS) { slot = 0; ++use; }
                    }
        }
        return;
    }
    if (warp == 17) {                                   // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t slot = 0, use = 0, rpar = 0;
            const uint32_t idesc = umma_idesc_f16(UM, 256);
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
                for (int L = 1; L <= 5; ++L) {
                    const uint32_t ta = tbase + 256u * ((L - 1) & 1), td = tbase + 256u * (L & 1);
                    for (int j = 0; j < 4; ++j) {
                        mbar_wait(&ready[j], (rpar >> j) & 1u);
                        rpar ^= (1u << j);
                        mbar_wait(&full[slot], use & 1u);
                        tc_fence_after();
                        const uint32_t b = smem_u32(ring + slot * 32768);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ts(td, ta + 64u * j + 16u * k, umma_smem_desc_sw128(b + 32u * k), idesc, (j > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&empty[slot]);
                        if (++slot == F16_NSLOTS) { slot = 0; ++use; }
                    }
                    umma_commit(done);
                }
        }
        __syncwarp();
        asm volatile("bar.sync 2, 544;" ::: "memory");  // epilogue warps are out of tensor memory
        tmem_dealloc(tbase, 512);
        return;
    }
    // ===== epilogue warps: q = lane quarter, u = which 16 columns of every K-chunk =====
    const int q = warp & 3, u = warp >> 2, r = 32 * q + lane;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    uint32_t done_par = 0;
    auto sync_epi = [&]() { asm volatile("bar.sync 1, 512;" ::: "memory"); };
    // 16 activations (K = 64 j + 16 u ..) -> 8 packed columns at 64 j + 16 u of region `reg`; publishes the warp's share of chunk j
    auto put = [&](int reg, int j, const float (&v)[16]) {
        uint32_t p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            p[i] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        tmem_st8(trow + 256u * reg + 64u * j + 16u * u, p);
    };
    // the warp's share of chunk j is in tensor memory
    auto publish = [&](int j) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[j]);
    };
    // the query point of row `tid` of a tile; called one tile AHEAD (the two dependent loads shade_list -> smp_xn would otherwise
    // cost two L2 / HBM latencies at the start of every tile, with all sixteen warps waiting at the barrier behind them)
    auto fetch = [&](int tile, int& sl, float (&xn)[3]) {
        sl = -1; xn[0] = 0.f; xn[1] = 0.f; xn[2] = 0.f;
        const int i = tile * UM + tid;
        if (tile < ntiles && i < n) {
            if (g.out) {                                            // lattice coordinates with the reference's arithmetic (sdf_meshing.py:25-38)
                sl = i;
                const int iz = i % g.N, iy = (i / g.N) % g.N, ix = i / (g.N * g.N);
                xn[0] = __fadd_rn(__fmul_rn((float)ix, g.voxel), -1.0f);
                xn[1] = __fadd_rn(__fmul_rn((float)iy, g.voxel), -1.0f);
                xn[2] = __fadd_rn(__fmul_rn((float)iz, g.voxel), -1.0f);
            } else { sl = w.shade_list[i]; xn[0] = w.smp_xn[3 * (size_t)sl]; xn[1] = w.smp_xn[3 * (size_t)sl + 1]; xn[2] = w.smp_xn[3 * (size_t)sl + 2]; }
        }
    };
    int sl_next = -1;
    float xn_next[3] = {0.f, 0.f, 0.f};
    if (tid < UM) fetch(blockIdx.x, sl_next, xn_next);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int sl = -1;
        if (tid < UM) {
            sl = sl_next;
            xs[tid][0] = xn_next[0]; xs[tid][1] = xn_next[1]; xs[tid][2] = xn_next[2]; xs[tid][3] = 0.f;
            fetch(tile + (int)gridDim.x, sl_next, xn_next);         // in flight while this tile is evaluated
        }
        sync_epi();
        {   // layer 0 (K = 3) on the FP32 pipe -> operand of layer 1 in R0
            const float x = xs[r][0], y = xs[r][1], z = xs[r][2];
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int col0 = 64 * j + 16 * u;
                float v[16];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {                          // 128-bit shared-memory loads: LDS and MUFU share the MIO queue
                    const int cc = col0 + 4 * i4;
                    const float4 wx = *reinterpret_cast<const float4*>(sW0 + cc), wy = *reinterpret_cast<const float4*>(sW0 + 256 + cc),
                                 wz = *reinterpret_cast<const float4*>(sW0 + 512 + cc), f4 = *reinterpret_cast<const float4*>(sF + cc),
                                 g4 = *reinterpret_cast<const float4*>(sG + cc);
                    v[4 * i4 + 0] = __sinf(fmaf(fmaf(wz.x, z, fmaf(wy.x, y, wx.x * x)), f4.x, g4.x));
                    v[4 * i4 + 1] = __sinf(fmaf(fmaf(wz.y, z, fmaf(wy.y, y, wx.y * x)), f4.y, g4.y));
                    v[4 * i4 + 2] = __sinf(fmaf(fmaf(wz.z, z, fmaf(wy.z, y, wx.z * x)), f4.z, g4.z));
                    v[4 * i4 + 3] = __sinf(fmaf(fmaf(wz.w, z, fmaf(wy.w, y, wx.w * x)), f4.w, g4.w));
                }
                put(0, j, v);
                publish(j);
            }
        }
        float dot = 0.f;
#pragma unroll 1
        for (int L = 1; L <= 5; ++L) {
            mbar_wait(done, done_par);
            done_par ^= 1u;
            __syncwarp();
            tc_fence_after();
            const int dreg = L & 1;
            const float* F = sF + L * 256;
            const float* G = sG + L * 256;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int col0 = 64 * j + 16 * u;
                float v[16];
                tmem_ld16(trow + 256u * dreg + (uint32_t)col0, v);
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float4 f4 = *reinterpret_cast<const float4*>(F + col0 + 4 * i4), g4 = *reinterpret_cast<const float4*>(G + col0 + 4 * i4);
                    v[4 * i4 + 0] = __sinf(fmaf(v[4 * i4 + 0], f4.x, g4.x));
                    v[4 * i4 + 1] = __sinf(fmaf(v[4 * i4 + 1], f4.y, g4.y));
                    v[4 * i4 + 2] = __sinf(fmaf(v[4 * i4 + 2], f4.z, g4.z));
                    v[4 * i4 + 3] = __sinf(fmaf(v[4 * i4 + 3], f4.w, g4.w));
                }
                if (L < 5) { put(dreg, j, v); publish(j); }      // in place: the 8 packed columns lie inside the 16 just read
                else {
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(sW6 + col0 + 4 * i4);
                        dot = fmaf(v[4 * i4 + 3], w4.w, fmaf(v[4 * i4 + 2], w4.z, fmaf(v[4 * i4 + 1], w4.y, fmaf(v[4 * i4], w4.x, dot))));
                    }
                }
            }
        }
        tc_fence_before();
        part[u][r] = dot;
        sync_epi();
        if (tid < UM && sl >= 0) {
            const float raw = ((part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid])) + __ldg(sd.b6);
            if (g.out) g.out[sl] = raw;
            else w.smp_sdf[sl] = sdf_to_metres(raw, fp.cmin, fp.cmax);
        }
        // (xs / part are rewritten only after the next sync_epi)
    }
    tc_fence_before();
    asm volatile("bar.sync 2, 544;" ::: "memory");
}

}  // namespace arah
