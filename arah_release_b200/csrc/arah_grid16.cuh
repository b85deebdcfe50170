// arah_grid16.cuh — k_sdf_grid16: the canonical SDF on the N^3 lattice over [-1, 1]^3 (utils/sdf_meshing.py:13-58)
// on the fp16 split-precision engine (arah_sdf16.cuh): same values to ~1e-6 as the fp32 FFMA path, at twice the MMA rate of
// round 1's 3xTF32 kernel and with sixteen epilogue warps.  Lattice coordinates are generated in the kernel with the reference's
// own arithmetic (index * voxel_size + origin, one fp32 rounding per operation, sdf_meshing.py:25-38);
// out[(ix * N + iy) * N + iz] = raw network output (what `decoder(model_input)` returns, :49-54).
#pragma once
#include "arah_sdf16.cuh"

namespace arah {

__host__ __device__ constexpr size_t sdf_grid16_smem_bytes() { return (size_t)S16_NSLOTS * S16_SLOT_BYTES + (size_t)(4 * UM + 8) * 4 + sizeof(S16Ctl) + 64; }

// list != nullptr (the refinement pass of the banded lattice): the points are list[0 .. *list_n) (lattice indices); a refined
// value that differs from the coarse value already in out[] by more than eps counts into *violations.
__global__ void __launch_bounds__(S16_THREADS, 1) k_sdf_grid16(SdfF16 sd, int N, float voxel, long long n_total_, float* __restrict__ out,
                                                               const int* __restrict__ list, const int* __restrict__ list_n, float eps,
                                                               int* __restrict__ violations) {
    extern __shared__ __align__(1024) uint8_t raw_smem[];
    const long long n_total = list ? (long long)*list_n : n_total_;
    const long long ntiles = (n_total + UM - 1) / UM;
    if ((long long)blockIdx.x >= ntiles) return;
    if (smem_u32(raw_smem) & 1023u) __trap();
    uint8_t* ring = raw_smem;
    float (*part)[UM] = reinterpret_cast<float (*)[UM]>(ring + S16_NSLOTS * S16_SLOT_BYTES);
    float* sInv = reinterpret_cast<float*>(part) + 4 * UM;
    S16Ctl* ctl = reinterpret_cast<S16Ctl*>(sInv + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s16_ctl_init(ctl);
    if (warp == 17) tmem_alloc(&ctl->tslot, 512);
    if (tid < 5) sInv[tid] = __ldg(sd.scale + 2 * tid + 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = ctl->tslot;
    if (warp == 16) {                                       // TMA producer: the tile count is known, no speculation needed
        if (lane == 0) { S16Prod p; for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) s16_produce_sdf(ring, ctl, p, sd, false); }
        return;
    }
    if (warp == 17) {                                       // MMA issuer
        if (lane == 0) { S16Mma m; for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) s16_mma_sdf(ring, ctl, m, tbase); }
        __syncwarp();
        s16_sync_exit();
        tmem_dealloc(tbase, 512);
        return;
    }
    const int q = warp & 3, u = warp >> 2, r = 32 * q + lane;
    uint32_t done_par = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long e = tile * UM + r;                  // the lattice point of this thread's row
        const long long i = (list && e < n_total) ? (long long)list[e] : e;
        float x = 0.f, y = 0.f, z = 0.f;
        if (e < n_total) {
            const int iz = (int)(i % N), iy = (int)((i / N) % N), ix = (int)(i / ((long long)N * N));
            x = __fadd_rn(__fmul_rn((float)ix, voxel), -1.0f);
            y = __fadd_rn(__fmul_rn((float)iy, voxel), -1.0f);
            z = __fadd_rn(__fmul_rn((float)iz, voxel), -1.0f);
        }
        const float dot = s16_compute_sdf(sd, x, y, z, ctl, done_par, tbase, sInv);
        part[u][r] = dot;
        s16_sync();
        if (u == 0 && e < n_total) {
            const float val = ((part[0][r] + part[1][r]) + (part[2][r] + part[3][r])) + __ldg(sd.b6);
            if (list && !(fabsf(val - out[i]) <= eps)) atomicAdd(violations, 1);
            out[i] = val;
        }
        s16_sync();                                         // part is rewritten by the next tile
    }
    tc_fence_before();
    s16_sync_exit();
}

// ---- banded lattice: which points need the split-precision value ------------------------------------------------------------
// A cell can contribute to the iso-surface only if its 8 corner values straddle `level`.  With |coarse - exact| <= eps at every
// point, a cell whose coarse corner values satisfy min - eps > level or max + eps < level cannot straddle it, and all its corners
// keep the sign (relative to level) of their exact values.  Every other cell gets all 8 corners refined, so marching cubes sees
// exact values wherever it interpolates and exact signs everywhere: its output is bit-identical to that of the full lattice.
__global__ void k_grid_band_flag(const float* __restrict__ vol, int N, float level, float eps, uint8_t* __restrict__ flag) {
    const int iz = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y, ix = blockIdx.z;
    if (iz >= N - 1) return;
    const size_t b = ((size_t)ix * N + iy) * N + iz, sx = (size_t)N * N, sy = (size_t)N;
    const size_t o[8] = {b, b + 1, b + sy, b + sy + 1, b + sx, b + sx + 1, b + sx + sy, b + sx + sy + 1};
    float mn = 3.4e38f, mx = -3.4e38f;
    bool finite = true;
#pragma unroll
    for (int c = 0; c < 8; ++c) { const float v = vol[o[c]]; mn = fminf(mn, v); mx = fmaxf(mx, v); finite = finite && (fabsf(v) < 3.0e38f); }
    if (!finite || (mn - eps <= level && level <= mx + eps)) {
#pragma unroll
        for (int c = 0; c < 8; ++c) flag[o[c]] = 1;
    }
}
__global__ void k_grid_band_list(const uint8_t* __restrict__ flag, int n, int* __restrict__ list, int* __restrict__ counter) {
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {     // warp-uniform trip count (n rounded up by the caller)
        const int i = base + threadIdx.x;
        warp_append(i < n && flag[i] != 0, i, list, counter);
    }
}

}  // namespace arah
