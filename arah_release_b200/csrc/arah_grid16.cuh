// arah_grid16.cuh — k_sdf_grid16: the canonical SDF on the N^3 lattice over [-1, 1]^3 (utils/sdf_meshing.py:13-58)
// on the fp16 split-precision engine (arah_sdf16.cuh): same values to ~1e-6 as the fp32 FFMA path, at twice the MMA rate of
// round 1's 3xTF32 kernel and with sixteen epilogue warps.  Lattice coordinates are generated in the kernel with the reference's
// own arithmetic (index * voxel_size + origin, one fp32 rounding per operation, sdf_meshing.py:25-38);
// out[(ix * N + iy) * N + iz] = raw network output (what `decoder(model_input)` returns, :49-54).
#pragma once
#include "arah_sdf16.cuh"

namespace arah {

__host__ __device__ constexpr size_t sdf_grid16_smem_bytes() { return (size_t)S16_NSLOTS * S16_SLOT_BYTES + (size_t)(4 * UM + 8) * 4 + sizeof(S16Ctl) + 64; }

__global__ void __launch_bounds__(S16_THREADS, 1) k_sdf_grid16(SdfF16 sd, int N, float voxel, long long n_total, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t raw_smem[];
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
        const long long i = tile * UM + r;                  // the lattice point of this thread's row
        float x = 0.f, y = 0.f, z = 0.f;
        if (i < n_total) {
            const int iz = (int)(i % N), iy = (int)((i / N) % N), ix = (int)(i / ((long long)N * N));
            x = __fadd_rn(__fmul_rn((float)ix, voxel), -1.0f);
            y = __fadd_rn(__fmul_rn((float)iy, voxel), -1.0f);
            z = __fadd_rn(__fmul_rn((float)iz, voxel), -1.0f);
        }
        const float dot = s16_compute_sdf(sd, x, y, z, ctl, done_par, tbase, sInv);
        part[u][r] = dot;
        s16_sync();
        if (u == 0 && i < n_total) out[i] = ((part[0][r] + part[1][r]) + (part[2][r] + part[3][r])) + __ldg(sd.b6);
        s16_sync();                                         // part is rewritten by the next tile
    }
    tc_fence_before();
    s16_sync_exit();
}

}  // namespace arah
