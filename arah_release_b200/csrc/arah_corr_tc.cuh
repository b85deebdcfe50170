// arah_corr_tc.cuh — k_corr_tc: one step of the per-sample correspondence search with the skinning MLP on the tensor
// cores in split precision (3xTF32 ~ fp32), 128 samples per tile.
//
// Same algorithm and bookkeeping as k_corr_step (utils/root_finding_utils.py:267-362 + utils/broyden.py of the reference);
// only the 128-wide hidden layers and the 128->25 output layer change engine:
//   D = A_hi.B_hi + A_lo.B_hi + A_hi.B_lo  with hi = RN_tf32(x), lo = RN_tf32(x - hi)  (arah_umma.cuh)
// A_hi / A_lo tiles live in shared memory (2 x 64 KB), weights arrive as [hi | lo] chunk images through the TMA ring,
// accumulators in TMEM (128 columns).  Layer 0 (3 -> 128) and all per-point math (hierarchical softmax, LBS blend,
// Broyden update) stay fp32 on the CUDA cores.
#pragma once
#include "arah_kernels.cuh"
#include "arah_umma.cuh"

namespace arah {

struct SkinTC {
    const float* Wt0;      // [3][128]
    const float* b[5];     // biases (b[4] padded to 32)
    const float* hid[3];   // layers 1..3: 4 chunks of [hi | lo] images, N = 128
    const float* out;      // layer 4: 4 chunks of [hi | lo] images, N = 32 (25 padded)
};

__host__ __device__ constexpr size_t corr_tc_smem_bytes() {
    return (size_t)(2 * 4 * A_CHUNK_FLOATS + 2 * RING_SLOT_FLOATS + UM * 32 + UM * 4) * 4 + 256 + 1024;
}

// skinning MLP for the 128 rows whose normalised inputs sit in xs; result: logits[r][0..31]
__device__ __forceinline__ void skin_tile_tc(const SkinTC& sk, const float (*xs)[4], float* A_hi, float* A_lo, URing& rg, uint64_t* done_bar,
                                             uint32_t& done_par, uint32_t tbase, float (*logits)[32]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, half = warp >> 2, r = 32 * q + lane;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    auto handoff = [&]() { fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after(); };
    {   // layer 0 on the FP32 pipe: 64 columns per thread
        const float x = xs[r][0], y = xs[r][1], z = xs[r][2];
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
            const int col0 = 64 * half + 32 * b;
            float h[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int cc = col0 + i;
                const float a = fmaf(__ldg(sk.Wt0 + 256 + cc), z, fmaf(__ldg(sk.Wt0 + 128 + cc), y, __ldg(sk.Wt0 + cc) * x)) + __ldg(sk.b[0] + cc);
                h[i] = softplus100_fast(a);
            }
            a_store_chunk_split(A_hi, A_lo, r, col0 / 32, h);
        }
    }
    handoff();
    for (int l = 1; l < 4; ++l) {
        if (tid == 0) umma_layer_issue_x3(rg, A_hi, A_lo, sk.hid[l - 1], 4, 128, tbase, done_bar);
        mbar_wait(done_bar, done_par);
        done_par ^= 1u;
        __syncwarp();
        tc_fence_after();
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
            const int col0 = 64 * half + 32 * b;
            float v[32];
            tmem_ld32(trow + (uint32_t)col0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = softplus100_fast(v[i] + __ldg(sk.b[l] + col0 + i));
            a_store_chunk_split(A_hi, A_lo, r, col0 / 32, v);
        }
        handoff();
    }
    if (tid == 0) umma_layer_issue_x3(rg, A_hi, A_lo, sk.out, 4, 32, tbase, done_bar);
    mbar_wait(done_bar, done_par);
    done_par ^= 1u;
    __syncwarp();
    tc_fence_after();
    if (half == 0) {
        float v[32];
        tmem_ld32(trow, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) logits[r][i] = v[i] + __ldg(sk.b[4] + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}

__global__ void __launch_bounds__(256, 1) k_corr_tc(FrameParams fp, SkinTC sk, Work w, int iter) {
    extern __shared__ uint8_t raw_smem[];
    const int n = (iter < 0) ? w.counters[C_ON] : w.counters[C_CORR + iter];
    if ((int)blockIdx.x * UM >= n) return;
    const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;
    float* sm = reinterpret_cast<float*>(raw_smem + (base - smem_u32(raw_smem)));
    float* A_hi = sm;
    float* A_lo = A_hi + 4 * A_CHUNK_FLOATS;
    float* ring = A_lo + 4 * A_CHUNK_FLOATS;
    float (*logits)[32] = reinterpret_cast<float (*)[32]>(ring + 2 * RING_SLOT_FLOATS);
    float (*xs)[4] = reinterpret_cast<float (*)[4]>(reinterpret_cast<float*>(logits) + UM * 32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(xs) + UM * 4);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tslot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    URing rg; rg.buf = ring; rg.full = bars; rg.empty = bars + 2; rg.fill_cnt = 0; rg.mma_cnt = 0;
    uint32_t done_par = 0;
    const int* list = (iter <= 0) ? nullptr : ((iter & 1) ? w.listB : w.listA);
    int* next = (iter & 1) ? w.listA : w.listB;
    for (int tile = blockIdx.x; tile * UM < n; tile += gridDim.x) {
        int id = -1;
        BroydenState<3> st;
        float dx[3];
        if (tid < UM) {
            const int i = tile * UM + tid;
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) {
                id = list ? list[i] : i;
                state_load(st, &w.corr_state[id]);
                if (iter >= 0) broyden_advance<3>(st, dx);
                normalize3(fp, st.x, xn);
            }
            xs[tid][0] = xn[0]; xs[tid][1] = xn[1]; xs[tid][2] = xn[2]; xs[tid][3] = 0.f;
        }
        __syncthreads();
        skin_tile_tc(sk, xs, A_hi, A_lo, rg, &bars[4], done_par, tbase, logits);
        if (tid < UM) {
            bool active = false;
            if (id >= 0) {
                float T12[12], xb[3], g[3];
                skin_point(fp, logits[tid], st.x, T12, xb);
#pragma unroll
                for (int k = 0; k < 3; ++k) g[k] = xb[k] - st.tgt[k];
                if (iter < 0) {
                    float A3[9], Ai[9], Tinit[12];
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                        for (int c = 0; c < 3; ++c) A3[rr * 3 + c] = T12[rr * 4 + c];
                    invert3(A3, Ai);
#pragma unroll
                    for (int e = 0; e < 12; ++e) Tinit[e] = st.best_T[e];
                    const float x0[3] = {st.x[0], st.x[1], st.x[2]};
                    const int owner = st.owner;
                    const float tg[3] = {st.tgt[0], st.tgt[1], st.tgt[2]};
                    broyden_begin<3>(st, x0, g, Ai, Tinit);
                    st.owner = owner; st.tgt[0] = tg[0]; st.tgt[1] = tg[1]; st.tgt[2] = tg[2];
                    st.g_evals = 2;
                    state_store(&w.corr_state[id], st);
                } else {
                    active = broyden_update<3>(st, dx, g, T12);
                    if (iter + 1 >= BROYDEN_ITERS) active = false;
                    if (active) state_store(&w.corr_state[id], st);
                    else corr_finalize(fp, w, st);
                }
            }
            if (iter >= 0) {
                if (iter + 1 < BROYDEN_ITERS) warp_append(active, id, next, &w.counters[C_CORR + iter + 1]);
                const bool done = (id >= 0) && !active;
                warp_append(done && st.best_n < CVG_THRESH, done ? st.owner : 0, w.shade_list, &w.counters[C_SHADE]);
                warp_stat_add(done ? st.g_evals : 0, &w.counters[C_STAT_CORR_EVALS]);
            }
        }
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128);
}

}  // namespace arah
