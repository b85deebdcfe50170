// arah_corr_tc2.cuh — k_corr_tc2: correspondence-search step with the skinning MLP in 3xTF32 on engine v2
// (arah_tc2.cuh): A_hi / A_lo in TMEM columns [0,128) / [128,256), accumulators in [256,384), weights ([hi | lo] chunk
// images) through the 6-slot ring filled by the producer warp.  Algorithm and bookkeeping: k_corr_step / k_corr_tc.
#pragma once
#include "arah_corr_tc.cuh"
#include "arah_tc2.cuh"

namespace arah {

__host__ __device__ constexpr size_t corr_tc2_smem_bytes() {
    return (size_t)(TC_NSLOTS * RING_SLOT_FLOATS + UM * 32 + UM * 4) * 4 + 256 + 1024;
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_corr_tc2(FrameParams fp, SkinTC sk, Work w, int iter) {
    extern __shared__ uint8_t raw_smem[];
    const int n = (iter < 0) ? w.counters[C_ON] : w.counters[C_CORR + iter];
    if ((int)blockIdx.x * UM >= n) return;
    const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;
    float* sm = reinterpret_cast<float*>(raw_smem + (base - smem_u32(raw_smem)));
    float* ring = sm;
    float (*logits)[32] = reinterpret_cast<float (*)[32]>(ring + TC_NSLOTS * RING_SLOT_FLOATS);
    float (*xs)[4] = reinterpret_cast<float (*)[4]>(reinterpret_cast<float*>(logits) + UM * 32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(xs) + UM * 4);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * TC_NSLOTS + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    TCRing rg; rg.buf = ring; rg.full = bars; rg.empty = bars + TC_NSLOTS;
    uint64_t* done_bar = bars + 2 * TC_NSLOTS;
    if (tid == 0) { tcring_init(rg); mbar_init(done_bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    if (warp == 8) {                                    // ===== TMA producer warp =====
        if (lane == 0) {
            RingPos pp; pp.slot = 0; pp.use = 0;
            for (int tile = blockIdx.x; tile * UM < n; tile += gridDim.x) {
                for (int l = 0; l < 3; ++l) tcring_produce(rg, pp, sk.hid[l], 4, 32768u);
                tcring_produce(rg, pp, sk.out, 4, 8192u);
            }
        }
        return;
    }
    const int q = warp & 3, half = warp >> 2, r = 32 * q + lane;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    const uint32_t tAhi = trow, tAlo = trow + 128u, tD = trow + 256u;
    RingPos cp; cp.slot = 0; cp.use = 0;
    uint32_t done_par = 0;
    auto handoff = [&]() { tmem_st_wait(); tc_fence_before(); cta_sync_compute(); tc_fence_after(); };
    const int* list = (iter <= 0) ? nullptr : ((iter & 1) ? w.listB : w.listA);
    int* next = (iter & 1) ? w.listA : w.listB;
    PhaseClk pc; pc.start((tid == 32) ? w.phase_clk : nullptr);
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
        cta_sync_compute();
        pc.mark(0);                                   // gather + advance
        {   // layer 0 (3 -> 128) on the FP32 pipe, 64 columns per thread
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
                a_tmem_store_split(tAhi + (uint32_t)col0, tAlo + (uint32_t)col0, h);
            }
        }
        handoff();
        pc.mark(1);                                   // layer 0
        for (int l = 1; l < 4; ++l) {
            if (tid == 0) tcring_mma_layer_x3(rg, cp, tbase, tbase + 128u, 4, 128, tbase + 256u, done_bar);
            mbar_wait(done_bar, done_par);
            done_par ^= 1u;
            __syncwarp();
            tc_fence_after();
            pc.mark(2);                               // waiting for the layer's MMAs
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
                const int col0 = 64 * half + 32 * b;
                float v[32];
                tmem_ld32(tD + (uint32_t)col0, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = softplus100_fast(v[i] + __ldg(sk.b[l] + col0 + i));
                a_tmem_store_split(tAhi + (uint32_t)col0, tAlo + (uint32_t)col0, v);
            }
            handoff();
            pc.mark(3);                               // hidden-layer epilogue
        }
        if (tid == 0) tcring_mma_layer_x3(rg, cp, tbase, tbase + 128u, 4, 32, tbase + 256u, done_bar);
        mbar_wait(done_bar, done_par);
        done_par ^= 1u;
        __syncwarp();
        tc_fence_after();
        if (half == 0) {
            float v[32];
            tmem_ld32(tD, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) logits[r][i] = v[i] + __ldg(sk.b[4] + i);
        }
        tc_fence_before();
        cta_sync_compute();
        tc_fence_after();
        pc.mark(4);                                   // output layer (MMA wait + logits)
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
        cta_sync_compute();
        pc.mark(5);                                   // per-point phase (hsoftmax, blend, Broyden, state I/O)
    }
    tc_fence_before();
    cta_sync_compute();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace arah
