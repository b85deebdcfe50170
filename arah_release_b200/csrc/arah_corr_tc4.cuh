// arah_corr_tc4.cuh — k_corr_tc4 = k_corr_tc3 + 2-CTA cluster sharing (multicasting) the skinning-weight stream
// (see arah_shade_tc4.cuh for the protocol: each CTA fetches half of every [hi | lo] chunk image and multicasts it to both
// shared memories; ring slots are recycled when both MMA warps have retired them).
#pragma once
#include "arah_corr_tc3.cuh"
#include "arah_shade_tc4.cuh"

namespace arah {

__global__ void __launch_bounds__(TC3_THREADS, 1) k_corr_tc4(FrameParams fp, SkinTC sk, Work w, int iter) {
    extern __shared__ uint8_t raw_smem[];
    const int n = (iter < 0) ? w.counters[C_ON] : w.counters[C_CORR + iter];
    const int ntiles_mine = (n + UM - 1) / UM;
    const int first = (int)(blockIdx.x & ~1u);                       // the pair's even CTA fixes the tile count of both CTAs
    if (first >= ntiles_mine) return;
    const int ntrips = (ntiles_mine - 1 - first) / (int)gridDim.x + 1;   // tiles per CTA (the odd CTA's last one may be padding)
    const uint32_t cta_rank = cluster_ctarank();
    const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;
    float* sm = reinterpret_cast<float*>(raw_smem + (base - smem_u32(raw_smem)));
    float* ring = sm;
    float (*logits)[LGS] = reinterpret_cast<float (*)[LGS]>(ring + TC3_NSLOTS * RING_SLOT_FLOATS);   // [2*UM][33]: tiles A, B (odd stride: conflict-free)
    float (*xs)[4] = reinterpret_cast<float (*)[4]>(reinterpret_cast<float*>(logits) + 2 * UM * LGS);  // [2*UM][4]
    float* sB = reinterpret_cast<float*>(xs) + 2 * UM * 4;        // bone transforms [24][16]
    float* sW0 = sB + 24 * 16;                                    // layer-0 weights [3][128]
    float* sb = sW0 + 3 * 128;                                    // biases: 4 x 128 then 32  (sb + 128*l)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb + 5 * 128);
    uint64_t* full = bars;
    uint64_t* empty = bars + TC3_NSLOTS;
    uint64_t* ready = bars + 2 * TC3_NSLOTS;     // [4]
    uint64_t* done_bar = ready + 4;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(done_bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < TC3_NSLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }
        for (int i = 0; i < 4; ++i) mbar_init(&ready[i], 4);
        mbar_init(done_bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tslot, 512);
    for (int i = tid; i < 24 * 16; i += TC3_THREADS) sB[i] = __ldg(fp.bone_T + i);
    for (int i = tid; i < 3 * 128; i += TC3_THREADS) sW0[i] = __ldg(sk.Wt0 + i);
    for (int i = tid; i < 4 * 128; i += TC3_THREADS) sb[i] = __ldg(sk.b[i >> 7] + (i & 127));
    if (tid < 32) sb[512 + tid] = __ldg(sk.b[4] + tid);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tbase = *tslot;

    if (warp == 8) {                                    // ===== TMA producer =====
        if (lane == 0) {
            uint32_t slot = 0, use = 0;
            for (int trip = 0; trip < ntrips; ++trip) {
                for (int s = 0; s < 4; ++s) {
                    const float* wsrc = (s < 3) ? sk.hid[s] : sk.out;
                    const uint32_t bytes = (s < 3) ? 32768u : 8192u, hb = bytes >> 1;
                    for (int i = 0; i < 4; ++i) {
                        const int c = seg_chunk(1, i);
                        if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1u);
                        mbar_expect_tx(&full[slot], bytes);
                        bulk_g2s_mc2(reinterpret_cast<char*>(ring + slot * RING_SLOT_FLOATS) + cta_rank * hb,
                                     reinterpret_cast<const char*>(wsrc) + (size_t)c * bytes + cta_rank * hb, hb, &full[slot]);
                        if (++slot == TC3_NSLOTS) { slot = 0; ++use; }
                    }
                }
            }
        }
        cluster_sync_exit();
        return;
    }
    if (warp == 9) {                                    // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t slot = 0, use = 0, rpar = 0;
            for (int trip = 0; trip < ntrips; ++trip) {
                for (int s = 0; s < 4; ++s) {
                    const int N = (s < 3) ? 128 : 32;
                    const uint32_t idesc = umma_idesc_tf32(UM, N);
                    const uint32_t td = tbase + ((s & 1) ? 384u : 256u);
                    for (int i = 0; i < 4; ++i) {
                        const int c = seg_chunk(1, i);
                        mbar_wait(&ready[c], (rpar >> c) & 1u);
                        rpar ^= (1u << c);
                        mbar_wait(&full[slot], use & 1u);
                        tc_fence_after();
                        const uint32_t bh = smem_u32(ring + slot * RING_SLOT_FLOATS), bl = bh + (uint32_t)N * UK * 4;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t col = (uint32_t)(c * UK + k * 8), ko = k * 32;
                            umma_tf32_ts(td, tbase + 128u + col, umma_smem_desc_sw128(bh + ko), idesc, (i > 0 || k > 0) ? 1u : 0u);   // A_lo . B_hi
                            umma_tf32_ts(td, tbase + col, umma_smem_desc_sw128(bl + ko), idesc, 1u);                                  // A_hi . B_lo
                            umma_tf32_ts(td, tbase + col, umma_smem_desc_sw128(bh + ko), idesc, 1u);                                  // A_hi . B_hi
                        }
                        umma_commit_mc2(&empty[slot]);
                        if (++slot == TC3_NSLOTS) { slot = 0; ++use; }
                    }
                    umma_commit(done_bar);
                }
            }
        }
        cluster_sync_exit();
        return;
    }
    // ===== compute warps =====
    const int q = warp & 3, half = warp >> 2, r = 32 * q + lane;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    uint32_t done_par = 0;
    auto a_put = [&](int chunk, const float (&v)[32]) {
        a_tmem_store_split(trow + 32u * chunk, trow + 128u + 32u * chunk, v);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[chunk]);
    };
    auto wait_done = [&]() {
        mbar_wait(done_bar, done_par);
        done_par ^= 1u;
        __syncwarp();
        tc_fence_after();
    };
    const int* list = (iter <= 0) ? nullptr : ((iter & 1) ? w.listB : w.listA);
    int* next = (iter & 1) ? w.listA : w.listB;
    PhaseClk pc; pc.start((tid == 32) ? w.phase_clk : nullptr);
    // Two tiles (A, B) per trip: their MLPs run back to back on all 8 warps, their per-point phases (state gather, hierarchical
    // softmax, LBS blend, Broyden update) run CONCURRENTLY: warps 0-3 own tile A's points, warps 4-7 tile B's.
    const int sub = tid >> 7, pt = tid & (UM - 1);                 // which tile of the pair / which point this thread owns
    for (int trip = 0; trip < ntrips; trip += 2) {
        const int tileA = (int)blockIdx.x + trip * (int)gridDim.x;
        const int tileB = tileA + (int)gridDim.x;
        const bool haveB = trip + 1 < ntrips;                          // B exists as a (possibly all-padding) tile of this CTA
        const int my_tile = sub ? (haveB ? tileB : ntiles_mine) : tileA;
        int id = -1;
        BroydenState<3> st;
        float dx[3];
        {
            const int i = my_tile * UM + pt;
            float xn[3] = {0.f, 0.f, 0.f};
            if (my_tile < ntiles_mine && i < n) {
                id = list ? list[i] : i;
                state_load(st, &w.corr_state[id]);
                if (iter >= 0) broyden_advance<3>(st, dx);
                normalize3(fp, st.x, xn);
            }
            xs[tid][0] = xn[0]; xs[tid][1] = xn[1]; xs[tid][2] = xn[2]; xs[tid][3] = 0.f;
        }
        cta_sync_compute();
        pc.mark(0);
        for (int t = 0; t < 2; ++t) {
            if (t == 1 && !haveB) break;
            {   // layer 0 (3 -> 128) on the FP32 pipe
                const float x = xs[t * UM + r][0], y = xs[t * UM + r][1], z = xs[t * UM + r][2];
#pragma unroll 1
                for (int b = 0; b < 2; ++b) {
                    const int col0 = 64 * half + 32 * b;
                    float h[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int cc = col0 + i;
                        h[i] = softplus100_fast(fmaf(sW0[256 + cc], z, fmaf(sW0[128 + cc], y, sW0[cc] * x)) + sb[cc]);
                    }
                    a_put(col0 / 32, h);
                }
            }
            pc.mark(1);
            for (int l = 1; l < 4; ++l) {
                wait_done();
                pc.mark(2);
                const uint32_t tD = trow + ((l & 1) ? 256u : 384u);
#pragma unroll 1
                for (int b = 0; b < 2; ++b) {
                    const int col0 = 64 * half + 32 * b;
                    float v[32];
                    tmem_ld32(tD + (uint32_t)col0, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = softplus100_fast(v[i] + sb[128 * l + col0 + i]);
                    a_put(col0 / 32, v);
                }
                pc.mark(3);
            }
            wait_done();                                                 // output layer: D = Db; X is free for the next tile's layer 0
            if (half == 0) {
                float v[32];
                tmem_ld32(trow + 384u, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) logits[t * UM + r][i] = v[i] + sb[512 + i];
            }
            tc_fence_before();
            cta_sync_compute();                                          // Db read before the next tile's second GEMM rewrites it
            tc_fence_after();
            pc.mark(4);
        }
        {
            bool active = false;
            if (id >= 0) {
                float T12[12], xb[3], g[3], lg[25], wj[NJ];
#pragma unroll
                for (int k = 0; k < 25; ++k) lg[k] = logits[tid][k] * 20.0f;
                hierarchical_softmax(lg, wj);
                blend_T(wj, sB, T12, nullptr);
                apply_T(T12, st.x, xb);
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
        pc.mark(5);
    }
    tc_fence_before();
    cta_sync_compute();
    if (warp == 0) tmem_dealloc(tbase, 512);
    cluster_sync_exit();
}

}  // namespace arah
