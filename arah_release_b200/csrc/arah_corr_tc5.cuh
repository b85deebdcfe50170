// arah_corr_tc5.cuh — k_corr_tc5: correspondence-search step with the two tiles of a trip INTERLEAVED on the tensor pipe.
//
// k_corr_tc3/tc4 run the skinning MLP of tile A and then of tile B; inside one tile every layer's epilogue (bias + softplus +
// hi/lo split, ~2.7 k cycles for 8 warps) and the MMA tail that follows it (~2.4 k) are serialised: tensor pipe 26 % active
// (profiles/r01_k_corr_tc4_ncu.md).  Two fully resident 3xTF32 tiles (X_hi + X_lo + D = 384 TMEM columns each) do not fit in
// 512 columns, but the activation columns can be TIME-SHARED: only the accumulators are per tile.
//
//   TMEM: X = [0,128) hi | [128,256) lo  (input of whichever job runs), D_A = [256,384), D_B = [384,512).
//   jobs issued by the MMA warp per trip: (A,l1) (B,l1) (A,l2) (B,l2) (A,l3) (B,l3) (A,out) (B,out).
//   compute warps: while job J(k-1) (other tile) runs they turn the accumulators of J(k-2) (same tile, previous layer) into the
//   input of job J(k) in REGISTERS (64 values per thread), then store it to X chunk by chunk as J(k-1) releases the chunks
//   (xfree[c], committed by the MMA warp after the MMAs that read chunk c) and publish each chunk (ready[c]): J(k) can start
//   the moment J(k-1) ends.  An epilogue always overlaps the other tile's MMAs and the store hand-off overlaps them too.
// Weight ring, 2-CTA cluster multicast, bookkeeping and arithmetic are those of k_corr_tc4 (same per-row results: a row's
// activations, MMAs and accumulation order do not depend on the schedule).  The Broyden state is not kept in registers across
// the MLP phase (it is re-read from L2 for the per-point phase) to make room for the 64 staged activations.
#pragma once
#include "arah_corr_tc4.cuh"

namespace arah {

__global__ void __launch_bounds__(TC3_THREADS, 1) k_corr_tc5(FrameParams fp, SkinTC sk, Work w, int iter) {
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
    float (*logits)[LGS] = reinterpret_cast<float (*)[LGS]>(ring + TC3_NSLOTS * RING_SLOT_FLOATS);   // [2*UM][33]: tiles A, B
    float (*xs)[4] = reinterpret_cast<float (*)[4]>(reinterpret_cast<float*>(logits) + 2 * UM * LGS);  // [2*UM][4]
    float* sB = reinterpret_cast<float*>(xs) + 2 * UM * 4;        // bone transforms [24][16]
    float* sW0 = sB + 24 * 16;                                    // layer-0 weights [3][128]
    float* sb = sW0 + 3 * 128;                                    // biases: 4 x 128 then 32  (sb + 128*l)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb + 5 * 128);
    uint64_t* full = bars;
    uint64_t* empty = bars + TC3_NSLOTS;
    uint64_t* ready = bars + 2 * TC3_NSLOTS;     // [4] X chunk c stored by its 4 warps
    uint64_t* done = ready + 4;                  // [2] job of tile A / B complete (its D ready)
    uint64_t* xfree = done + 2;                  // [4] the running job has finished reading X chunk c
    uint32_t* tslot = reinterpret_cast<uint32_t*>(xfree + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < TC3_NSLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }
        for (int i = 0; i < 4; ++i) mbar_init(&ready[i], 4);
        mbar_init(&done[0], 1); mbar_init(&done[1], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&xfree[i], 1);
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

    if (warp == 8) {                                    // ===== TMA producer: one pass over the skinning weights per job =====
        if (lane == 0) {
            uint32_t slot = 0, use = 0;
            for (int trip = 0; trip < ntrips; trip += 2) {
                const int ntile = (trip + 1 < ntrips) ? 2 : 1;
                for (int s = 0; s < 4; ++s) {
                    const float* wsrc = (s < 3) ? sk.hid[s] : sk.out;
                    const uint32_t bytes = (s < 3) ? 32768u : 8192u, hb = bytes >> 1;
                    for (int t = 0; t < ntile; ++t)
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
            for (int trip = 0; trip < ntrips; trip += 2) {
                const int ntile = (trip + 1 < ntrips) ? 2 : 1;
                for (int s = 0; s < 4; ++s) {
                    const int N = (s < 3) ? 128 : 32;
                    const uint32_t idesc = umma_idesc_tf32(UM, N);
                    for (int t = 0; t < ntile; ++t) {
                        const uint32_t td = tbase + (t ? 384u : 256u);
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
                            umma_commit(&xfree[c]);
                            if (++slot == TC3_NSLOTS) { slot = 0; ++use; }
                        }
                        umma_commit(&done[t]);
                    }
                }
            }
        }
        cluster_sync_exit();
        return;
    }
    // ===== compute warps =====
    const int q = warp & 3, half = warp >> 2, r = 32 * q + lane;
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    uint32_t done_par = 0, xpar = 0;                                   // parity bits of done[t] / xfree[c]
    bool x_virgin = true;                                              // nothing has read X yet (first job of the kernel)
    float v[2][32];                                                    // the staged input of the next job: columns 64 half + 32 b + i
    auto wait_done = [&](int t) {
        mbar_wait(&done[t], (done_par >> t) & 1u);
        done_par ^= (1u << t);
        __syncwarp();
        tc_fence_after();
    };
    // v -> X (hi | lo), chunk by chunk: wait until the running job has read the chunk, store, publish
    auto store_x = [&]() {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int chunk = 2 * half + b;
            if (!x_virgin) {
                mbar_wait(&xfree[chunk], (xpar >> chunk) & 1u);
                xpar ^= (1u << chunk);
                __syncwarp();
                tc_fence_after();
            }
            a_tmem_store_split(trow + 32u * chunk, trow + 128u + 32u * chunk, v[b]);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[chunk]);
        }
        // xfree[c] of the chunks owned by the OTHER half also completes once per job: keep their parity in step
        if (!x_virgin) xpar ^= (3u << (2 * (1 - half)));
        x_virgin = false;
    };
    // all compute warps have finished reading a D region (the job triggered by the next store_x overwrites it)
    auto sync_reads = [&]() { tc_fence_before(); cta_sync_compute(); tc_fence_after(); };
    auto layer0 = [&](int t) {                                         // 3 -> 128 on the FP32 pipe, into v
        const float x = xs[t * UM + r][0], y = xs[t * UM + r][1], z = xs[t * UM + r][2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int col0 = 64 * half + 32 * b;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int cc = col0 + i;
                v[b][i] = softplus100_fast(fmaf(sW0[256 + cc], z, fmaf(sW0[128 + cc], y, sW0[cc] * x)) + sb[cc]);
            }
        }
    };
    auto epilogue = [&](int t, int l) {                                // D_t (output of layer l - 1 ... i.e. pre-activation of layer l) -> v
        const uint32_t tD = trow + (t ? 384u : 256u);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const int col0 = 64 * half + 32 * b;
            tmem_ld32(tD + (uint32_t)col0, v[b]);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[b][i] = softplus100_fast(v[b][i] + sb[128 * l + col0 + i]);
        }
        tc_fence_before();
    };
    auto take_logits = [&](int t) {
        if (half == 0) {
            float o[32];
            tmem_ld32(trow + (t ? 384u : 256u), o);
#pragma unroll
            for (int i = 0; i < 32; ++i) logits[t * UM + r][i] = o[i] + sb[512 + i];
        }
        tc_fence_before();
    };
    const int* list = (iter <= 0) ? nullptr : ((iter & 1) ? w.listB : w.listA);
    int* next = (iter & 1) ? w.listA : w.listB;
    PhaseClk pc; pc.start((tid == 32) ? w.phase_clk : nullptr);
    const int sub = tid >> 7, pt = tid & (UM - 1);                 // which tile of the pair / which point this thread owns
    // gather of a trip: the query point of this thread's sample (x + update, normalised) -> xs; returns the sample id (-1: none).
    // The state itself is re-read for the per-point phase.
    auto gather = [&](int trip) -> int {
        const int tileA = (int)blockIdx.x + trip * (int)gridDim.x;
        const bool hb = trip + 1 < ntrips;
        const int my_tile = sub ? (hb ? tileA + (int)gridDim.x : ntiles_mine) : tileA;
        const int i = my_tile * UM + pt;
        int sid = -1;
        float xn[3] = {0.f, 0.f, 0.f};
        if (my_tile < ntiles_mine && i < n) {
            sid = list ? list[i] : i;
            const BroydenState<3>* sp = &w.corr_state[sid];
            float x3[3] = {sp->x[0], sp->x[1], sp->x[2]};
            if (iter >= 0) { x3[0] += sp->upd[0]; x3[1] += sp->upd[1]; x3[2] += sp->upd[2]; }      // broyden_advance
            normalize3(fp, x3, xn);
        }
        xs[tid][0] = xn[0]; xs[tid][1] = xn[1]; xs[tid][2] = xn[2]; xs[tid][3] = 0.f;
        return sid;
    };
    // Software pipeline across trips: the first two jobs of trip i+1 are fed (gather, layer 0, stores) BEFORE the per-point
    // phase of trip i, so the hierarchical softmax / Broyden update / state traffic of trip i runs under J(0), J(1) of trip i+1,
    // and the next gather + layer 0 run under the drain of trip i's last jobs.
    int id = gather(0);
    cta_sync_compute();
    pc.mark(0);
    layer0(0);
    store_x();                                                         // -> J(0) of the first trip
    if (ntrips > 1) { layer0(1); store_x(); }                          // -> J(1)
    pc.mark(1);
    for (int trip = 0; trip < ntrips; trip += 2) {
        const bool haveB = trip + 1 < ntrips;                          // B exists as a (possibly all-padding) tile of this CTA
        const bool more = trip + 2 < ntrips;                           // another trip follows
        // ---- interleaved MLPs: jobs J(k), k = 0 .. 4 ntile - 1, tile k % ntile, layer k / ntile (+1); J(0), J(1) are already fed
        const int ntile = haveB ? 2 : 1, njobs = 4 * ntile;
        for (int k = ntile; k < njobs; ++k) {
            const int t = k % ntile, s = k / ntile;
            wait_done(t);                                              // J(k - ntile) complete: D_t holds the pre-activations of layer s
            pc.mark(2);
            epilogue(t, s);                                            // under J(k - 1) (the other tile)
            sync_reads();                                              // J(k) overwrites D_t
            store_x();                                                 // as J(k - 1) releases the X chunks -> J(k)
            pc.mark(3);
        }
        int id_next = -1;
        if (more) {                                                    // under the last jobs of this trip
            id_next = gather(trip + 2);
            cta_sync_compute();
            layer0(0);
        }
        pc.mark(0);
        for (int t = 0; t < ntile; ++t) { wait_done(t); take_logits(t); }
        sync_reads();                                                  // logits visible to their owner threads; D_A / D_B read by everyone
        pc.mark(4);
        if (more) {
            store_x();                                                 // -> J(0) of the next trip
            if (trip + 3 < ntrips) { layer0(1); store_x(); }           // -> J(1)
            pc.mark(1);
        }
        {
            bool active = false;
            BroydenState<3> st;
            if (id >= 0) {
                float dx[3] = {0.f, 0.f, 0.f};
                state_load(st, &w.corr_state[id]);
                if (iter >= 0) broyden_advance<3>(st, dx);
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
                const bool fin = (id >= 0) && !active;
                warp_append(fin && st.best_n < CVG_THRESH, fin ? st.owner : 0, w.shade_list, &w.counters[C_SHADE]);
                warp_stat_add(fin ? st.g_evals : 0, &w.counters[C_STAT_CORR_EVALS]);
            }
        }
        cta_sync_compute();                                            // logits consumed before the next trip's take_logits
        id = id_next;
        pc.mark(5);
    }
    tc_fence_before();
    cta_sync_compute();
    if (warp == 0) tmem_dealloc(tbase, 512);
    cluster_sync_exit();
}

}  // namespace arah
