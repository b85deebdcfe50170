// arah_trace_p.cuh — k_trace_persist: sphere tracing (BodyRayTracing.sphere_tracing,
// /root/reference/im2mesh/metaavatar_render/renderer/ray_tracing.py:174-241) as ONE persistent kernel.
//
// Round 1 launched, for each of the 50 marching steps, one 1-NN kernel and one SDF kernel over the list of rays still marching
// (100 launches; from step ~10 on a launch held a handful of rays and cost its fixed 70-120 us).  Here a CTA (one per SM) keeps a
// resident tile of 128 rays: per step it finds each ray's nearest posed SMPL vertex (exact clustered 1-NN over the vertex index
// held in shared memory, arah_work.cuh), inverts the blended vertex transform (ray_tracing.py:382-400), evaluates the SDF of the
// 128 canonical points on the tensor cores (arah_sdf16.cuh) and advances the rays (:228-241).  A ray that has converged,
// diverged or used its 50 steps is written back and its row is re-filled from the device-wide list of rays at once, so the tile
// stays full while rays are left; every ray runs exactly the step sequence of the reference (rays are independent).
// Warps: 0-15 compute (row work + epilogues), 16 TMA producer (runs ahead speculatively, stops on the `stop` flag), 17 MMA issuer.
#pragma once
#include "arah_sdf16.cuh"

namespace arah {

// per-row words, SoA in shared memory: word f of row r at st[f * 128 + r]
enum { TR_RAY = 0, TR_T = 1, TR_FAR = 2, TR_D = 3, TR_IT = 6, TR_S = 7, TR_T12 = 8, TR_XN = 20, TR_SLOT = 23, TR_WORDS = 24 };

__host__ __device__ constexpr size_t trace_persist_smem_bytes(int n_verts) {
    return (size_t)S16_NSLOTS * S16_SLOT_BYTES + knn_smem_bytes(n_verts) + (size_t)(TR_WORDS * UM + 4 * UM + 8) * 4 + sizeof(S16Ctl) + 64;
}

__global__ void __launch_bounds__(S16_THREADS, 1) k_trace_persist(FrameParams fp, SdfF16 sd, KnnIndex ix, Work w) {
    extern __shared__ __align__(1024) uint8_t raw_smem[];
    const int n = w.counters[C_TRACE];                      // rays with near < far (k_trace_begin), listed in w.listA
    if (n <= 0) return;
    if (smem_u32(raw_smem) & 1023u) __trap();               // SWIZZLE_128B images need the segment 1024-aligned
    uint8_t* ring = raw_smem;
    float4* sknn = reinterpret_cast<float4*>(ring + S16_NSLOTS * S16_SLOT_BYTES);
    float* st = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sknn) + knn_smem_bytes(fp.n_verts));
    float (*part)[UM] = reinterpret_cast<float (*)[UM]>(st + TR_WORDS * UM);
    float* sInv = reinterpret_cast<float*>(part) + 4 * UM;                        // [5] (+3 pad)
    S16Ctl* ctl = reinterpret_cast<S16Ctl*>(sInv + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s16_ctl_init(ctl);
    if (warp == 17) tmem_alloc(&ctl->tslot, 512);
    if (tid < 5) sInv[tid] = __ldg(sd.scale + 2 * tid + 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = ctl->tslot;

    if (warp == 16) {                                       // ===== TMA producer =====
        if (lane == 0) {
            S16Prod p;
            while (s16_produce_sdf(ring, ctl, p, sd, true)) {}
            s16_drain(ctl, p);
        }
        return;
    }
    if (warp == 17) {                                       // ===== MMA issuer =====
        if (lane == 0) {
            S16Mma m;
            for (uint32_t e = 0;; ++e) {
                mbar_wait(&ctl->go, e & 1u);
                if (!ctl->cont[e & 1u]) break;
                s16_mma_sdf(ring, ctl, m, tbase);
            }
        }
        __syncwarp();
        s16_sync_exit();                                    // compute warps are out of tensor memory
        tmem_dealloc(tbase, 512);
        return;
    }
    // ===== compute warps =====
    // the 1-NN index (sorted vertices + cluster / super boxes) goes to shared memory once per CTA; only compute warps take part
    KnnSmem kk;
    {
        const int nv = ix.nc * KNN_CLUSTER, ns = (ix.nc + KNN_SUPER - 1) / KNN_SUPER;
        for (int v = tid; v < nv; v += S16_CTHREADS) sknn[v] = __ldg(ix.sv + v);
        for (int c = tid; c < ix.nc; c += S16_CTHREADS) { sknn[nv + c] = __ldg(ix.cmin + c); sknn[nv + ix.nc + c] = __ldg(ix.cmax + c); }
        s16_sync();
        for (int g = tid; g < ns; g += S16_CTHREADS) {
            float4 mn = make_float4(1e30f, 1e30f, 1e30f, 0.f), mx = make_float4(-1e30f, -1e30f, -1e30f, 0.f);
            for (int c = g * KNN_SUPER; c < min(ix.nc, (g + 1) * KNN_SUPER); ++c) {
                const float4 a = sknn[nv + c], b = sknn[nv + ix.nc + c];
                mn.x = fminf(mn.x, a.x); mn.y = fminf(mn.y, a.y); mn.z = fminf(mn.z, a.z);
                mx.x = fmaxf(mx.x, b.x); mx.y = fmaxf(mx.y, b.y); mx.z = fmaxf(mx.z, b.z);
            }
            sknn[nv + 2 * ix.nc + g] = mn; sknn[nv + 2 * ix.nc + ns + g] = mx;
        }
        kk.sv = sknn; kk.cmin = sknn + nv; kk.cmax = sknn + nv + ix.nc; kk.smin = sknn + nv + 2 * ix.nc; kk.smax = kk.smin + ns; kk.nc = ix.nc; kk.ns = ns;
    }
    const int q = warp & 3, u = warp >> 2, r = 32 * q + lane;           // TMEM row of this thread / its 16 columns per chunk (SDF epilogues)
    uint32_t done_par = 0;
    int evals = 0;
    // row `tid` (tid < 128) asks for a ray: warp-aggregated claim on the list cursor
    auto refill = [&](bool need) {
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (!m) return;
        int base = 0;
        if (lane == (__ffs(m) - 1)) base = atomicAdd(&w.counters[C_TRACE_CURSOR], __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (need) {
            const int i = base + __popc(m & ((1u << lane) - 1u));
            int ray = -1;
            if (i < n) {
                ray = w.listA[i];
                st[TR_T * UM + tid] = w.near_far[2 * ray];
                st[TR_FAR * UM + tid] = w.near_far[2 * ray + 1];
#pragma unroll
                for (int k = 0; k < 3; ++k) st[(TR_D + k) * UM + tid] = w.ray_dirs[3 * ray + k];
                st[TR_IT * UM + tid] = __int_as_float(0);
                st[TR_SLOT * UM + tid] = __int_as_float(-1);               // no previous nearest vertex yet
            }
            st[TR_RAY * UM + tid] = __int_as_float(ray);
        }
    };
    if (tid < UM) refill(true);
    bool live = s16_sync_or(tid < UM && __float_as_int(st[TR_RAY * UM + tid]) >= 0);
    if (tid == 0) { ctl->cont[0] = live ? 1 : 0; if (!live) ctl->stop = 1; __threadfence_block(); mbar_arrive(&ctl->go); }
    uint32_t e = 0;
    PhaseClk pc; pc.start((tid == 32 && w.phase_clk) ? w.phase_clk + 16 : nullptr);      // [0] 1-NN, [1] layer 0, [2] MMA wait, [3] epilogues, [4] marching
    while (live) {
        // ---- nearest posed vertex + inverse NN skinning.  Two shapes (w.trace_knn): 1 = one row per lane on warps 0-3, the exact
        // scan seeded with the row's previous winner (a marching ray moves little between steps, so the first bound is tight);
        // 0 = four rows at a time per warp, 8 lanes each (octet form), all 16 warps
        if (w.trace_knn) {
            if (tid < UM) {
                const int ray = __float_as_int(st[TR_RAY * UM + tid]);
                float xn[3] = {0.f, 0.f, 0.f};
                if (ray >= 0) {
                    const float t = st[TR_T * UM + tid];
                    float x[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) x[k] = st[(TR_D + k) * UM + tid] * t + fp.cam_loc[k];
                    int slot = __float_as_int(st[TR_SLOT * UM + tid]);
                    const int idx = knn_scan_seeded(kk, x[0], x[1], x[2], slot);
                    st[TR_SLOT * UM + tid] = __int_as_float(slot);
                    float T12[12], s_, xh[3];
                    nn_inverse_skinning(fp, idx, x, T12, &s_, xh);
                    normalize3(fp, xh, xn);
#pragma unroll
                    for (int k = 0; k < 12; ++k) st[(TR_T12 + k) * UM + tid] = T12[k];
                    st[TR_S * UM + tid] = s_;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) st[(TR_XN + k) * UM + tid] = xn[k];
            }
        } else {
            const int row = 8 * warp + (lane & 7);
            const bool mine_row = lane < 8;
            const int ray = __float_as_int(st[TR_RAY * UM + row]);
            float x[3] = {0.f, 0.f, 0.f};
            if (ray >= 0) {
                const float t = st[TR_T * UM + row];
#pragma unroll
                for (int k = 0; k < 3; ++k) x[k] = st[(TR_D + k) * UM + row] * t + fp.cam_loc[k];
            }
            // TR_SLOT holds the cluster of the row's previous nearest vertex (-1 on the first step): the seed of this step's scan
            const int seed = (ray >= 0) ? __float_as_int(st[TR_SLOT * UM + row]) : -1;
            int mine = 0, mine_c = -1;
#pragma unroll 1
            for (int r4 = 0; r4 < 8; r4 += 4) {                          // four queries at a time, one per octet (knn_warp_batches)
                const int qi = r4 + (lane >> 3);
                const float qx = __shfl_sync(0xffffffffu, x[0], qi), qy = __shfl_sync(0xffffffffu, x[1], qi), qz = __shfl_sync(0xffffffffu, x[2], qi);
                const bool qv = __shfl_sync(0xffffffffu, (int)(ray >= 0), qi) != 0;
                const int qs = __shfl_sync(0xffffffffu, seed, qi);
                int wc = -1;
                const int idx = knn_scan_octet(kk, qx, qy, qz, qv, qv ? qs : 0, &wc);
                const int got = __shfl_sync(0xffffffffu, idx, (lane & 3) * 8), got_c = __shfl_sync(0xffffffffu, wc, (lane & 3) * 8);
                if ((lane >> 2) == (r4 >> 2)) { mine = got; mine_c = got_c; }
            }
            if (mine_row && ray >= 0) st[TR_SLOT * UM + row] = __int_as_float(mine_c);
            float xn[3] = {0.f, 0.f, 0.f};
            if (mine_row && ray >= 0) {
                float T12[12], s_, xh[3];
                nn_inverse_skinning(fp, mine, x, T12, &s_, xh);
                normalize3(fp, xh, xn);
#pragma unroll
                for (int k = 0; k < 12; ++k) st[(TR_T12 + k) * UM + row] = T12[k];
                st[TR_S * UM + row] = s_;
            }
            if (mine_row) {
#pragma unroll
                for (int k = 0; k < 3; ++k) st[(TR_XN + k) * UM + row] = xn[k];
            }
        }
        s16_sync();
        pc.mark(0);
        // ---- SDF of the 128 canonical points
        const float dot = s16_compute_sdf(sd, st[TR_XN * UM + r], st[(TR_XN + 1) * UM + r], st[(TR_XN + 2) * UM + r], ctl, done_par, tbase, sInv, &pc);
        part[u][r] = dot;
        s16_sync();
        // ---- marching logic (ray_tracing.py:228-241), one thread per row
        bool row_live = false;
        if (tid < UM) {
            const int ray = __float_as_int(st[TR_RAY * UM + tid]);
            bool need = false;
            if (ray >= 0) {
                const float sdf = sdf_to_metres(((part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid])) + __ldg(sd.b6), fp.cmin, fp.cmax);
                float t = st[TR_T * UM + tid];
                const float far_ = st[TR_FAR * UM + tid];
                const float sm = fminf(fmaxf(sdf, -0.1f), 0.1f);
                bool diverge = false;
                if (fabsf(sm) > CVG_THRESH && fabsf(sdf) < 1e6f) { t = t + sm; diverge = t >= far_; st[TR_T * UM + tid] = t; }
                const bool still = !(fabsf(sdf) <= CVG_THRESH || diverge);
                const int it = __float_as_int(st[TR_IT * UM + tid]) + 1;
                ++evals;
                if (!still || it >= TRACE_ITERS) {                       // this ray is done: last depth, flags and last evaluated point
                    w.ray_t[ray] = t;
                    w.ray_flags[ray] = (still ? 1 : 0) | (diverge ? 2 : 0);
                    RayCur c;
#pragma unroll
                    for (int k = 0; k < 3; ++k) c.xn[k] = st[(TR_XN + k) * UM + tid];
                    c.s = st[TR_S * UM + tid];
#pragma unroll
                    for (int k = 0; k < 12; ++k) c.T[k] = st[(TR_T12 + k) * UM + tid];
                    w.ray_cur[ray] = c;
                    need = true;
                } else st[TR_IT * UM + tid] = __int_as_float(it);
            }
            refill(need);
            row_live = __float_as_int(st[TR_RAY * UM + tid]) >= 0;
        }
        live = s16_sync_or(row_live);
        pc.mark(4);
        if (pc.dst) atomicAdd(pc.dst + 5, 1ull);                         // evaluations of this CTA
        ++e;
        if (tid == 0) { ctl->cont[e & 1u] = live ? 1 : 0; if (!live) ctl->stop = 1; __threadfence_block(); mbar_arrive(&ctl->go); }
    }
    warp_stat_add(evals, &w.counters[C_STAT_TRACE_EVALS]);
    tc_fence_before();
    s16_sync_exit();
}

}  // namespace arah
