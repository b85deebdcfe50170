// arah_iso_p.cuh — k_iso_persist: the joint iso-surface / correspondence search (search_iso_surface_depth,
// /root/reference/im2mesh/utils/root_finding_utils.py:426-461, driven by broyden.py:47-76 with D = 4) as ONE persistent kernel.
//
// Round 1 launched one kernel per Broyden step (50 launches of ~120 us although ~2 steps per ray suffice on average: 45 of them
// found almost nothing to do).  Here a CTA keeps a resident tile of 128 rays; every ray iterates on its own until its residual is
// below 1e-5, diverges or has used 50 steps, and a finished row is re-filled from the list of rays at once.  Per step a row
// evaluates g(u) = [ sdf(x_hat) ; LBS(x_hat) - (o + z d - trans) ]: the skinning MLP and the SDF both on tcgen05 in fp16 split
// precision (arah_f16x3.cuh, arah_sdf16.cuh); residual, rank-1 Jacobian update and bookkeeping are the per-point restatement of
// broyden.py used everywhere else (broyden_update<4>).  The Broyden state (208 B per ray) lives in shared memory while a ray is
// resident; it is read from / written to w.iso_state once per ray (k_iso_init_tc3 before, k_trace_finish after).
#pragma once
#include "arah_corr_p.cuh"
#include "arah_sdf16.cuh"

namespace arah {

constexpr int IP_STATE_WORDS = 52;          // BroydenState<4> as 32-bit words
enum { IS_X = 0, IS_J = 4, IS_GX = 20, IS_UPD = 24, IS_BX = 28, IS_BT = 32, IS_BN = 44, IS_OWNER = 45, IS_TG = 46, IS_EV = 49,
       IP_RAY = 52, IP_IT = 53, IP_DX = 54, IP_XN = 58, IP_WORDS = 61 };
static_assert(sizeof(BroydenState<4>) == IP_STATE_WORDS * 4, "state words");

__host__ __device__ constexpr size_t iso_persist_smem_bytes() {
    return (size_t)S16_NSLOTS * S16_SLOT_BYTES + (size_t)(IP_WORDS * UM + 25 * UM + 4 * UM + 3 * 128 + 5 * 128 + 16) * 4 + sizeof(S16Ctl) + 64;
}

__global__ void __launch_bounds__(S16_THREADS, 1) k_iso_persist(FrameParams fp, SdfF16 sd, SkinF16 sk, Work w) {
    extern __shared__ __align__(1024) uint8_t raw_smem[];
    const int n = w.counters[C_ISO];
    if (n <= 0) return;
    if (smem_u32(raw_smem) & 1023u) __trap();
    uint8_t* ring = raw_smem;
    float* st = reinterpret_cast<float*>(ring + S16_NSLOTS * S16_SLOT_BYTES);      // [IP_WORDS][128]
    float* lgs = st + IP_WORDS * UM;                                               // [25][128] logits (incl. bias)
    float (*part)[UM] = reinterpret_cast<float (*)[UM]>(lgs + 25 * UM);
    float* sW0 = reinterpret_cast<float*>(part) + 4 * UM;                          // skinning layer 0 [3][128]
    float* sb = sW0 + 3 * 128;                                                     // skinning biases 4 x 128, then 32
    float* sInv = sb + 5 * 128;                                                    // [0..3] skinning layers 1..4, [4..8] SDF layers 1..5
    S16Ctl* ctl = reinterpret_cast<S16Ctl*>(sInv + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) s16_ctl_init(ctl);
    if (warp == 17) tmem_alloc(&ctl->tslot, 512);
    for (int i = tid; i < 3 * 128; i += S16_THREADS) sW0[i] = __ldg(sk.Wt0 + i);
    for (int i = tid; i < 4 * 128; i += S16_THREADS) sb[i] = __ldg(sk.b[i >> 7] + (i & 127));
    if (tid < 32) sb[512 + tid] = __ldg(sk.b[4] + tid);
    if (tid < 4) sInv[tid] = __ldg(sk.scale + 2 * tid + 1);
    if (tid < 5) sInv[4 + tid] = __ldg(sd.scale + 2 * tid + 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = ctl->tslot;

    if (warp == 16) {                                       // ===== TMA producer: skinning images then SDF images, per evaluation =====
        if (lane == 0) {
            S16Prod p;
            bool ok = true;
            while (ok) {
                for (int s = 0; s < 4 && ok; ++s) {
                    const uint32_t bytes = s < 3 ? 32768u : 8192u;
                    ok = s16_put(ring, ctl, p, reinterpret_cast<const char*>(sk.hi) + (size_t)s * 32768, bytes, true);
                    if (ok) ok = s16_put(ring, ctl, p, reinterpret_cast<const char*>(sk.lo) + (size_t)s * 32768, bytes, true);
                }
                if (ok) ok = s16_produce_sdf(ring, ctl, p, sd, true);
            }
            s16_drain(ctl, p);
        }
        return;
    }
    if (warp == 17) {                                       // ===== MMA issuer =====
        if (lane == 0) {
            S16Mma m;
            uint32_t skpar = 0;
            for (uint32_t e = 0;; ++e) {
                mbar_wait(&ctl->go, e & 1u);
                if (!ctl->cont[e & 1u]) break;
                for (int s = 0; s < 4; ++s) {               // skinning MLP: X_hi [0,64) | X_lo [64,128) | D [128,256)
                    const int N = (s < 3) ? 128 : 32;
                    const uint32_t idesc = umma_idesc_f16(UM, N), img = (uint32_t)N * HK * 2;
                    mbar_wait(&ctl->ready_sk, skpar);
                    skpar ^= 1u;
                    mbar_wait(&ctl->full[m.slot], m.use & 1u);                 // hi image (both K-chunks)
                    tc_fence_after();
                    uint32_t b = smem_u32(ring + m.slot * S16_SLOT_BYTES);
#pragma unroll
                    for (int kc = 0; kc < 2; ++kc)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t col = 32u * kc + 8u * k;
                            umma_f16_ts(tbase + 128u, tbase + 64u + col, umma_smem_desc_sw128(b + kc * img + 32u * k), idesc, (kc > 0 || k > 0) ? 1u : 0u);
                            umma_f16_ts(tbase + 128u, tbase + col, umma_smem_desc_sw128(b + kc * img + 32u * k), idesc, 1u);
                        }
                    umma_commit(&ctl->empty[m.slot]);
                    if (++m.slot == S16_NSLOTS) { m.slot = 0; ++m.use; }
                    mbar_wait(&ctl->full[m.slot], m.use & 1u);                 // lo image
                    tc_fence_after();
                    b = smem_u32(ring + m.slot * S16_SLOT_BYTES);
#pragma unroll
                    for (int kc = 0; kc < 2; ++kc)
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ts(tbase + 128u, tbase + 32u * kc + 8u * k, umma_smem_desc_sw128(b + kc * img + 32u * k), idesc, 1u);
                    umma_commit(&ctl->empty[m.slot]);
                    if (++m.slot == S16_NSLOTS) { m.slot = 0; ++m.use; }
                    umma_commit(&ctl->done);
                }
                s16_mma_sdf(ring, ctl, m, tbase);
            }
        }
        __syncwarp();
        s16_sync_exit();
        tmem_dealloc(tbase, 512);
        return;
    }
    // ===== compute warps =====
    const int q = warp & 3, u = warp >> 2, r = 32 * q + lane;        // TMEM lane quarter, column quarter, row
    const uint32_t trow = tbase + ((uint32_t)(32 * q) << 16);
    uint32_t done_par = 0;
    int evals = 0;
    auto wait_done = [&]() { mbar_wait(&ctl->done, done_par); done_par ^= 1u; __syncwarp(); tc_fence_after(); };
    auto publish_sk = [&]() { tmem_st_wait(); tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&ctl->ready_sk); };
    auto emit = [&](float (&v)[32], int xcol) {
        uint32_t hi[16], lo[16];
        split_pack_f16(v, hi, lo);
        tmem_st16(trow + (uint32_t)xcol, hi);
        tmem_st16(trow + 64u + (uint32_t)xcol, lo);
    };
    // row `tid` takes the next ray of the list: state from w.iso_state (k_iso_init_tc3), first step applied
    auto refill = [&](bool need) {
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (!m) return;
        int base = 0;
        if (lane == (__ffs(m) - 1)) base = atomicAdd(&w.counters[C_ISO_CURSOR], __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (need) {
            const int i = base + __popc(m & ((1u << lane) - 1u));
            int ray = -1;
            if (i < n) {
                ray = w.listA[i];
                const uint32_t* sp = reinterpret_cast<const uint32_t*>(&w.iso_state[ray]);
#pragma unroll 4
                for (int k = 0; k < IP_STATE_WORDS; ++k) st[k * UM + tid] = __uint_as_float(sp[k]);
                st[IP_IT * UM + tid] = __int_as_float(0);
            }
            st[IP_RAY * UM + tid] = __int_as_float(ray);
        }
    };
    // x += update (broyden.py:50-51); keeps the applied step and publishes the normalised query point
    auto advance = [&]() {
        float xq[3], xn[3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float u = st[(IS_UPD + k) * UM + tid];
            const float x = st[(IS_X + k) * UM + tid] + u;
            st[(IS_X + k) * UM + tid] = x;
            st[(IP_DX + k) * UM + tid] = u;
            if (k < 3) xq[k] = x;
        }
        normalize3(fp, xq, xn);
#pragma unroll
        for (int k = 0; k < 3; ++k) st[(IP_XN + k) * UM + tid] = xn[k];
    };
    if (tid < UM) {
        refill(true);
        if (__float_as_int(st[IP_RAY * UM + tid]) >= 0) advance();
        else { st[(IP_XN) * UM + tid] = 0.f; st[(IP_XN + 1) * UM + tid] = 0.f; st[(IP_XN + 2) * UM + tid] = 0.f; }
    }
    bool live = s16_sync_or(tid < UM && __float_as_int(st[IP_RAY * UM + tid]) >= 0);
    if (tid == 0) { ctl->cont[0] = live ? 1 : 0; if (!live) ctl->stop = 1; __threadfence_block(); mbar_arrive(&ctl->go); }
    uint32_t e = 0;
    while (live) {
        const float x = st[IP_XN * UM + r], y = st[(IP_XN + 1) * UM + r], z = st[(IP_XN + 2) * UM + r];
        // ---- skinning MLP (layer 0 on the FP32 pipe, layers 1..4 on the tensor cores)
        {
            const int col0 = 32 * u;
            float v[32];
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {                            // 128-bit shared-memory loads (LDS shares the MIO queue with MUFU)
                const int cc = col0 + 4 * i4;
                const float4 wx = *reinterpret_cast<const float4*>(sW0 + cc), wy = *reinterpret_cast<const float4*>(sW0 + 128 + cc),
                             wz = *reinterpret_cast<const float4*>(sW0 + 256 + cc), bb = *reinterpret_cast<const float4*>(sb + cc);
                v[4 * i4 + 0] = softplus100_fast(fmaf(wz.x, z, fmaf(wy.x, y, wx.x * x)) + bb.x);
                v[4 * i4 + 1] = softplus100_fast(fmaf(wz.y, z, fmaf(wy.y, y, wx.y * x)) + bb.y);
                v[4 * i4 + 2] = softplus100_fast(fmaf(wz.z, z, fmaf(wy.z, y, wx.z * x)) + bb.z);
                v[4 * i4 + 3] = softplus100_fast(fmaf(wz.w, z, fmaf(wy.w, y, wx.w * x)) + bb.w);
            }
            emit(v, col0 / 2);
        }
        publish_sk();
#pragma unroll 1
        for (int l = 1; l < 4; ++l) {
            wait_done();
            const float inv = sInv[l - 1];
            {
                const int col0 = 32 * u;
                float v[32];
                tmem_ld32(trow + 128u + (uint32_t)col0, v);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 bb = *reinterpret_cast<const float4*>(sb + 128 * l + col0 + 4 * i4);
                    v[4 * i4 + 0] = softplus100_fast(fmaf(v[4 * i4 + 0], inv, bb.x));
                    v[4 * i4 + 1] = softplus100_fast(fmaf(v[4 * i4 + 1], inv, bb.y));
                    v[4 * i4 + 2] = softplus100_fast(fmaf(v[4 * i4 + 2], inv, bb.z));
                    v[4 * i4 + 3] = softplus100_fast(fmaf(v[4 * i4 + 3], inv, bb.w));
                }
                emit(v, col0 / 2);
            }
            publish_sk();
        }
        wait_done();
        if (u == 0) {                                                   // logits of row r (= tid for the row threads)
            float v[32];
            tmem_ld32(trow + 128u, v);
            const float inv4 = sInv[3];
#pragma unroll
            for (int k = 0; k < 25; ++k) lgs[k * UM + r] = fmaf(v[k], inv4, sb[512 + k]);
        }
        tc_fence_before();
        s16_sync();                                                     // tensor memory is free for the SDF
        tc_fence_after();
        // ---- SDF
        const float dot = s16_compute_sdf(sd, x, y, z, ctl, done_par, tbase, sInv + 4);
        part[u][r] = dot;
        s16_sync();
        // ---- residual + Broyden update, one thread per row
        bool row_live = false;
        if (tid < UM) {
            const int ray = __float_as_int(st[IP_RAY * UM + tid]);
            bool need = false;
            if (ray >= 0) {
                BroydenState<4> s;
                float dx[4], g[4], T12[12], lg32[32];
                uint32_t* sw = reinterpret_cast<uint32_t*>(&s);
#pragma unroll
                for (int k = 0; k < IP_STATE_WORDS; ++k) sw[k] = __float_as_uint(st[k * UM + tid]);
#pragma unroll
                for (int k = 0; k < 4; ++k) dx[k] = st[(IP_DX + k) * UM + tid];
#pragma unroll
                for (int k = 0; k < 25; ++k) lg32[k] = lgs[k * UM + tid];
                iso_residual(fp, w, ray, s.x, lg32, ((part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid])) + __ldg(sd.b6), g, T12);
                bool active = broyden_update<4>(s, dx, g, T12);
                const int it = __float_as_int(st[IP_IT * UM + tid]);
                if (it + 1 >= BROYDEN_ITERS) active = false;
                ++evals;
#pragma unroll
                for (int k = 0; k < IP_STATE_WORDS; ++k) st[k * UM + tid] = __uint_as_float(sw[k]);
                if (active) { st[IP_IT * UM + tid] = __int_as_float(it + 1); }
                else { state_store(&w.iso_state[ray], s); need = true; }
            }
            refill(need);
            row_live = __float_as_int(st[IP_RAY * UM + tid]) >= 0;
            if (row_live) advance();
        }
        live = s16_sync_or(row_live);
        ++e;
        if (tid == 0) { ctl->cont[e & 1u] = live ? 1 : 0; if (!live) ctl->stop = 1; __threadfence_block(); mbar_arrive(&ctl->go); }
    }
    warp_stat_add(evals, &w.counters[C_STAT_ISO_EVALS]);
    tc_fence_before();
    s16_sync_exit();
}

}  // namespace arah
