// arah_work.cuh — what the stage kernels share: the per-frame workspace (Work), device counters, the Broyden-state copies,
// warp-aggregated list appends, the exact clustered 1-NN over the posed SMPL vertices and the per-point helpers.
// Only inline device code and plain structs live here, so that several translation units can include it (the __global__ kernels
// of arah_kernels.cuh belong to arah_api.cu; the persistent root-finding kernels to arah_root.cu).
#pragma once
#include "arah_tile.cuh"

namespace arah {

constexpr int TRACE_ITERS = 50;
constexpr int BROYDEN_ITERS = 50;
constexpr int MAX_STEPS = 256;

// counter slots (int32) in Work::counters
enum Ctr {
    C_TRACE = 0,                         // [0..50] active rays entering sphere-tracing step i
    C_ISO = C_TRACE + TRACE_ITERS + 1,   // [0..50] active rays entering joint-search step i
    C_CORR = C_ISO + BROYDEN_ITERS + 1,  // [0..50] active samples entering correspondence step i
    C_ON = C_CORR + BROYDEN_ITERS + 1,   // number of "on" samples
    C_SHADE,                             // number of converged samples to shade
    C_SHADE2,                            // ... of which alpha != 0 (exact cull): samples that need gradient + colour
    C_STAT_TRACE_EVALS, C_STAT_ISO_EVALS, C_STAT_CORR_EVALS, C_STAT_HIT_RAYS, C_STAT_VOL_RAYS,
    C_CORR_CURSOR,                       // persistent correspondence kernel: next unclaimed block of on-samples
    C_TRACE_CURSOR, C_ISO_CURSOR,        // persistent tracing / joint-search kernels: next unclaimed ray / list entry
    C_COUNT
};

struct RayCur { float xn[3]; float s; float T[12]; };   // 64 B: last sphere-tracing evaluation of a ray
// what the persistent correspondence kernel needs to start a sample (written by k_knn_samples): the kNN-skinning start point
// x0 (metres), the kNN transform (returned as the best transform if no iterate beats the start, broyden.py:38-40), the posed
// target x - trans and the sample slot
struct alignas(16) CorrSeed { float x[3]; int32_t owner; float T[12]; float tgt[3]; float pad; };   // 80 B

struct Work {
    int P, S;
    const float* ray_dirs;    // [P][3]
    const float* near_far;    // [P][2]
    float* ray_t;             // [P]
    uint8_t* ray_flags;       // [P] bit0 unfinished, bit1 diverged
    RayCur* ray_cur;          // [P]
    BroydenState<4>* iso_state;   // [P]
    uint8_t* ray_conv;        // [P]   BodyRayTracing network_body_mask
    float* ray_dist;          // [P]   dists
    float* ray_pnorm;         // [P][3] points_hat_norm
    float* z_vals;            // [P][S]
    float* smp_xn;            // [P*S][3]
    float* smp_T;             // [P*S][12]
    uint8_t* smp_conv;        // [P*S]
    float* smp_sdf;           // [P*S]   metres
    float* smp_rgb;           // [P*S][3]
    BroydenState<3>* corr_state;  // [P*S]   per-iteration launches (fp32 root mode)
    CorrSeed* corr_seed;          // [P*S]   persistent kernel (aliases corr_state's storage); null selects corr_state
    int* listA; int* listB;   // [P*S]
    int* on_list;             // [P*S] sample slot index of the k-th on-sample
    int* ray_on_base;         // [P]   position of the ray's first on-sample in on_list (its on-samples are consecutive)
    int* shade_list;          // [P*S]
    int* counters;            // [C_COUNT]
    int knn_seed;             // k_knn_samples: 2 = one ray per lane, every sample seeded by the previous one (default); 1 = runs of 4
                              // on-samples per lane; 0 = unseeded batches (all exact; switchable for A/B)
    int trace_knn;            // k_trace_persist: 1 = seeded one-row-per-lane 1-NN, 0 = octet form (exact both; switchable for A/B)
    int shade_ctr;            // counter slot holding the length of shade_list for the tensor-core shading kernel (C_SHADE / C_SHADE2)
    int shade_keep_sdf;       // full shading pass leaves smp_sdf alone (k_sdf_fwd16 has written the value compositing uses)
    float* scratch;           // shade kernel: per-CTA [7][TM][256]
    float* out_rgb;           // [P][3]
    uint8_t* out_mask;        // [P]
    float* out_points_cam;    // [P][3]
    float* out_wsum;          // [P]
    unsigned long long* phase_clk;   // [16] debug: SM-clock cycles per kernel phase, accumulated by one thread per CTA (may be null)
    // training-mode tracing (BodyRayTracing.forward(eval_mode=False), ray_tracing.py:249,298-311): all rays enter the joint
    // search and the z samples are jittered with the caller's three torch.rand draws
    int train;                // 0 = eval
    const float* u_all;       // [P][S]
    const float* u_near;      // [P][near+1]
    const float* u_far;       // [P][far]
};

// phase timer used by one designated thread per CTA: adds the cycles since the previous mark to slot `i`
struct PhaseClk {
    unsigned long long* dst; long long t;
    __device__ __forceinline__ void start(unsigned long long* d) { dst = d; t = clock64(); }
    __device__ __forceinline__ void mark(int i) { if (dst) { const long long n = clock64(); atomicAdd(dst + i, (unsigned long long)(n - t)); t = n; } }
};

// 128-bit copies of a state record (the struct is alignas(16) and a multiple of 16 bytes)
template <class T> __device__ __forceinline__ void state_load(T& dst, const T* src) {
    static_assert(sizeof(T) % 16 == 0, "state records are 16-byte multiples");
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(&dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = s[i];
}
template <class T> __device__ __forceinline__ void state_store(T* dst, const T& src) {
    const uint4* s = reinterpret_cast<const uint4*>(&src);
    uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 16); ++i) d[i] = s[i];
}

// warp-aggregated append of `value` to list (returns nothing); all 32 lanes must call
__device__ __forceinline__ void warp_append(bool pred, int value, int* list, int* counter) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == (__ffs(m) - 1)) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}
__device__ __forceinline__ void warp_stat_add(int v, int* counter) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counter, v);
}


// ---- exact 1-NN over the posed SMPL vertices ---------------------------------------------------------------------
// k_knn_build (once per frame) sorts the vertices along a 30-bit Morton curve and groups them into clusters of 32 with an
// AABB each.  A query first finds the cluster with the smallest box distance, scans it, then scans only clusters whose box
// could still hold a closer vertex: typically 2-6 of 216 clusters instead of all 6890 vertices.  The result is the exact
// argmin of (x-v).(x-v) with the lowest original index on ties == the brute-force answer
// (pytorch3d.ops.knn_points K=1, ray_tracing.py:386,407).
constexpr int KNN_CLUSTER = 32;
struct KnnIndex {
    const float4* sv;      // [nc*32] sorted vertices (x, y, z, original index as int bits); padding = +1e30
    const float4* cmin;    // [nc]
    const float4* cmax;    // [nc]
    int nc;
};
constexpr int KNN_SUPER = 8;          // cluster boxes per super box (Morton-contiguous, so spatially compact)
__host__ __device__ constexpr size_t knn_smem_bytes(int n_verts) {
    return (size_t)((n_verts + KNN_CLUSTER - 1) / KNN_CLUSTER) * (KNN_CLUSTER * 16 + 32)
         + (size_t)(((n_verts + KNN_CLUSTER - 1) / KNN_CLUSTER + KNN_SUPER - 1) / KNN_SUPER) * 32;
}
__device__ __forceinline__ uint32_t morton_spread10(uint32_t v) {
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

struct KnnSmem {
    const float4* sv; const float4* cmin; const float4* cmax; const float4* smin; const float4* smax; int nc, ns;
    const float4* qmin = nullptr; const float4* qmax = nullptr;      // optional: boxes of the four 8-vertex quarters of every cluster (load_knn_quarters)
};
constexpr int KNN_QUARTER = 8;
__host__ __device__ constexpr size_t knn_quarter_smem_bytes(int n_verts) { return (size_t)((n_verts + KNN_CLUSTER - 1) / KNN_CLUSTER) * 4 * 32; }
__device__ __forceinline__ KnnSmem load_knn(float4* smem, const KnnIndex& ix) {
    const int nv = ix.nc * KNN_CLUSTER, ns = (ix.nc + KNN_SUPER - 1) / KNN_SUPER;
    for (int v = threadIdx.x; v < nv; v += blockDim.x) smem[v] = __ldg(ix.sv + v);
    for (int c = threadIdx.x; c < ix.nc; c += blockDim.x) { smem[nv + c] = __ldg(ix.cmin + c); smem[nv + ix.nc + c] = __ldg(ix.cmax + c); }
    __syncthreads();
    // super boxes: the AABB of 8 consecutive cluster boxes (a box-distance test on it bounds all 8 from below)
    for (int g = threadIdx.x; g < ns; g += blockDim.x) {
        float4 mn = make_float4(1e30f, 1e30f, 1e30f, 0.f), mx = make_float4(-1e30f, -1e30f, -1e30f, 0.f);
        for (int c = g * KNN_SUPER; c < min(ix.nc, (g + 1) * KNN_SUPER); ++c) {
            const float4 a = smem[nv + c], b = smem[nv + ix.nc + c];
            mn.x = fminf(mn.x, a.x); mn.y = fminf(mn.y, a.y); mn.z = fminf(mn.z, a.z);
            mx.x = fmaxf(mx.x, b.x); mx.y = fmaxf(mx.y, b.y); mx.z = fmaxf(mx.z, b.z);
        }
        smem[nv + 2 * ix.nc + g] = mn; smem[nv + 2 * ix.nc + ns + g] = mx;
    }
    __syncthreads();
    KnnSmem k; k.sv = smem; k.cmin = smem + nv; k.cmax = smem + nv + ix.nc; k.smin = smem + nv + 2 * ix.nc; k.smax = k.smin + ns; k.nc = ix.nc; k.ns = ns;
    return k;
}
// Quarter boxes: the vertices of a cluster are Morton-sorted, so 8 consecutive ones are compact.  A cluster whose box passes the
// bound test usually only grazes the search ball: testing its four quarter boxes first (4 tests) replaces the scan of 32 vertices
// by that of the 8 or 16 that can matter.  `q` points at 8 * nc float4 of shared memory behind the index; call after load_knn.
__device__ __forceinline__ void load_knn_quarters(float4* q, KnnSmem& k) {
    for (int b = threadIdx.x; b < 4 * k.nc; b += blockDim.x) {
        float4 mn = make_float4(1e30f, 1e30f, 1e30f, 0.f), mx = make_float4(-1e30f, -1e30f, -1e30f, 0.f);
        for (int v = b * KNN_QUARTER; v < (b + 1) * KNN_QUARTER; ++v) {
            const float4 p = k.sv[v];
            mn.x = fminf(mn.x, p.x); mn.y = fminf(mn.y, p.y); mn.z = fminf(mn.z, p.z);
            mx.x = fmaxf(mx.x, p.x); mx.y = fmaxf(mx.y, p.y); mx.z = fmaxf(mx.z, p.z);
        }
        q[b] = mn; q[4 * k.nc + b] = mx;
    }
    __syncthreads();
    k.qmin = q; k.qmax = q + 4 * k.nc;
}
__device__ __forceinline__ float box_dist2(const float4 mn, const float4 mx, float x, float y, float z) {
    const float dx = fmaxf(fmaxf(mn.x - x, x - mx.x), 0.f), dy = fmaxf(fmaxf(mn.y - y, y - mx.y), 0.f), dz = fmaxf(fmaxf(mn.z - z, z - mx.z), 0.f);
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}
__device__ __forceinline__ void knn_scan_cluster(const float4* sv, int c, float x, float y, float z, float& bd, int& bi) {
#pragma unroll 8
    for (int v = c * KNN_CLUSTER; v < (c + 1) * KNN_CLUSTER; ++v) {
        const float4 p = sv[v];
        const float dx = x - p.x, dy = y - p.y, dz = z - p.z;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const int id = __float_as_int(p.w);
        if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; }
    }
}
// One query per lane.  Two-level pruning: 27 super boxes are tested instead of 216 cluster boxes; clusters are visited in index
// order exactly as a flat scan would (a cluster that passes `lb <= best` lies in a super box that passes it too: the super
// AABB contains the cluster AABB, and the fp32 box distance is monotone under containment), so the result is unchanged.
__device__ __forceinline__ int knn_scan(const KnnSmem& k, float x, float y, float z) {
    float lb0 = INFINITY;
    int s0 = 0;
    for (int g = 0; g < k.ns; ++g) {
        const float lb = box_dist2(k.smin[g], k.smax[g], x, y, z);
        if (lb < lb0) { lb0 = lb; s0 = g; }
    }
    lb0 = INFINITY;
    int c0 = s0 * KNN_SUPER;
    for (int c = s0 * KNN_SUPER; c < min(k.nc, (s0 + 1) * KNN_SUPER); ++c) {
        const float lb = box_dist2(k.cmin[c], k.cmax[c], x, y, z);
        if (lb < lb0) { lb0 = lb; c0 = c; }
    }
    float bd = INFINITY;
    int bi = 0x7fffffff;
    knn_scan_cluster(k.sv, c0, x, y, z, bd, bi);
    for (int g = 0; g < k.ns; ++g) {
        // the box distance is a lower bound computed in the same fp32 form; keep a 1-ulp-safe margin
        if (box_dist2(k.smin[g], k.smax[g], x, y, z) > bd * 1.000001f) continue;
        for (int c = g * KNN_SUPER; c < min(k.nc, (g + 1) * KNN_SUPER); ++c) {
            if (c == c0) continue;
            if (box_dist2(k.cmin[c], k.cmax[c], x, y, z) <= bd * 1.000001f) knn_scan_cluster(k.sv, c, x, y, z, bd, bi);
        }
    }
    return bi;
}
// Seeded per-lane scan for a RUN of neighbouring queries (consecutive samples of one ray): the winner of the previous query
// (its slot in the sorted vertex array) gives a tight initial bound, so the "find the best box, scan it" pass is skipped and
// pruning bites from the first box on.  Still exact: the seed is a real vertex, every cluster that could hold a closer (or
// equally close, lower-index) vertex has box distance <= the bound and is scanned.  slot < 0: unseeded.  Returns the vertex
// index, `slot` is updated to the winner's slot.
__device__ __forceinline__ int knn_scan_seeded(const KnnSmem& k, float x, float y, float z, int& slot) {
    float bd = INFINITY;
    int bi = 0x7fffffff, bs = 0;
    auto scan_range = [&](int v0, int v1) {
#pragma unroll 8
        for (int v = v0; v < v1; ++v) {
            const float4 p = k.sv[v];
            const float dx = x - p.x, dy = y - p.y, dz = z - p.z;
            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const int id = __float_as_int(p.w);
            if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; bs = v; }
        }
    };
    auto scan = [&](int c) {
        if (k.qmin) {
#pragma unroll 1
            for (int q = 4 * c; q < 4 * c + 4; ++q)
                if (box_dist2(k.qmin[q], k.qmax[q], x, y, z) <= bd * 1.000001f) scan_range(q * KNN_QUARTER, (q + 1) * KNN_QUARTER);
        } else scan_range(c * KNN_CLUSTER, (c + 1) * KNN_CLUSTER);
    };
    if (slot >= 0) {
        const float4 p = k.sv[slot];
        const float dx = x - p.x, dy = y - p.y, dz = z - p.z;
        bd = fmaf(dz, dz, fmaf(dy, dy, dx * dx)); bi = __float_as_int(p.w); bs = slot;
    } else {
        float lb0 = INFINITY;
        int s0 = 0;
        for (int g = 0; g < k.ns; ++g) { const float lb = box_dist2(k.smin[g], k.smax[g], x, y, z); if (lb < lb0) { lb0 = lb; s0 = g; } }
        lb0 = INFINITY;
        int c0 = s0 * KNN_SUPER;
        for (int c = s0 * KNN_SUPER; c < min(k.nc, (s0 + 1) * KNN_SUPER); ++c) {
            const float lb = box_dist2(k.cmin[c], k.cmax[c], x, y, z);
            if (lb < lb0) { lb0 = lb; c0 = c; }
        }
        scan(c0);                                  // (scanned again below if it still qualifies: harmless)
    }
    for (int g = 0; g < k.ns; ++g) {
        if (box_dist2(k.smin[g], k.smax[g], x, y, z) > bd * 1.000001f) continue;
        for (int c = g * KNN_SUPER; c < min(k.nc, (g + 1) * KNN_SUPER); ++c)
            if (box_dist2(k.cmin[c], k.cmax[c], x, y, z) <= bd * 1.000001f) scan(c);
    }
    slot = bs;
    return bi;
}

// Warp-cooperative form of knn_scan: all 32 lanes hold the SAME query.  Lanes split the cluster boxes (lower bounds kept in
// registers), then every candidate cluster is scanned one vertex per lane (conflict-free LDS.128) and reduced with a
// lexicographic (distance, original index) butterfly.  Clusters are visited in the same order as knn_scan, so the running
// best evolves identically; the result is the same exact argmin with the lowest index on ties.
__device__ __forceinline__ void knn_warp_argmin(float& d, int& id) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float d2 = __shfl_xor_sync(0xffffffffu, d, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, id, o);
        if (d2 < d || (d2 == d && i2 < id)) { d = d2; id = i2; }
    }
}
__device__ __forceinline__ int knn_scan_warp(const KnnSmem& k, float x, float y, float z) {
    const int lane = threadIdx.x & 31;
    float lbs[8];                                          // nc <= 256 (n_verts <= 8192)
    float lb0 = INFINITY;
    int c0 = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = lane + 32 * j;
        float lb = INFINITY;
        if (c < k.nc) lb = box_dist2(k.cmin[c], k.cmax[c], x, y, z);
        lbs[j] = lb;
        if (lb < lb0) { lb0 = lb; c0 = c; }
    }
    knn_warp_argmin(lb0, c0);
    if (c0 >= k.nc) c0 = 0;                                // all boxes at infinite distance (NaN/inf query): any cluster
    float bd = INFINITY;
    int bi = 0x7fffffff;
    auto scan = [&](int c) {
        const float4 p = k.sv[c * KNN_CLUSTER + lane];
        const float dx = x - p.x, dy = y - p.y, dz = z - p.z;
        float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        int id = __float_as_int(p.w);
        knn_warp_argmin(d, id);
        if (d < bd || (d == bd && id < bi)) { bd = d; bi = id; }
    };
    scan(c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (32 * j >= k.nc) break;
        unsigned done = 0u;
        while (true) {
            const bool want = (lane + 32 * j != c0) && (lbs[j] <= bd * 1.000001f);
            const unsigned m = __ballot_sync(0xffffffffu, want) & ~done;
            if (!m) break;
            const int src = __ffs(m) - 1;
            done = (src == 31) ? 0xffffffffu : ((2u << src) - 1u);
            scan(32 * j + src);
        }
    }
    return bi;
}
// Octet form: the warp works on FOUR queries at once, 8 lanes each (all 8 lanes of an octet hold the same query).  Two-level
// pruning like knn_scan: the octet's lanes split the super boxes, then the 8 cluster boxes of a passing super box (one per lane);
// a cluster is scanned 4 vertices per lane and reduced with a 3-step butterfly inside the octet.  ~3x fewer instructions per
// query than knn_scan_warp (shorter butterflies, no replicated box tests), same exact result.  `valid`: the octet has a query.
__device__ __forceinline__ void knn_octet_argmin(float& d, int& id) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        const float d2 = __shfl_xor_sync(0xffffffffu, d, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, id, o);
        if (d2 < d || (d2 == d && i2 < id)) { d = d2; id = i2; }
    }
}
// seed_c (octet-uniform): cluster of the query's previous winner, or < 0 — a marching ray moves little between steps, so scanning
// that cluster first gives a tight bound and the search for the best box is skipped (still exact: every cluster whose box is
// within the bound is scanned below).  *win_c receives the winner's cluster.
__device__ __forceinline__ int knn_scan_octet(const KnnSmem& k, float x, float y, float z, bool valid, int seed_c = -1, int* win_c = nullptr) {
    const int lane = threadIdx.x & 31, sub = lane & 7, oct = lane >> 3;
    float bd = INFINITY;
    int bi = 0x7fffffff, bc = 0;
    // executed by ALL 32 lanes (the butterfly uses full-mask shuffles); `on` is octet-uniform: does this octet scan cluster c
    auto scan = [&](int c, bool on) {
        float d = INFINITY;
        int id = 0x7fffffff;
        if (on) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 p = k.sv[c * KNN_CLUSTER + 4 * sub + i];
                const float dx = x - p.x, dy = y - p.y, dz = z - p.z;
                const float di = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                const int ii = __float_as_int(p.w);
                if (di < d || (di == d && ii < id)) { d = di; id = ii; }
            }
        }
        knn_octet_argmin(d, id);
        if (on && (d < bd || (d == bd && id < bi))) { bd = d; bi = id; bc = c; }
    };
    int c0 = seed_c;
    if (__any_sync(0xffffffffu, seed_c < 0)) {
        // ---- best super box, best cluster in it (for the octets without a seed)
        float lb0 = INFINITY;
        int s0 = 0x7fffffff;
        for (int g = sub; g < k.ns; g += 8) {
            const float lb = box_dist2(k.smin[g], k.smax[g], x, y, z);
            if (lb < lb0) { lb0 = lb; s0 = g; }
        }
        knn_octet_argmin(lb0, s0);
        if (s0 >= k.ns) s0 = 0;
        int cb = s0 * KNN_SUPER + sub;
        float lc = (cb < k.nc) ? box_dist2(k.cmin[cb], k.cmax[cb], x, y, z) : INFINITY;
        knn_octet_argmin(lc, cb);
        if (cb >= k.nc) cb = s0 * KNN_SUPER;
        if (seed_c < 0) c0 = cb;
    }
    scan(c0, valid);
    // ---- every other cluster that can still hold a closer vertex, super box by super box
    const int ns8 = (k.ns + 7) & ~7;
    for (int g0 = 0; g0 < ns8; g0 += 8) {
        const int g = g0 + sub;
        const float lbs = (g < k.ns) ? box_dist2(k.smin[g], k.smax[g], x, y, z) : INFINITY;
        unsigned sdone = 0u;
        while (true) {
            const bool swant = valid && lbs <= bd * 1.000001f;
            const unsigned sm_all = __ballot_sync(0xffffffffu, swant);
            if (!sm_all) break;                                // no octet has a super box left in this group of 8
            const unsigned sm = (sm_all >> (8 * oct)) & 0xffu & ~sdone;
            const bool have = sm != 0u;
            const int sb = have ? (__ffs(sm) - 1) : 0;
            if (!__ballot_sync(0xffffffffu, have)) break;      // every octet has exhausted its bits (all covered by sdone)
            if (have) sdone |= (2u << sb) - 1u;
            // the 8 clusters of super box g0 + sb, one per lane
            const int c = (g0 + sb) * KNN_SUPER + sub;
            const float lbc = (have && c < k.nc && c != c0) ? box_dist2(k.cmin[c], k.cmax[c], x, y, z) : INFINITY;
            unsigned cdone = 0u;
            while (true) {
                const bool cwant = have && lbc <= bd * 1.000001f;
                const unsigned cm = (__ballot_sync(0xffffffffu, cwant) >> (8 * oct)) & 0xffu & ~cdone;
                const bool chave = cm != 0u;
                if (!__ballot_sync(0xffffffffu, chave)) break;
                const int cb = chave ? (__ffs(cm) - 1) : 0;
                if (chave) cdone |= (2u << cb) - 1u;
                scan((g0 + sb) * KNN_SUPER + cb, chave);
            }
        }
    }
    if (win_c) *win_c = bc;
    return bi;
}

// Batch driver: a warp takes `B` consecutive queries (B = 1..32 so that every warp of the grid has work when few queries are
// left); lane l loads query l, the queries are scanned cooperatively one after the other, lane l finishes query l.
// Measured (B200, 15 M queries): with a full warp of queries the per-lane scan (knn_scan, 32 queries in SIMT) is ~3x
// cheaper per query than the cooperative one (shuffle reductions); the cooperative form wins when a warp would otherwise hold
// only a few queries (tail iterations of sphere tracing, training-size batches), where latency, not throughput, counts.
// Measured (512x512 frame, stage times of bench.py): samples along one ray are coherent, the per-lane scan wins beyond ~12
// queries per warp (k_knn_samples 11 ms vs 33 ms cooperative; 7 ms with the super-box pruning).  Rays stay cooperative at
// every size: sending the full-warp launches of the first sphere-tracing iterations through the per-lane scan made the
// tracing stage slower (21.7 vs 15.4 ms) although ncu had timed those launches alone at 0.3 vs 0.9 ms.
constexpr int KNN_COOP_SAMPLES = 12, KNN_COOP_RAYS = 32;
template <int COOP_MAX_B, class LoadQ, class Finish>
__device__ __forceinline__ void knn_warp_batches(const KnnSmem& kk, int n, int B, LoadQ load, Finish fin) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5, total_warps = gridDim.x * wpb, gw = blockIdx.x * wpb + (threadIdx.x >> 5);
    const int nchunks = (n + B - 1) / B;
    for (int c = gw; c < nchunks; c += total_warps) {
        const int i0 = c * B, cnt = min(B, n - i0);
        float x[3] = {0.f, 0.f, 0.f};
        if (lane < cnt) load(i0 + lane, x);
        int mine = 0;
        if (B > COOP_MAX_B) {
            if (lane < cnt) mine = knn_scan(kk, x[0], x[1], x[2]);
        } else {
            if (cnt >= 3) {                                     // four queries at a time, one per octet
                for (int r4 = 0; r4 < cnt; r4 += 4) {
                    const int q = r4 + (lane >> 3);
                    const float qx = __shfl_sync(0xffffffffu, x[0], q & 31), qy = __shfl_sync(0xffffffffu, x[1], q & 31), qz = __shfl_sync(0xffffffffu, x[2], q & 31);
                    const int idx = knn_scan_octet(kk, qx, qy, qz, q < cnt);
                    const int got = __shfl_sync(0xffffffffu, idx, (lane & 3) * 8);      // result of query r4 + (lane & 3)
                    if ((lane >> 2) == (r4 >> 2)) mine = got;
                }
            } else {
                for (int j = 0; j < cnt; ++j) {
                    const float qx = __shfl_sync(0xffffffffu, x[0], j), qy = __shfl_sync(0xffffffffu, x[1], j), qz = __shfl_sync(0xffffffffu, x[2], j);
                    const int idx = knn_scan_warp(kk, qx, qy, qz);
                    if (lane == j) mine = idx;
                }
            }
        }
        if (lane < cnt) fin(i0 + lane, x, mine);
    }
}
// queries per warp: spread over all warps of the grid while that keeps a warp at <= COOP_MAX_B queries, else full warps
template <int COOP_MAX_B>
__device__ __forceinline__ int knn_batch_size(int n) {
    const int total_warps = gridDim.x * (blockDim.x >> 5);
    const int b = max(1, (n + total_warps - 1) / total_warps);
    return b > COOP_MAX_B ? 32 : b;
}

// NN-skinning inverse of one posed point x (incl. trans): T = sum_j W[idx][j] B_j, x_hat = T^-1 (x - trans)
__device__ __forceinline__ void nn_inverse_skinning(const FrameParams& fp, int idx, const float* x, float* T12, float* s, float* x_hat) {
    float wj[NJ];
    const float4* wp = reinterpret_cast<const float4*>(fp.smpl_w + (size_t)idx * NJ);
#pragma unroll
    for (int q = 0; q < NJ / 4; ++q) { const float4 t = __ldg(wp + q); wj[4 * q] = t.x; wj[4 * q + 1] = t.y; wj[4 * q + 2] = t.z; wj[4 * q + 3] = t.w; }
    blend_T(wj, fp.bone_T, T12, s);
    const float xl[3] = {x[0] - fp.trans[0], x[1] - fp.trans[1], x[2] - fp.trans[2]};
    affine_inverse_apply(T12, *s, xl, x_hat);
}


__device__ __forceinline__ void iso_residual(const FrameParams& fp, const Work& w, int r, const float* u, const float* lg32,
                                             float sdf_raw, float* g, float* T12) {
    float xb[3];
    skin_point(fp, lg32, u, T12, xb);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float xbar = w.ray_dirs[3 * r + k] * u[3] + fp.cam_loc[k];
        g[1 + k] = xb[k] - (xbar - fp.trans[k]);
    }
    g[0] = sdf_to_metres(sdf_raw, fp.cmin, fp.cmax);
}


constexpr int LDA_SKIN = 132;
__device__ __forceinline__ void corr_finalize(const FrameParams& fp, const Work& w, const BroydenState<3>& st) {
    const size_t sl = (size_t)st.owner;
    float xn[3];
    normalize3(fp, st.best_x, xn);
    w.smp_xn[3 * sl] = xn[0]; w.smp_xn[3 * sl + 1] = xn[1]; w.smp_xn[3 * sl + 2] = xn[2];
    float4* Tp = reinterpret_cast<float4*>(w.smp_T + 12 * sl);
    Tp[0] = make_float4(st.best_T[0], st.best_T[1], st.best_T[2], st.best_T[3]);
    Tp[1] = make_float4(st.best_T[4], st.best_T[5], st.best_T[6], st.best_T[7]);
    Tp[2] = make_float4(st.best_T[8], st.best_T[9], st.best_T[10], st.best_T[11]);
    w.smp_conv[sl] = (st.best_n < CVG_THRESH) ? 1 : 0;
}


}  // namespace arah
