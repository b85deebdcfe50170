// arah_api.cu — C ABI (include/arah_b200.h), workspace management, weight packing and the launch sequence.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/arah_b200.h"
#include "arah_kernels.cuh"
#include "arah_umma.cuh"
#include "arah_sdf3x.cuh"
#include "arah_iso_init_tc.cuh"
#include "arah_train_cuda.cuh"
#include "arah_root.h"
#include "arah_shade.h"
#include <stdlib.h>

using namespace arah;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
extern "C" int arah_internal_fail(int code, const char* msg) { return fail(code, msg ? msg : ""); }      // arah_mesh.cu
#define CU(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(ARAH_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }
static inline unsigned grid_min(size_t a, size_t b) { return (unsigned)(a < b ? a : b); }

extern "C" const char* arah_last_error(void) { return g_err.c_str(); }
extern "C" int arah_version(void) { return 100; }

// ------------------------------------------------------------------------------------------------ pack kernels
// dst[k][n] = (k < K && n < N) ? src[n][col(k)] : 0   with col(k) = (k < split) ? k + off_lo : k - split + off_hi
__global__ void k_pack_transpose(const float* __restrict__ src, int src_ld, float* __restrict__ dst, int K, int N,
                                 int Kpad, int Npad, int split, int off_lo, int off_hi) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Kpad * Npad) return;
    const int k = idx / Npad, n = idx % Npad;
    float v = 0.f;
    if (k < K && n < N) {
        const int col = (k < split) ? (k + off_lo) : (k - split + off_hi);
        v = src[(size_t)n * src_ld + col];
    }
    dst[idx] = v;
}
// b'[o] = b[o] + sum_j W[o][col0 + j] * latent[j]   (sequential j; folds the per-frame-constant latent)
__global__ void k_fold_latent(const float* __restrict__ W, int ld, int col0, const float* __restrict__ latent, int L,
                              const float* __restrict__ b, float* __restrict__ out, int n_out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    float acc = 0.f;
    for (int j = 0; j < L; ++j) acc = fmaf(W[(size_t)o * ld + col0 + j], latent[j], acc);
    out[o] = b[o] + acc;
}
__global__ void k_copy_pad(const float* __restrict__ src, float* __restrict__ dst, int n, int npad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) dst[i] = (i < n) ? src[i] : 0.f;
}
__global__ void k_verts4(const float* __restrict__ v3, float4* __restrict__ v4, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v4[i] = make_float4(v3[3 * i], v3[3 * i + 1], v3[3 * i + 2], 0.f);
}
// F = 30 f, G = 30 (f b + phi): the FiLM-sine argument becomes fma(acc, F, G) in the tensor-core epilogue
__global__ void k_pack_film(const float* __restrict__ b, const float* __restrict__ f, const float* __restrict__ ph, float* __restrict__ F, float* __restrict__ G) {
    const int c = threadIdx.x;
    F[c] = 30.0f * f[c];
    G[c] = 30.0f * (f[c] * b[c] + ph[c]);
}
__global__ void k_read_b6(const float* __restrict__ b6, float* __restrict__ dst) { dst[0] = b6[0]; }

// ------------------------------------------------------------------------------------------------ unit kernels
__global__ void __launch_bounds__(256, 1) k_eval_sdf(FrameParams fp, const float* xn, int n, float* sdf, float* grad, float* feat_out, float* scratch) {
    extern __shared__ __align__(128) float smem[];
    if ((int)blockIdx.x * TM >= n) return;
    TileSmem s = carve(smem, LDA_SDF);
    WPipe wp;
    wpipe_init(wp, s.wbuf, s.bars);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* cf = scratch + (size_t)blockIdx.x * 7 * TM * SDF_H;
    float (*g)[4] = reinterpret_cast<float (*)[4]>(s.logits);
    for (int tile = blockIdx.x; tile * TM < n; tile += gridDim.x) {
        if (tid < TM) {
            const int i = tile * TM + tid;
            for (int k = 0; k < 3; ++k) s.xs[tid][k] = (i < n) ? xn[3 * i + k] : 0.f;
            s.xs[tid][3] = 0.f;
        }
        __syncthreads();
        sdf_tile_forward<PLAIN>(fp, s.xs, s.A, LDA_SDF, wp, s.sdfo, cf, 0.f);
        if (feat_out) {
            for (int r = 0; r < 8; ++r) {
                const int i = tile * TM + warp * 8 + r;
                if (i < n) for (int c = lane; c < SDF_H; c += 32) feat_out[(size_t)i * SDF_H + c] = s.A[(warp * 8 + r) * LDA_SDF + c];
            }
        }
        __syncwarp();
        sdf_tile_backward(fp, s.A, LDA_SDF, wp, cf, g);
        __syncthreads();
        if (tid < TM) {
            const int i = tile * TM + tid;
            if (i < n) {
                sdf[i] = s.sdfo[tid];
                if (grad) for (int k = 0; k < 3; ++k) grad[3 * i + k] = g[tid][k];
            }
        }
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256, 2) k_eval_skin(FrameParams fp, const float* x_hat, int n, float* weights, float* x_bar) {
    extern __shared__ __align__(128) float smem[];
    if ((int)blockIdx.x * TM >= n) return;
    TileSmem s = carve(smem, LDA_SKIN);
    WPipe wp;
    wpipe_init(wp, s.wbuf, s.bars);
    const int tid = threadIdx.x;
    for (int tile = blockIdx.x; tile * TM < n; tile += gridDim.x) {
        float xh[3] = {0.f, 0.f, 0.f};
        const int i = tile * TM + tid;
        if (tid < TM) {
            float xn[3] = {0.f, 0.f, 0.f};
            if (i < n) { xh[0] = x_hat[3 * i]; xh[1] = x_hat[3 * i + 1]; xh[2] = x_hat[3 * i + 2]; normalize3(fp, xh, xn); }
            s.xs[tid][0] = xn[0]; s.xs[tid][1] = xn[1]; s.xs[tid][2] = xn[2]; s.xs[tid][3] = 0.f;
        }
        __syncthreads();
        skin_tile_forward<PLAIN>(fp, s.xs, s.A, LDA_SKIN, wp, s.logits, 0.f);
        __syncthreads();
        if (tid < TM && i < n) {
            float lg[25], wj[NJ], T12[12], xb[3];
            for (int k = 0; k < 25; ++k) lg[k] = s.logits[tid][k] * 20.0f;
            hierarchical_softmax(lg, wj);
            blend_T(wj, fp.bone_T, T12, nullptr);
            apply_T(T12, xh, xb);
            for (int k = 0; k < NJ; ++k) weights[(size_t)i * NJ + k] = wj[k];
            for (int k = 0; k < 3; ++k) x_bar[3 * i + k] = xb[k];
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------ tcgen05 packing + probe
// dst chunk c (k in [32c, 32c+32)) = shared-memory image of a K-major SWIZZLE_128B tile of B[N][K]:
//   float index  c*N*32 + (n/8)*256 + (n%8)*32 + ((j ^ (n%8))*4) + e   <-   src[n][col(32c + 4j + e)]   (0 beyond K)
// col(k) = (k < split) ? k + off_lo : k - split + off_hi   (same column permutation as k_pack_transpose)
__global__ void k_pack_umma(const float* __restrict__ src, int src_ld, float* __restrict__ dst, int N, int K, int nchunks,
                            int split, int off_lo, int off_hi, int transpose_src) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nchunks * N * 32) return;
    const int c = idx / (N * 32), rem = idx % (N * 32);
    const int n = rem / 32, q = rem % 32, j = q / 4, e = q % 4;
    const int k = 32 * c + 4 * j + e;
    float v = 0.f;
    if (k < K) {
        const int col = (k < split) ? (k + off_lo) : (k - split + off_hi);
        v = transpose_src ? src[(size_t)col * src_ld + n] : src[(size_t)n * src_ld + col];
        v = tf32_rn(v);                  // round once here so the tensor core's operand truncation is exact
    }
    dst[(size_t)c * N * 32 + (n >> 3) * 256 + (n & 7) * 32 + ((j ^ (n & 7)) << 2) + e] = v;
}

// split-precision variant: per K-chunk [hi image | lo image], Npad rows (rows >= N are zero)
__global__ void k_pack_umma_x3(const float* __restrict__ src, int src_ld, float* __restrict__ dst, int N, int Npad, int K, int nchunks) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nchunks * Npad * 32) return;
    const int c = idx / (Npad * 32), rem = idx % (Npad * 32);
    const int n = rem / 32, q = rem % 32, j = q / 4, e = q % 4;
    const int k = 32 * c + 4 * j + e;
    float v = 0.f;
    if (k < K && n < N) v = src[(size_t)n * src_ld + k];
    const float hi = tf32_rn(v), lo = tf32_rn(v - hi);
    const size_t o = (size_t)c * Npad * 32 * 2 + (n >> 3) * 256 + (n & 7) * 32 + ((j ^ (n & 7)) << 2) + e;
    dst[o] = hi;
    dst[o + (size_t)Npad * 32] = lo;
}

// D[128][N] = A[128][K] . W[N][K]^T on the tensor cores (TF32 operands, fp32 accumulate); validates descriptors/swizzle
__global__ void __launch_bounds__(256, 1) k_umma_probe(const float* __restrict__ A, const float* __restrict__ Wsw, int K, int N, float* __restrict__ D, int a_in_tmem) {
    extern __shared__ uint8_t raw_smem[];
    const uint32_t base = (smem_u32(raw_smem) + 1023u) & ~1023u;
    float* sm = reinterpret_cast<float*>(raw_smem + (base - smem_u32(raw_smem)));
    float* Abuf = sm;                                   // 8 chunks x 16 KB
    float* ring = sm + 8 * A_CHUNK_FLOATS;              // 2 x 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 2 * RING_SLOT_FLOATS);   // full[2], empty[2], done
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tslot, 512);
    if (tid < UM && !a_in_tmem) {
        for (int c = 0; c < K / UK; ++c) {
            float v[32];
            for (int i = 0; i < 32; ++i) v[i] = A[(size_t)tid * K + c * UK + i];
            a_store_chunk(Abuf, tid, c, v);
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tslot;
    const int q = warp & 3, half = warp >> 2;
    if (a_in_tmem) {          // A via tcgen05.st into TMEM columns [256, 256+K); D in columns [0, N)
        if (half == 0) {
            for (int c = 0; c < K / UK; ++c) {
                float v[32];
                for (int i = 0; i < 32; ++i) v[i] = A[(size_t)(32 * q + lane) * K + c * UK + i];
                a_tmem_store(tbase + ((uint32_t)(32 * q) << 16) + 256u + (uint32_t)(c * UK), v);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (tid == 0) {
        URing rg; rg.buf = ring; rg.full = bars; rg.empty = bars + 2; rg.fill_cnt = 0; rg.mma_cnt = 0;
        if (!a_in_tmem) umma_layer_issue(rg, Abuf, Wsw, K / UK, N, tbase, 0u, &bars[4]);
        else {
            const uint32_t idesc = umma_idesc_tf32(UM, N);
            for (int c = 0; c < K / UK; ++c) {
                mbar_expect_tx(&bars[0], (uint32_t)N * UK * 4);
                bulk_g2s(ring, Wsw + (size_t)c * N * UK, (uint32_t)N * UK * 4, &bars[0]);
                mbar_wait(&bars[0], c & 1);
                tc_fence_after();
                for (int k = 0; k < 4; ++k)
                    umma_tf32_ts(tbase, tbase + 256u + (uint32_t)(c * UK + k * 8), umma_smem_desc_sw128(smem_u32(ring) + k * 32), idesc, (c > 0 || k > 0) ? 1u : 0u);
                umma_commit(&bars[2]);
                mbar_wait(&bars[2], c & 1);          // serialise: the probe reuses one slot
            }
            umma_commit(&bars[4]);
        }
    }
    mbar_wait(&bars[4], 0);
    tc_fence_after();
    for (int b = 0; b < (N / 2) / 32; ++b) {
        const int col0 = half * (N / 2) + 32 * b;
        float v[32];
        tmem_ld32(tbase + ((uint32_t)(32 * q) << 16) + (uint32_t)col0, v);
        for (int i = 0; i < 32; ++i) D[(size_t)(32 * q + lane) * N + col0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

extern "C" int arah_debug_umma_f16(const float* A, const float* W, int32_t K, int32_t N, float* D, int32_t mode, void* stream) {
    if (!A || !W || !D) return fail(ARAH_EINVAL, "null buffer");
    if ((K != 64 && K != 128) || (N != 32 && N != 128 && N != 256)) return fail(ARAH_EINVAL, "K in {64,128}, N in {32,128,256}");
    CU(root_init());
    CU(root_probe_f16(A, W, K, N, D, mode, (cudaStream_t)stream));
    return ARAH_OK;
}

extern "C" int arah_debug_umma_gemm(const float* A, const float* W, int32_t K, int32_t N, float* D, int32_t a_in_tmem, void* stream) {
    if (!A || !W || !D) return fail(ARAH_EINVAL, "null buffer");
    if (K <= 0 || K > 256 || (K % 32) != 0 || (N != 256 && N != 128)) return fail(ARAH_EINVAL, "K must be a multiple of 32 <= 256, N in {128,256}");
    cudaStream_t st = (cudaStream_t)stream;
    float* Wsw = nullptr;
    CU(cudaMalloc(&Wsw, (size_t)K * N * 4));
    k_pack_umma<<<cdiv((size_t)K * N, 256), 256, 0, st>>>(W, K, Wsw, N, K, K / 32, K, 0, 0, 0);
    const int smem = (8 * A_CHUNK_FLOATS + 2 * RING_SLOT_FLOATS) * 4 + 128 + 1024;
    CU(cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_umma_probe<<<1, 256, smem, st>>>(A, Wsw, K, N, D, a_in_tmem);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    CU(cudaFree(Wsw));
    return ARAH_OK;
}

// ------------------------------------------------------------------------------------------------ handle
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t n) {
        if (n <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        if (cudaMalloc(&p, n) != cudaSuccess) return -1;
        bytes = n;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

struct ArahHandle {
    ArahConfig cfg;
    int n_sms = 148;
    bool frame_set = false, rendered = false, have_smpl_w = false;
    FrameParams fp;
    // packed weights (one arena)
    DevBuf arena;
    float* sdf_Wt[6]; float* sdf_W[6]; float* sdf_b[6]; float* sdf_w6; float* sdf_freq; float* sdf_phase; float* d_b6;
    float* skin_Wt[5]; float* skin_b[5];
    float* col_Wt0; float* col_Wt1; float* col_Wt2; float* col_Wt3a; float* col_Wt3b; float* col_Wt4; float* col_W5; float* col_b[6];
    float* bone_T; float4* verts4; float* verts3; float* smpl_w;
    float* knn_sv; float* knn_cmin; float* knn_cmax; KnnIndex knn;
    // tensor-core shading: pre-swizzled chunk images (arah_umma.cuh)
    float* tc_F; float* tc_G;  // FiLM factors 30 f, 30 (f b + phi) of the six SDF layers (k_pack_film)
    float* tc_skin_hid[3]; float* tc_skin_out;
    SkinTC sk;
    float* tc_sdf3x[5];
    SdfTC sd;
    int knn_seed = 2;          // k_knn_samples: 2 = ray-major seeded per-lane 1-NN for full frames, runs of 4 for small batches; 3 = always ray-major; 1 = runs of 4 on-samples (round 1); 0 = unseeded
    int trace_knn = 0;         // k_trace_persist: 0 = octet-cooperative 1-NN (measured faster), 1 = seeded one-row-per-lane scan
    SkinF16Dev skin16{};
    int trace_persist = 1;     // k_trace_persist: sphere tracing as one persistent kernel (1-NN + SDF per step, resident rays)
    SdfF16Dev sdf16{};
    Shade16Dev shade16_img{};
    bool shade_cull_ran = false;
    int shade_cull = 1;        // exact alpha cull before the gradient / colour pass (k_alpha_cull)
    // workspace
    DevBuf ws, scratch, io_in, io_out;
    Work w;
    int cap_rays = 0;
    int64_t launches = 0;
    int last_P = 0;
    int64_t pack_launches = 0;
    bool profile = false, profiled = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // training (arah_train.h): raw reference-layout weight copies + the engine's saved activations
    bool training = false, train_traced = false;
    int train_precision = ARAH_TRAIN_3XTF32;
    DevBuf raw;
    arah::train::AllParams tp{};
    arah::train::Session<arah::train::CudaBK>* sess = nullptr;
    int col_d0 = 0;
};

extern "C" int arah_destroy(ArahHandle* h);

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int alloc_arena(ArahHandle* h) {
    const int L = h->cfg.latent_dim;
    (void)L;
    size_t off = 0;
    auto take = [&](size_t floats) { size_t o = off; off += align_up(floats * 4, 256); return o; };
    std::vector<std::pair<float**, size_t>> slots;
    auto reg = [&](float** p, size_t floats) { slots.push_back({p, take(floats)}); };
    reg(&h->sdf_Wt[0], 3 * 256); reg(&h->sdf_W[0], 256 * 3); reg(&h->sdf_b[0], 256);
    for (int l = 1; l < 6; ++l) { reg(&h->sdf_Wt[l], 256 * 256); reg(&h->sdf_W[l], 256 * 256); reg(&h->sdf_b[l], 256); }
    reg(&h->sdf_w6, 256); reg(&h->sdf_freq, 6 * 256); reg(&h->sdf_phase, 6 * 256); reg(&h->d_b6, 64);
    reg(&h->skin_Wt[0], 3 * 128); reg(&h->skin_b[0], 128);
    for (int l = 1; l < 4; ++l) { reg(&h->skin_Wt[l], 128 * 128); reg(&h->skin_b[l], 128); }
    reg(&h->skin_Wt[4], 128 * 32); reg(&h->skin_b[4], 32);
    reg(&h->col_Wt0, COL_IN_PAD * 256); reg(&h->col_Wt1, 256 * 256); reg(&h->col_Wt2, 256 * 128);
    reg(&h->col_Wt3a, COL_IN_PAD * 256); reg(&h->col_Wt3b, 128 * 256); reg(&h->col_Wt4, 256 * 256); reg(&h->col_W5, 3 * 256);
    reg(&h->tc_F, 6 * 256); reg(&h->tc_G, 6 * 256);
    for (int l = 0; l < 5; ++l) reg(&h->tc_sdf3x[l], 8 * 256 * 32 * 2);
    for (int l = 0; l < 3; ++l) reg(&h->tc_skin_hid[l], 4 * 128 * 32 * 2);
    reg(&h->tc_skin_out, 4 * 32 * 32 * 2);
    float *s16_hi = nullptr, *s16_lo = nullptr;
    reg(&s16_hi, SKIN_F16_IMAGE_BYTES / 4); reg(&s16_lo, SKIN_F16_IMAGE_BYTES / 4); reg(&h->skin16.scale, 8);
    float *d16_hi = nullptr, *d16_lo = nullptr;
    reg(&d16_hi, SDF_F16_DEV_BYTES / 4); reg(&d16_lo, SDF_F16_DEV_BYTES / 4); reg(&h->sdf16.scale, 16);
    float *sh16_bwd = nullptr, *sh16_col = nullptr;
    reg(&sh16_bwd, SHADE16_BWD_DEV_BYTES / 4); reg(&sh16_col, SHADE16_COL_DEV_BYTES / 4);
    reg(&h->col_b[0], 256); reg(&h->col_b[1], 256); reg(&h->col_b[2], 256); reg(&h->col_b[3], 256); reg(&h->col_b[4], 256); reg(&h->col_b[5], 64);
    reg(&h->knn_sv, (size_t)((h->cfg.n_verts + 31) / 32) * 32 * 4); reg(&h->knn_cmin, (size_t)((h->cfg.n_verts + 31) / 32) * 4);
    reg(&h->knn_cmax, (size_t)((h->cfg.n_verts + 31) / 32) * 4);
    reg(&h->bone_T, 24 * 16);
    float* v4 = nullptr;
    reg(&v4, (size_t)h->cfg.n_verts * 4);
    reg(&h->verts3, (size_t)h->cfg.n_verts * 3);
    reg(&h->smpl_w, (size_t)h->cfg.n_verts * 24);
    if (h->arena.ensure(off) != 0) return -1;
    for (auto& s : slots) *s.first = reinterpret_cast<float*>(static_cast<char*>(h->arena.p) + s.second);
    h->verts4 = reinterpret_cast<float4*>(static_cast<char*>(h->arena.p) + slots[slots.size() - 3].second);
    h->skin16.hi = s16_hi; h->skin16.lo = s16_lo;
    h->sdf16.hi = d16_hi; h->sdf16.lo = d16_lo;
    h->shade16_img.bwd = sh16_bwd; h->shade16_img.col = sh16_col;
    return 0;
}

static int ensure_workspace(ArahHandle* h, int P) {
    if (P <= h->cap_rays) return 0;
    const int cap = (int)align_up((size_t)P, 1024);
    const size_t S = h->cfg.n_steps, PS = (size_t)cap * S;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    const size_t o_t = take(cap * 4), o_fl = take(cap), o_cur = take(cap * sizeof(RayCur)), o_iso = take(cap * sizeof(BroydenState<4>));
    const size_t o_conv = take(cap), o_dist = take(cap * 4), o_pn = take(cap * 12);
    const size_t o_z = take(PS * 4), o_xn = take(PS * 12), o_T = take(PS * 48), o_sc = take(PS), o_sdf = take(PS * 4), o_rgb = take(PS * 12);
    const size_t o_cs = take(PS * sizeof(BroydenState<3>));
    const size_t o_la = take(PS * 4), o_lb = take(PS * 4), o_on = take(PS * 4), o_sh = take(PS * 4), o_ctr = take(C_COUNT * 4 + 64), o_clk = take(32 * 8);
    const size_t o_rb = take(cap * 4);
    if (h->ws.ensure(off) != 0) return -1;
    char* b = static_cast<char*>(h->ws.p);
    Work& w = h->w;
    w.S = (int)S;
    w.ray_t = (float*)(b + o_t); w.ray_flags = (uint8_t*)(b + o_fl); w.ray_cur = (RayCur*)(b + o_cur);
    w.iso_state = (BroydenState<4>*)(b + o_iso); w.ray_conv = (uint8_t*)(b + o_conv); w.ray_dist = (float*)(b + o_dist);
    w.ray_pnorm = (float*)(b + o_pn); w.z_vals = (float*)(b + o_z); w.smp_xn = (float*)(b + o_xn); w.smp_T = (float*)(b + o_T);
    w.smp_conv = (uint8_t*)(b + o_sc); w.smp_sdf = (float*)(b + o_sdf); w.smp_rgb = (float*)(b + o_rgb);
    w.corr_state = (BroydenState<3>*)(b + o_cs); w.listA = (int*)(b + o_la); w.listB = (int*)(b + o_lb);
    w.on_list = (int*)(b + o_on); w.shade_list = (int*)(b + o_sh); w.counters = (int*)(b + o_ctr); w.ray_on_base = (int*)(b + o_rb);
    w.phase_clk = (unsigned long long*)(b + o_clk);
    h->cap_rays = cap;
    return 0;
}

extern "C" int arah_create(const ArahConfig* cfg, ArahHandle** out) {
    if (!cfg || !out) return fail(ARAH_EINVAL, "null argument");
    if (cfg->n_steps <= 0 || cfg->n_steps > MAX_STEPS) return fail(ARAH_EINVAL, "n_steps must be in [1,256]");
    if (cfg->near_samples + 1 + cfg->far_samples > cfg->n_steps)
        return fail(ARAH_EINVAL, "near_samples + 1 + far_samples must be <= n_steps (ray_tracing.py:336,346)");
    if (cfg->near_samples < 0 || cfg->far_samples < 0 || (cfg->near_samples == 0 && cfg->far_samples == 0))
        return fail(ARAH_EINVAL, "need near_samples > 0 or far_samples > 0 (ray_tracing.py:107)");
    if (cfg->shade_mode != ARAH_SHADE_TF32 && cfg->shade_mode != ARAH_SHADE_FP32) return fail(ARAH_EINVAL, "shade_mode must be ARAH_SHADE_TF32 or ARAH_SHADE_FP32");
    if (cfg->root_mode != ARAH_ROOT_3XTF32 && cfg->root_mode != ARAH_ROOT_FP32) return fail(ARAH_EINVAL, "root_mode must be ARAH_ROOT_3XTF32 or ARAH_ROOT_FP32");
    if (cfg->latent_dim < 0 || cfg->latent_dim > 512) return fail(ARAH_EINVAL, "latent_dim must be in [0,512]");
    if (cfg->n_verts <= 0 || cfg->n_verts > 8192) return fail(ARAH_EINVAL, "n_verts must be in [1, 8192] (kNN index: one-block Morton sort + shared-memory clusters)");
    CU(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major < 10) return fail(ARAH_EINVAL, "arah_b200 needs an sm_100-class GPU");
    ArahHandle* h = new ArahHandle();
    struct Guard { ArahHandle* h; ~Guard() { if (h) arah_destroy(h); } } guard{h};     // any early return below releases the handle and its buffers
    h->cfg = *cfg;
    h->n_sms = prop.multiProcessorCount;
    memset(&h->w, 0, sizeof(h->w));
    if (alloc_arena(h) != 0) return fail(ARAH_ENOMEM, "weight arena allocation failed");
    if (ensure_workspace(h, cfg->max_rays > 0 ? cfg->max_rays : 4096) != 0) return fail(ARAH_ENOMEM, "workspace allocation failed");
    const size_t scr_fp32 = (size_t)7 * TM * SDF_H * 4, scr_tc = shade16_scratch_bytes_per_cta();
    if (h->scratch.ensure((size_t)h->n_sms * (scr_fp32 > scr_tc ? scr_fp32 : scr_tc)) != 0) return fail(ARAH_ENOMEM, "scratch allocation failed");
    h->w.scratch = (float*)h->scratch.p;
    CU(cudaFuncSetAttribute(k_trace_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(LDA_SDF)));
    CU(cudaFuncSetAttribute(k_iso_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(LDA_SDF)));
    CU(cudaFuncSetAttribute(k_iso_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(LDA_SDF)));
    CU(cudaFuncSetAttribute(k_eval_sdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(LDA_SDF)));
    CU(cudaFuncSetAttribute(k_corr_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(LDA_SKIN)));
    CU(cudaFuncSetAttribute(k_eval_skin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(LDA_SKIN)));
    CU(cudaFuncSetAttribute(k_shade, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shade_smem_bytes()));
    h->shade_cull = cfg->shade_cull == ARAH_CULL_OFF ? 0 : 1;
    if (const char* e = getenv("ARAH_SHADE_CULL")) h->shade_cull = atoi(e) != 0;
    if (const char* e = getenv("ARAH_TRACE_PERSIST")) h->trace_persist = atoi(e) != 0;
    CU(shade16_init());
    if (!root_trace_fits(cfg->n_verts)) h->trace_persist = 0;      // vertex index + weight ring must fit in 227 KB of shared memory
    CU(root_init());
    CU(cudaFuncSetAttribute(k_trace_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trace_tc3_smem_bytes()));
    CU(cudaFuncSetAttribute(k_iso_init_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trace_tc3_smem_bytes()));
    if (const char* e = getenv("ARAH_KNN_SEED")) h->knn_seed = atoi(e);
    if (const char* e = getenv("ARAH_TRACE_KNN")) h->trace_knn = atoi(e) != 0;
    CU(cudaFuncSetAttribute(k_knn_rays, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem_bytes(cfg->n_verts)));
    CU(cudaFuncSetAttribute(k_knn_samples, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(knn_smem_bytes(cfg->n_verts) + knn_quarter_smem_bytes(cfg->n_verts))));
    CU(cudaFuncSetAttribute(k_knn_points, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem_bytes(cfg->n_verts)));
    CU(cudaFuncSetAttribute(k_knn_build, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
    guard.h = nullptr;
    *out = h;
    return ARAH_OK;
}

extern "C" int arah_set_profiling(ArahHandle* h, int32_t enable) {
    if (!h) return fail(ARAH_EINVAL, "null handle");
    if (enable && !h->ev[0]) for (int i = 0; i < 6; ++i) CU(cudaEventCreate(&h->ev[i]));
    h->profile = enable != 0;
    return ARAH_OK;
}

extern "C" int arah_destroy(ArahHandle* h) {
    if (!h) return ARAH_OK;
    for (int i = 0; i < 6; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    h->arena.release(); h->ws.release(); h->scratch.release(); h->io_in.release(); h->io_out.release(); h->raw.release();
    if (h->sess) { h->sess->release(); delete h->sess; }
    delete h;
    return ARAH_OK;
}


extern "C" int arah_set_frame(ArahHandle* h, const ArahFrame* f, void* stream_) {
    if (!h || !f) return fail(ARAH_EINVAL, "null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    CU(cudaSetDevice(h->cfg.device));
    for (int l = 0; l < 7; ++l) if (!f->sdf_W[l] || !f->sdf_b[l]) return fail(ARAH_EINVAL, "sdf weights missing");
    for (int l = 0; l < 5; ++l) if (!f->skin_W[l] || !f->skin_b[l]) return fail(ARAH_EINVAL, "skinning weights missing");
    for (int l = 0; l < 6; ++l) if (!f->col_W[l] || !f->col_b[l]) return fail(ARAH_EINVAL, "colour weights missing");
    if (!f->sdf_freq || !f->sdf_phase || !f->bone_transforms || !f->smpl_verts) return fail(ARAH_EINVAL, "frame buffers missing");
    if (h->cfg.latent_dim > 0 && !f->latent) return fail(ARAH_EINVAL, "latent missing");
    if (!f->smpl_weights && !h->have_smpl_w) return fail(ARAH_EINVAL, "smpl_weights required on the first frame");
    const int L = h->cfg.latent_dim;
    const int din = 3 + 27 + 3 + 256 + L;     // colour-net input width in the reference's order [x|PE|n|feat|latent]
    int64_t npack = 0;
    auto tp = [&](const float* src, int ld, float* dst, int K, int N, int Kp, int Np, int split, int lo, int hi) {
        k_pack_transpose<<<cdiv((size_t)Kp * Np, 256), 256, 0, st>>>(src, ld, dst, K, N, Kp, Np, split, lo, hi);
        ++npack;
    };
    // SDF
    tp(f->sdf_W[0], 3, h->sdf_Wt[0], 3, 256, 3, 256, 3, 0, 0);
    CU(cudaMemcpyAsync(h->sdf_W[0], f->sdf_W[0], 256 * 3 * 4, cudaMemcpyDeviceToDevice, st));
    for (int l = 1; l < 6; ++l) {
        tp(f->sdf_W[l], 256, h->sdf_Wt[l], 256, 256, 256, 256, 256, 0, 0);
        CU(cudaMemcpyAsync(h->sdf_W[l], f->sdf_W[l], 256 * 256 * 4, cudaMemcpyDeviceToDevice, st));
    }
    for (int l = 0; l < 6; ++l) CU(cudaMemcpyAsync(h->sdf_b[l], f->sdf_b[l], 256 * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(h->sdf_w6, f->sdf_W[6], 256 * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(h->sdf_freq, f->sdf_freq, 6 * 256 * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(h->sdf_phase, f->sdf_phase, 6 * 256 * 4, cudaMemcpyDeviceToDevice, st));
    // the scalar output bias stays on the device: kernels read it through a pointer (no D2H, no stream synchronisation here)
    CU(cudaMemcpyAsync(h->d_b6, f->sdf_b[6], 4, cudaMemcpyDeviceToDevice, st));
    const float* b6 = h->d_b6;
    // skinning
    tp(f->skin_W[0], 3, h->skin_Wt[0], 3, 128, 3, 128, 3, 0, 0);
    for (int l = 1; l < 4; ++l) tp(f->skin_W[l], 128, h->skin_Wt[l], 128, 128, 128, 128, 128, 0, 0);
    tp(f->skin_W[4], 128, h->skin_Wt[4], 128, 25, 128, 32, 128, 0, 0);
    for (int l = 0; l < 4; ++l) CU(cudaMemcpyAsync(h->skin_b[l], f->skin_b[l], 128 * 4, cudaMemcpyDeviceToDevice, st));
    k_copy_pad<<<1, 32, 0, st>>>(f->skin_b[4], h->skin_b[4], 25, 32); ++npack;
    // colour: reference input order [x 3 | PE 27 | n 3 | feat 256 | latent L]; ours [feat | x | PE | n]
    tp(f->col_W[0], din, h->col_Wt0, COL_IN, 256, COL_IN_PAD, 256, 256, 33, 0);
    tp(f->col_W[1], 256, h->col_Wt1, 256, 256, 256, 256, 256, 0, 0);
    tp(f->col_W[2], 256, h->col_Wt2, 256, 128, 256, 128, 256, 0, 0);
    tp(f->col_W[3], din + 128, h->col_Wt3a, COL_IN, 256, COL_IN_PAD, 256, 256, 33, 0);
    tp(f->col_W[3], din + 128, h->col_Wt3b, 128, 256, 128, 256, 128, din, 0);
    tp(f->col_W[4], 256, h->col_Wt4, 256, 256, 256, 256, 256, 0, 0);
    CU(cudaMemcpyAsync(h->col_W5, f->col_W[5], 3 * 256 * 4, cudaMemcpyDeviceToDevice, st));
    if (L > 0) {
        k_fold_latent<<<1, 256, 0, st>>>(f->col_W[0], din, 289, f->latent, L, f->col_b[0], h->col_b[0], 256);
        k_fold_latent<<<1, 256, 0, st>>>(f->col_W[3], din + 128, 289, f->latent, L, f->col_b[3], h->col_b[3], 256);
        npack += 2;
    } else {
        CU(cudaMemcpyAsync(h->col_b[0], f->col_b[0], 256 * 4, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(h->col_b[3], f->col_b[3], 256 * 4, cudaMemcpyDeviceToDevice, st));
    }
    CU(cudaMemcpyAsync(h->col_b[1], f->col_b[1], 256 * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(h->col_b[2], f->col_b[2], 128 * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(h->col_b[4], f->col_b[4], 256 * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(h->col_b[5], f->col_b[5], 3 * 4, cudaMemcpyDeviceToDevice, st));
    const bool tc_shade = h->cfg.shade_mode == ARAH_SHADE_TF32, tc_root = h->cfg.root_mode == ARAH_ROOT_3XTF32;
    if (tc_shade || tc_root) { long long n = 0; CU(root_pack_sdf_f16(f->sdf_W, h->sdf16, st, &n)); npack += n; }      // fp16 hi / lo images of the SDF
    if (tc_shade) {
        for (int l = 0; l < 6; ++l) { k_pack_film<<<1, 256, 0, st>>>(f->sdf_b[l], f->sdf_freq + l * 256, f->sdf_phase + l * 256, h->tc_F + l * 256, h->tc_G + l * 256); ++npack; }
        long long n = 0;
        CU(shade16_pack(f->sdf_W, f->col_W, din, h->shade16_img, st, &n));
        npack += n;
    }
    if (tc_root) {
        // 3xTF32 images: k_iso_init_tc3 (both MLPs) and the k_trace_tc3 fallback
        for (int l = 1; l < 4; ++l) { k_pack_umma_x3<<<cdiv((size_t)4 * 128 * 32, 256), 256, 0, st>>>(f->skin_W[l], 128, h->tc_skin_hid[l - 1], 128, 128, 128, 4); ++npack; }
        k_pack_umma_x3<<<cdiv((size_t)4 * 32 * 32, 256), 256, 0, st>>>(f->skin_W[4], 128, h->tc_skin_out, 25, 32, 128, 4); ++npack;
        for (int l = 1; l < 6; ++l) { k_pack_umma_x3<<<cdiv((size_t)8 * 256 * 32, 256), 256, 0, st>>>(f->sdf_W[l], 256, h->tc_sdf3x[l - 1], 256, 256, 256, 8); ++npack; }
        long long n = 0;
        CU(root_pack_skin_f16(f->skin_W, h->skin16, st, &n));
        npack += n;
    }
    // pose buffers
    const cudaMemcpyKind kind = f->pose_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    CU(cudaMemcpyAsync(h->bone_T, f->bone_transforms, 24 * 16 * 4, kind, st));
    CU(cudaMemcpyAsync(h->verts3, f->smpl_verts, (size_t)h->cfg.n_verts * 12, kind, st));
    k_verts4<<<cdiv(h->cfg.n_verts, 256), 256, 0, st>>>(h->verts3, h->verts4, h->cfg.n_verts); ++npack;
    k_knn_build<<<1, 1024, 8192 * 8, st>>>(h->verts3, h->cfg.n_verts, (float4*)h->knn_sv, (float4*)h->knn_cmin, (float4*)h->knn_cmax); ++npack;
    h->knn.sv = (const float4*)h->knn_sv; h->knn.cmin = (const float4*)h->knn_cmin; h->knn.cmax = (const float4*)h->knn_cmax;
    h->knn.nc = (h->cfg.n_verts + 31) / 32;
    h->pack_launches = npack;
    if (f->smpl_weights) {
        CU(cudaMemcpyAsync(h->smpl_w, f->smpl_weights, (size_t)h->cfg.n_verts * 24 * 4, kind, st));
        h->have_smpl_w = true;
    }
    CU(cudaGetLastError());
    FrameParams& fp = h->fp;
    for (int l = 0; l < 6; ++l) { fp.sdf_Wt[l] = h->sdf_Wt[l]; fp.sdf_W[l] = h->sdf_W[l]; fp.sdf_b[l] = h->sdf_b[l]; }
    fp.sdf_w6 = h->sdf_w6; fp.sdf_b6 = b6; fp.sdf_freq = h->sdf_freq; fp.sdf_phase = h->sdf_phase;
    for (int l = 0; l < 5; ++l) { fp.skin_Wt[l] = h->skin_Wt[l]; fp.skin_b[l] = h->skin_b[l]; }
    fp.col_Wt0 = h->col_Wt0; fp.col_Wt1 = h->col_Wt1; fp.col_Wt2 = h->col_Wt2; fp.col_Wt3a = h->col_Wt3a;
    fp.col_Wt3b = h->col_Wt3b; fp.col_Wt4 = h->col_Wt4; fp.col_W5 = h->col_W5;
    for (int l = 0; l < 6; ++l) fp.col_b[l] = h->col_b[l];
    fp.bone_T = h->bone_T; fp.verts4 = h->verts4; fp.smpl_w = h->smpl_w; fp.n_verts = h->cfg.n_verts;
    for (int k = 0; k < 3; ++k) { fp.trans[k] = f->trans[k]; fp.center[k] = f->center[k]; fp.cam_loc[k] = f->cam_loc[k]; }
    fp.cmin = f->coord_min; fp.cmax = f->coord_max;
    for (int k = 0; k < 16; ++k) fp.pose[k] = f->pose[k];
    fp.beta = f->beta;
    fp.n_steps = h->cfg.n_steps; fp.near_samples = h->cfg.near_samples; fp.far_samples = h->cfg.far_samples;
    fp.cano_view_dirs = h->cfg.cano_view_dirs;
    fp.render_last_pt = h->cfg.render_last_pt ? 1 : 0;
    SdfTC& sd = h->sd;
    sd.Wt0 = h->sdf_Wt[0]; sd.freq = h->sdf_freq; sd.phase = h->sdf_phase; sd.w6 = h->sdf_w6; sd.b6 = b6;
    for (int l = 0; l < 6; ++l) sd.b[l] = h->sdf_b[l];
    for (int l = 0; l < 5; ++l) sd.hid[l] = h->tc_sdf3x[l];
    SkinTC& sk = h->sk;
    sk.Wt0 = h->skin_Wt[0];
    for (int l = 0; l < 5; ++l) sk.b[l] = h->skin_b[l];
    for (int l = 0; l < 3; ++l) sk.hid[l] = h->tc_skin_hid[l];
    sk.out = h->tc_skin_out;
    if (!(fp.cmax > fp.cmin)) return fail(ARAH_EINVAL, "coord_max must exceed coord_min");
    if (h->training) {
        // raw reference-layout copies for the training engine (arah_train.h reads [out][in] weights directly)
        const int d0 = 33 + 256 + L, d3 = d0 + 128;
        const size_t n_sdfW[7] = {256 * 3, 65536, 65536, 65536, 65536, 65536, 256}, n_sdfb[7] = {256, 256, 256, 256, 256, 256, 1};
        const size_t n_skW[5] = {128 * 3, 16384, 16384, 16384, 25 * 128}, n_skb[5] = {128, 128, 128, 128, 25};
        const size_t n_cW[6] = {(size_t)256 * d0, 65536, 128 * 256, (size_t)256 * d3, 65536, 3 * 256}, n_cb[6] = {256, 256, 128, 256, 256, 3};
        size_t tot = 0;
        auto pad = [](size_t n) { return (n + 63) / 64 * 64; };
        for (int i = 0; i < 7; ++i) tot += pad(n_sdfW[i]) + pad(n_sdfb[i]);
        tot += 2 * pad(6 * 256);
        for (int i = 0; i < 5; ++i) tot += pad(n_skW[i]) + pad(n_skb[i]);
        for (int i = 0; i < 6; ++i) tot += pad(n_cW[i]) + pad(n_cb[i]);
        tot += pad((size_t)(L > 0 ? L : 1));
        if (h->raw.ensure(tot * 4) != 0) return fail(ARAH_ENOMEM, "training weight copies");
        float* cur = (float*)h->raw.p;
        cudaError_t ce = cudaSuccess;
        auto cp = [&](const float* src, size_t n) { float* d = cur; cur += pad(n); if (ce == cudaSuccess) ce = cudaMemcpyAsync(d, src, n * 4, cudaMemcpyDeviceToDevice, st); return (const float*)d; };
        arah::train::AllParams& T = h->tp;
        for (int i = 0; i < 7; ++i) { T.sdf.W[i] = cp(f->sdf_W[i], n_sdfW[i]); T.sdf.b[i] = cp(f->sdf_b[i], n_sdfb[i]); }
        T.sdf.freq = cp(f->sdf_freq, 6 * 256); T.sdf.phase = cp(f->sdf_phase, 6 * 256);
        for (int i = 0; i < 5; ++i) { T.skin.W[i] = cp(f->skin_W[i], n_skW[i]); T.skin.b[i] = cp(f->skin_b[i], n_skb[i]); }
        for (int i = 0; i < 6; ++i) { T.col.W[i] = cp(f->col_W[i], n_cW[i]); T.col.b[i] = cp(f->col_b[i], n_cb[i]); }
        T.col.latent = L > 0 ? cp(f->latent, (size_t)L) : nullptr;
        T.col.latent_dim = L;
        if (ce != cudaSuccess) return fail(ARAH_ECUDA, cudaGetErrorString(ce));
        T.bone_T = h->bone_T;
        T.nm.cmin = fp.cmin; T.nm.cmax = fp.cmax;
        for (int k = 0; k < 3; ++k) T.nm.center[k] = fp.center[k];
        h->col_d0 = d0;
    }
    h->train_traced = false;
    h->frame_set = true;
    return ARAH_OK;
}

static int render_device(ArahHandle* h, const float* ray_dirs, const float* near_far, int P, float* rgb, uint8_t* mask,
                         float* points_cam, float* wsum, cudaStream_t st, const float* u_all = nullptr,
                         const float* u_near = nullptr, const float* u_far = nullptr, bool trace_only = false) {
    if (!h->frame_set) return fail(ARAH_ESTATE, "arah_set_frame must be called before arah_render");
    if (P < 0) return fail(ARAH_EINVAL, "P < 0");
    h->launches = 0;
    h->shade_cull_ran = false;
    h->last_P = P;
    h->rendered = true;
    if (P == 0) return ARAH_OK;
    if (!ray_dirs || !near_far || (!trace_only && (!rgb || !mask))) return fail(ARAH_EINVAL, "null buffer");
    CU(cudaSetDevice(h->cfg.device));
    if (ensure_workspace(h, P) != 0) return fail(ARAH_ENOMEM, "workspace allocation failed");
    Work& w = h->w;
    w.P = P; w.ray_dirs = ray_dirs; w.near_far = near_far;
    w.train = u_all ? 1 : 0; w.u_all = u_all; w.u_near = u_near; w.u_far = u_far;
    w.out_rgb = rgb; w.out_mask = mask; w.out_points_cam = points_cam; w.out_wsum = wsum;
    const FrameParams& fp = h->fp;
    const int S = w.S;
    const size_t PS = (size_t)P * S;
    const int nsm = h->n_sms;
    const size_t sm_sdf = tile_smem_bytes(LDA_SDF), sm_skin = tile_smem_bytes(LDA_SKIN), sm_knn = knn_smem_bytes(fp.n_verts);
    auto L = [&]() { h->launches++; };
    const bool prof = h->profile;
    h->profiled = prof;
    CU(cudaMemsetAsync(w.counters, 0, C_COUNT * 4, st));
    w.shade_ctr = C_SHADE;
    w.knn_seed = h->knn_seed;
    w.trace_knn = h->trace_knn;
    Work wk = w;                      // kernels get the phase-clock pointer only while profiling
    if (prof) CU(cudaMemsetAsync(w.phase_clk, 0, 32 * 8, st)); else wk.phase_clk = nullptr;
    if (prof) CU(cudaEventRecord(h->ev[0], st));
    k_trace_begin<<<cdiv(P, 256), 256, 0, st>>>(w); L();
    const unsigned g_knn_rays = grid_min(cdiv(P, 16), (size_t)nsm);        // >= one query per warp; idle blocks exit before staging
    // root_mode 3xTF32: the persistent tensor-core kernels (fp16 split precision); fp32: one FFMA-tile launch per iteration
    const bool tc_root = h->cfg.root_mode == ARAH_ROOT_3XTF32;
    SdfF16Host sh16;
    sh16.Wt0 = h->sdf_Wt[0]; sh16.freq = h->sdf_freq; sh16.phase = h->sdf_phase; sh16.w6 = h->sdf_w6; sh16.b6 = h->d_b6;
    for (int l = 0; l < 6; ++l) sh16.b[l] = h->sdf_b[l];
    if (tc_root && h->trace_persist) {
        long long n = 0;
        CU(root_trace_persist(fp, sh16, h->sdf16, h->knn, wk, nsm, st, &n));
        h->launches += n;
    } else {
        // fp32 mode, or the vertex index does not fit next to the persistent kernel's weight ring (n_verts > ~7000; ARAH_TRACE_PERSIST=0
        // forces this path in the tests): one 1-NN + one SDF launch per sphere-tracing step
        for (int it = 0; it < TRACE_ITERS; ++it) {
            k_knn_rays<<<g_knn_rays, 512, sm_knn, st>>>(fp, h->knn, w, it); L();
            if (tc_root) k_trace_tc3<<<grid_min(cdiv(P, UM), (size_t)nsm), TC3_THREADS, trace_tc3_smem_bytes(), st>>>(fp, h->sd, wk, it);
            else k_trace_iter<<<grid_min(cdiv(P, TM), (size_t)2 * nsm), 256, sm_sdf, st>>>(fp, w, it);
            L();
        }
    }
    if (prof) CU(cudaEventRecord(h->ev[1], st));
    k_iso_prepare<<<cdiv(P, 256), 256, 0, st>>>(w); L();
    if (tc_root) {
        k_iso_init_tc3<<<grid_min(cdiv(P, ISO_TC_PTS), (size_t)nsm), TC3_THREADS, trace_tc3_smem_bytes(), st>>>(fp, h->sd, h->sk, w); L();
        long long n = 0;
        CU(root_iso_persist(fp, sh16, h->sdf16, h->skin_Wt[0], h->skin_b, h->skin16, wk, nsm, st, &n));
        h->launches += n;
    } else {
        k_iso_init<<<grid_min(cdiv(P, TM / 4), (size_t)nsm), 256, sm_sdf, st>>>(fp, w); L();
        for (int it = 0; it < BROYDEN_ITERS; ++it) { k_iso_iter<<<grid_min(cdiv(P, TM), (size_t)2 * nsm), 256, sm_sdf, st>>>(fp, w, it); L(); }
    }
    if (prof) CU(cudaEventRecord(h->ev[2], st));
    k_trace_finish<<<cdiv(P, 128), 128, 0, st>>>(fp, w); L();
    const unsigned g_knn_s = grid_min(cdiv(PS, 16), (size_t)nsm);
    w.corr_seed = tc_root ? reinterpret_cast<CorrSeed*>(w.corr_state) : nullptr;
    wk.corr_seed = w.corr_seed;
    k_knn_samples<<<g_knn_s, 512, sm_knn + knn_quarter_smem_bytes(fp.n_verts), st>>>(fp, h->knn, w); L();
    if (tc_root) {
        long long n = 0;
        CU(root_corr_persist(fp, h->skin_Wt[0], h->skin_b, h->skin16, wk, nsm, st, &n));
        h->launches += n;
    } else {
        const unsigned g_smp_tiles = grid_min(cdiv(PS, TM), (size_t)2 * nsm);
        for (int it = -1; it < BROYDEN_ITERS; ++it) { k_corr_step<<<g_smp_tiles, 256, sm_skin, st>>>(fp, w, it); L(); }
    }
    if (prof) CU(cudaEventRecord(h->ev[3], st));
    if (trace_only) {                                  // (training: the shading runs in arah_train.h; stage events stay well defined)
        if (prof) { CU(cudaEventRecord(h->ev[4], st)); CU(cudaEventRecord(h->ev[5], st)); }
        CU(cudaGetLastError());
        return ARAH_OK;
    }
    if (h->cfg.shade_mode == ARAH_SHADE_TF32) {
        // tensor-core shading (11-bit operands: fp16 images; round 1 used TF32, hence the mode's name)
        const unsigned g = grid_min(cdiv(PS, UM), (size_t)nsm);
        h->shade_cull_ran = h->shade_cull != 0;
        wk.shade_keep_sdf = 1;
        Shade16Host s16h{};
        s16h.sdf_Wt0 = h->sdf_Wt[0]; s16h.sdf_W0 = h->sdf_W[0]; s16h.sdf_F = h->tc_F; s16h.sdf_G = h->tc_G; s16h.sdf_scale = h->sdf16.scale;
        s16h.sdf_fwd_hi = h->sdf16.hi; s16h.sdf_w6 = h->sdf_w6; s16h.sdf_b6 = h->d_b6; s16h.col_W5 = h->col_W5;
        for (int l = 0; l < 6; ++l) s16h.col_b[l] = h->col_b[l];
        // 1. the SDF value of every converged sample (what compositing turns into sigma) in one fp16 single-pass sweep; the gradient +
        //    colour pass below then only supplies colours, with or without the cull: both settings composite identical inputs
        long long n = 0;
        CU(root_sdf_fwd16(fp, sh16, h->sdf16, wk, nsm, st, &n));
        Work w2 = wk;
        if (h->shade_cull) {
            // 2. exact alpha cull: samples whose alpha is exactly 0 cannot influence any output bit (arah_kernels.cuh)
            k_alpha_cull<<<cdiv(P, COMP_WARPS), 32 * COMP_WARPS, 0, st>>>(fp, w, w.listA); L();
            w2.shade_list = w.listA; w2.shade_ctr = C_SHADE2;
        }
        // 3. SDF gradient + colour of the survivors
        CU(shade16_launch(fp, s16h, h->shade16_img, w2, g, st, &n));
        h->launches += n;
    }
    else { k_shade<<<grid_min(cdiv(PS, TM), (size_t)nsm), 256, shade_smem_bytes(), st>>>(fp, w); L(); }
    if (prof) CU(cudaEventRecord(h->ev[4], st));
    k_composite<<<cdiv(P, COMP_WARPS), 32 * COMP_WARPS, 0, st>>>(fp, w); L();
    if (prof) CU(cudaEventRecord(h->ev[5], st));
    CU(cudaGetLastError());
    return ARAH_OK;
}

extern "C" int arah_render(ArahHandle* h, const float* ray_dirs, const float* near_far, int32_t P, float* rgb, uint8_t* mask,
                           float* points_cam, float* weights_sum, void* stream) {
    if (!h) return fail(ARAH_EINVAL, "null handle");
    return render_device(h, ray_dirs, near_far, P, rgb, mask, points_cam, weights_sum, (cudaStream_t)stream);
}

extern "C" int arah_render_host(ArahHandle* h, const float* ray_dirs, const float* near_far, int32_t P, float* rgb, uint8_t* mask,
                                float* points_cam, void* stream) {
    if (!h) return fail(ARAH_EINVAL, "null handle");
    if (P < 0) return fail(ARAH_EINVAL, "P < 0");
    if (P == 0) { h->launches = 0; h->last_P = 0; return ARAH_OK; }
    if (!ray_dirs || !near_far || !rgb || !mask) return fail(ARAH_EINVAL, "null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)P;
    if (h->io_in.ensure(n * 20 + 256) != 0 || h->io_out.ensure(n * 28 + 512) != 0) return fail(ARAH_ENOMEM, "io buffers");
    float* d_dirs = (float*)h->io_in.p;
    float* d_nf = d_dirs + n * 3;
    float* d_rgb = (float*)h->io_out.p;
    float* d_pc = d_rgb + n * 3;
    uint8_t* d_mask = (uint8_t*)(d_pc + n * 3);
    CU(cudaMemcpyAsync(d_dirs, ray_dirs, n * 12, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_nf, near_far, n * 8, cudaMemcpyHostToDevice, st));
    int rc = render_device(h, d_dirs, d_nf, P, d_rgb, d_mask, d_pc, nullptr, st);
    if (rc != ARAH_OK) return rc;
    CU(cudaMemcpyAsync(rgb, d_rgb, n * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(mask, d_mask, n, cudaMemcpyDeviceToHost, st));
    if (points_cam) CU(cudaMemcpyAsync(points_cam, d_pc, n * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return ARAH_OK;
}

extern "C" int arah_get_trace(ArahHandle* h, float* points_hat_norm, uint8_t* network_body_mask, float* dists, float* sampled_pts,
                              float* sampled_dists, float* sampled_transforms, uint8_t* sampler_converge_mask, void* stream) {
    if (!h) return fail(ARAH_EINVAL, "null handle");
    if (!h->rendered) return fail(ARAH_ESTATE, "arah_get_trace needs a preceding arah_render");
    cudaStream_t st = (cudaStream_t)stream;
    const int P = h->last_P;
    if (P == 0) return ARAH_OK;
    const Work& w = h->w;
    if (points_hat_norm) CU(cudaMemcpyAsync(points_hat_norm, w.ray_pnorm, (size_t)P * 12, cudaMemcpyDeviceToDevice, st));
    if (network_body_mask) CU(cudaMemcpyAsync(network_body_mask, w.ray_conv, (size_t)P, cudaMemcpyDeviceToDevice, st));
    if (dists) CU(cudaMemcpyAsync(dists, w.ray_dist, (size_t)P * 4, cudaMemcpyDeviceToDevice, st));
    if (sampled_pts || sampled_dists || sampled_transforms || sampler_converge_mask) {
        const size_t PS = (size_t)P * w.S;
        k_export_samples<<<cdiv(PS, 256), 256, 0, st>>>(w, h->cfg.near_samples + 1 + h->cfg.far_samples, sampled_pts, sampled_dists,
                                                        sampled_transforms, sampler_converge_mask);
        CU(cudaGetLastError());
    }
    return ARAH_OK;
}

extern "C" int arah_get_stats(ArahHandle* h, ArahStats* s, void* stream) {
    if (!h || !s) return fail(ARAH_EINVAL, "null argument");
    memset(s, 0, sizeof(*s));
    s->rays = h->last_P;
    s->kernel_launches = h->launches;
    s->pack_launches = h->pack_launches;
    if (!h->rendered || h->last_P == 0) return ARAH_OK;
    int c[C_COUNT];
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    CU(cudaMemcpy(c, h->w.counters, sizeof(c), cudaMemcpyDeviceToHost));
    s->trace_sdf_evals = c[C_STAT_TRACE_EVALS];
    s->iso_rays = c[C_ISO];
    s->iso_g_evals = c[C_STAT_ISO_EVALS];
    s->on_samples = c[C_ON];
    s->corr_skin_evals = c[C_STAT_CORR_EVALS];
    s->shaded_samples = c[C_SHADE];
    s->culled_samples = (h->shade_cull_ran) ? (int64_t)c[C_SHADE] - c[C_SHADE2] : 0;
    s->hit_rays = c[C_STAT_HIT_RAYS];
    s->vol_rays = c[C_STAT_VOL_RAYS];
    if (h->profiled) {
        float ms[5] = {0, 0, 0, 0, 0}, tot = 0;
        for (int i = 0; i < 5; ++i) CU(cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
        CU(cudaEventElapsedTime(&tot, h->ev[0], h->ev[5]));
        s->ms_trace = ms[0]; s->ms_iso = ms[1]; s->ms_sample_corr = ms[2]; s->ms_shade = ms[3]; s->ms_composite = ms[4]; s->ms_total = tot;
    }
    return ARAH_OK;
}

extern "C" int arah_debug_phase_clocks(ArahHandle* h, uint64_t* out16, void* stream) {
    if (!h || !out16) return fail(ARAH_EINVAL, "null argument");
    memset(out16, 0, 32 * 8);
    if (!h->rendered || !h->profiled || h->last_P == 0) return ARAH_OK;
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    CU(cudaMemcpy(out16, h->w.phase_clk, 32 * 8, cudaMemcpyDeviceToHost));
    return ARAH_OK;
}

extern "C" int arah_eval_sdf(ArahHandle* h, const float* xn, int32_t n, float* sdf, float* grad, float* feat, void* stream) {
    if (!h || !h->frame_set) return fail(ARAH_ESTATE, "arah_set_frame first");
    if (n <= 0) return ARAH_OK;
    if (!xn || !sdf) return fail(ARAH_EINVAL, "null buffer");
    const unsigned g = grid_min(cdiv(n, TM), (size_t)h->n_sms);
    k_eval_sdf<<<g, 256, tile_smem_bytes(LDA_SDF), (cudaStream_t)stream>>>(h->fp, xn, n, sdf, grad, feat, (float*)h->scratch.p);
    CU(cudaGetLastError());
    return ARAH_OK;
}
extern "C" int arah_eval_skin(ArahHandle* h, const float* x_hat, int32_t n, float* weights, float* x_bar, void* stream) {
    if (!h || !h->frame_set) return fail(ARAH_ESTATE, "arah_set_frame first");
    if (n <= 0) return ARAH_OK;
    if (!x_hat || !weights || !x_bar) return fail(ARAH_EINVAL, "null buffer");
    const unsigned g = grid_min(cdiv(n, TM), (size_t)2 * h->n_sms);
    k_eval_skin<<<g, 256, tile_smem_bytes(LDA_SKIN), (cudaStream_t)stream>>>(h->fp, x_hat, n, weights, x_bar);
    CU(cudaGetLastError());
    return ARAH_OK;
}

extern "C" int arah_debug_knn(ArahHandle* h, const float* pts, int32_t n, int32_t* idx, void* stream) {
    if (!h || !h->frame_set) return fail(ARAH_ESTATE, "arah_set_frame first");
    if (n <= 0) return ARAH_OK;
    if (!pts || !idx) return fail(ARAH_EINVAL, "null buffer");
    k_knn_points<<<grid_min(cdiv(n, 16), (size_t)h->n_sms), 512, knn_smem_bytes(h->fp.n_verts), (cudaStream_t)stream>>>(h->knn, pts, n, idx);
    CU(cudaGetLastError());
    return ARAH_OK;
}

// ------------------------------------------------------------------------------------------------ canonical SDF lattice (row f1)
__global__ void k_grid_points(int N, float voxel, long long i0, int n, float* __restrict__ xn) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long i = i0 + t;
    const int iz = (int)(i % N), iy = (int)((i / N) % N), ix = (int)(i / ((long long)N * N));
    xn[3 * t] = __fadd_rn(__fmul_rn((float)ix, voxel), -1.0f);
    xn[3 * t + 1] = __fadd_rn(__fmul_rn((float)iy, voxel), -1.0f);
    xn[3 * t + 2] = __fadd_rn(__fmul_rn((float)iz, voxel), -1.0f);
}

extern "C" int arah_sdf_grid(ArahHandle* h, int32_t N, float* sdf, void* stream) {
    if (!h || !h->frame_set) return fail(ARAH_ESTATE, "arah_set_frame first");
    if (N < 2 || N > 1024) return fail(ARAH_EINVAL, "lattice side must be in [2, 1024]");
    if (!sdf) return fail(ARAH_EINVAL, "null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(h->cfg.device));
    const float voxel = (float)(2.0 / (double)(N - 1));
    const long long n = (long long)N * N * N;
    if (h->cfg.root_mode == ARAH_ROOT_3XTF32) {
        SdfF16Host sh16;
        sh16.Wt0 = h->sdf_Wt[0]; sh16.freq = h->sdf_freq; sh16.phase = h->sdf_phase; sh16.w6 = h->sdf_w6; sh16.b6 = h->d_b6;
        for (int l = 0; l < 6; ++l) sh16.b[l] = h->sdf_b[l];
        CU(root_sdf_grid16(sh16, h->sdf16, N, voxel, n, sdf, h->n_sms, st));
    } else {
        // fp32 FFMA tiles: lattice points are staged in chunks of 64^3 (the reference's own max_batch, sdf_meshing.py:14)
        const int chunk = 64 * 64 * 64;
        if (h->io_in.ensure((size_t)chunk * 3 * 4) != 0) return fail(ARAH_ENOMEM, "lattice staging allocation failed");
        float* xn = (float*)h->io_in.p;
        for (long long i0 = 0; i0 < n; i0 += chunk) {
            const int m = (int)((n - i0 < chunk) ? (n - i0) : chunk);
            k_grid_points<<<cdiv(m, 256), 256, 0, st>>>(N, voxel, i0, m, xn);
            k_eval_sdf<<<grid_min(cdiv(m, TM), (size_t)h->n_sms), 256, tile_smem_bytes(LDA_SDF), st>>>(h->fp, xn, m, sdf + i0, nullptr, nullptr, (float*)h->scratch.p);
        }
    }
    CU(cudaGetLastError());
    return ARAH_OK;
}


extern "C" int arah_sdf_grid_banded(ArahHandle* h, int32_t N, float level, float eps, float* sdf, int32_t* stats, void* stream) {
    if (!h || !h->frame_set) return fail(ARAH_ESTATE, "arah_set_frame first");
    if (N < 2 || N > 1024) return fail(ARAH_EINVAL, "lattice side must be in [2, 1024]");
    if (!sdf || !stats) return fail(ARAH_EINVAL, "null buffer");
    if (!(eps >= 0.f)) return fail(ARAH_EINVAL, "eps must be >= 0");
    if (h->cfg.root_mode != ARAH_ROOT_3XTF32) return fail(ARAH_ESTATE, "the banded lattice needs root_mode 3xTF32 (fp16 weight images); use arah_sdf_grid");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)N * N * N;
    if (h->io_in.ensure(align_up(n, 256) + n * 4) != 0) return fail(ARAH_ENOMEM, "lattice flag / list allocation failed");
    uint8_t* flag = (uint8_t*)h->io_in.p;
    int* list = (int*)(flag + align_up(n, 256));
    SdfF16Host sh16;
    sh16.Wt0 = h->sdf_Wt[0]; sh16.freq = h->sdf_freq; sh16.phase = h->sdf_phase; sh16.w6 = h->sdf_w6; sh16.b6 = h->d_b6;
    for (int l = 0; l < 6; ++l) sh16.b[l] = h->sdf_b[l];
    CU(root_sdf_grid_banded(sh16, h->sdf16, N, (float)(2.0 / (double)(N - 1)), level, eps, sdf, flag, list, stats, h->n_sms, st, nullptr));
    return ARAH_OK;
}


// ================================================================================================ training (arah_train.h)
using arah::train::CudaBK;

static arah::train::AllGrads to_grads(const ArahTrainGrads* g) {
    arah::train::AllGrads G{};
    if (!g) return G;
    for (int i = 0; i < 7; ++i) { G.sdf.W[i] = g->sdf_W[i]; G.sdf.b[i] = g->sdf_b[i]; }
    G.sdf.freq = g->sdf_freq; G.sdf.phase = g->sdf_phase;
    for (int i = 0; i < 5; ++i) { G.skin.W[i] = g->skin_W[i]; G.skin.b[i] = g->skin_b[i]; }
    for (int i = 0; i < 6; ++i) { G.col.W[i] = g->col_W[i]; G.col.b[i] = g->col_b[i]; }
    G.col.latent = g->latent; G.beta = g->beta;
    return G;
}
static int train_ready(ArahHandle* h) {
    if (!h) return fail(ARAH_EINVAL, "null handle");
    if (!h->training) return fail(ARAH_ESTATE, "call arah_set_training(h, 1) before arah_set_frame");
    if (!h->frame_set || !h->raw.p) return fail(ARAH_ESTATE, "arah_set_frame must follow arah_set_training(h, 1)");
    if (!h->sess) h->sess = new arah::train::Session<CudaBK>();
    CudaBK::precision() = (h->train_precision == ARAH_TRAIN_FP32) ? 1 : (h->train_precision == ARAH_TRAIN_TF32 ? 2 : 0);
    return ARAH_OK;
}
static int train_done(ArahHandle* h, int rc, const char* what) {
    h->launches += CudaBK::launches(); CudaBK::launches() = 0;
    if (rc == -1) return fail(ARAH_ENOMEM, std::string(what) + ": device allocation failed");
    if (rc == -2) return fail(ARAH_EINVAL, std::string(what) + ": bad slot");
    CU(cudaGetLastError());
    return ARAH_OK;
}

extern "C" int arah_set_training(ArahHandle* h, int32_t enable) {
    if (!h) return fail(ARAH_EINVAL, "null handle");
    h->training = enable != 0;
    if (enable == ARAH_TRAIN_3XTF32 || enable == ARAH_TRAIN_FP32 || enable == ARAH_TRAIN_TF32) h->train_precision = enable;
    if (!h->training && h->sess) { h->sess->release(); delete h->sess; h->sess = nullptr; h->raw.release(); }
    return ARAH_OK;
}

extern "C" int arah_train_trace(ArahHandle* h, const float* ray_dirs, const float* near_far, int32_t P, const float* u_all,
                                const float* u_near, const float* u_far, void* stream) {
    if (!h) return fail(ARAH_EINVAL, "null handle");
    if (!u_all || !u_near || (h->cfg.far_samples > 0 && !u_far)) return fail(ARAH_EINVAL, "training-mode tracing needs the three uniform draws");
    const int rc = render_device(h, ray_dirs, near_far, P, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream, u_all, u_near, u_far, true);
    if (rc == ARAH_OK) h->train_traced = true;
    return rc;
}

extern "C" int arah_train_shade_forward(ArahHandle* h, const float* view_dirs, const float* view_dirs_orig, int32_t ray_augm,
                                        int32_t train_skinning_net, float* rgb, float* weights_sum, void* stream) {
    int rc = train_ready(h);
    if (rc != ARAH_OK) return rc;
    if (!h->train_traced) return fail(ARAH_ESTATE, "arah_train_shade_forward needs a preceding arah_train_trace");
    if (!view_dirs || !rgb || !weights_sum) return fail(ARAH_EINVAL, "null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaSetDevice(h->cfg.device));
    const Work& w = h->w;
    int M = 0;
    CU(cudaMemcpyAsync(&M, w.counters + C_SHADE, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    arah::train::ShadeGeom g{};
    g.P = h->last_P; g.S = w.S; g.cano_view_dirs = h->cfg.cano_view_dirs; g.ray_augm = ray_augm ? 1 : 0;
    g.list = w.shade_list; g.smp_xn = w.smp_xn; g.smp_T = w.smp_T; g.z_vals = w.z_vals; g.smp_conv = w.smp_conv;
    g.view = view_dirs; g.view_orig = view_dirs_orig ? view_dirs_orig : view_dirs;
    g.sdf_scale = h->fp.cmax - h->fp.cmin; g.beta_raw = h->fp.beta;
    CudaBK::launches() = 0;
    rc = h->sess->shade_forward(h->tp, g, M, train_skinning_net != 0, rgb, weights_sum, st);
    return train_done(h, rc, "arah_train_shade_forward");
}

extern "C" int arah_train_shade_backward(ArahHandle* h, const float* g_rgb, const float* g_weights_sum, const ArahTrainGrads* grads, void* stream) {
    int rc = train_ready(h);
    if (rc != ARAH_OK) return rc;
    if (!g_rgb || !grads) return fail(ARAH_EINVAL, "null buffer");
    CU(cudaSetDevice(h->cfg.device));
    CudaBK::launches() = 0;
    rc = h->sess->shade_backward(h->tp, to_grads(grads), g_rgb, g_weights_sum, (cudaStream_t)stream);
    return train_done(h, rc, "arah_train_shade_backward");
}

extern "C" int arah_train_sdf_forward(ArahHandle* h, int32_t slot, const float* points, int32_t n, int32_t with_grad, float* sdf,
                                      float* grad, void* stream) {
    int rc = train_ready(h);
    if (rc != ARAH_OK) return rc;
    if (n < 0 || (n > 0 && (!points || !sdf))) return fail(ARAH_EINVAL, "bad arguments");
    CU(cudaSetDevice(h->cfg.device));
    CudaBK::launches() = 0;
    rc = h->sess->sdf_forward(h->tp, slot, points, n, with_grad != 0, sdf, grad, (cudaStream_t)stream);
    return train_done(h, rc, "arah_train_sdf_forward");
}

extern "C" int arah_train_sdf_backward(ArahHandle* h, int32_t slot, const float* g_sdf, const float* g_grad, const ArahTrainGrads* grads, void* stream) {
    int rc = train_ready(h);
    if (rc != ARAH_OK) return rc;
    if (!grads) return fail(ARAH_EINVAL, "null gradient table");
    CU(cudaSetDevice(h->cfg.device));
    CudaBK::launches() = 0;
    rc = h->sess->sdf_backward(h->tp, to_grads(grads), slot, g_sdf, g_grad, (cudaStream_t)stream);
    return train_done(h, rc, "arah_train_sdf_backward");
}

extern "C" int arah_train_skin_forward(ArahHandle* h, const float* points, int32_t n, float* weights, void* stream) {
    int rc = train_ready(h);
    if (rc != ARAH_OK) return rc;
    if (n < 0 || (n > 0 && (!points || !weights))) return fail(ARAH_EINVAL, "bad arguments");
    CU(cudaSetDevice(h->cfg.device));
    CudaBK::launches() = 0;
    rc = h->sess->skin_forward(h->tp, points, n, weights, (cudaStream_t)stream);
    return train_done(h, rc, "arah_train_skin_forward");
}

extern "C" int arah_train_skin_backward(ArahHandle* h, const float* g_weights, const ArahTrainGrads* grads, void* stream) {
    int rc = train_ready(h);
    if (rc != ARAH_OK) return rc;
    if (!g_weights || !grads) return fail(ARAH_EINVAL, "null buffer");
    CU(cudaSetDevice(h->cfg.device));
    CudaBK::launches() = 0;
    rc = h->sess->skin_backward(h->tp, to_grads(grads), g_weights, (cudaStream_t)stream);
    return train_done(h, rc, "arah_train_skin_backward");
}

extern "C" int arah_debug_train_gemm(int32_t M, int32_t N, int32_t K, const float* A, int64_t sa_i, int64_t sa_k, const float* B, int64_t sb_k,
                                     int64_t sb_j, float* C, int32_t ldc, const float* bias, int32_t accumulate, int32_t mode, void* stream) {
    if (!A || !B || !C || M < 0 || N < 0 || K <= 0) return fail(ARAH_EINVAL, "bad arguments");
    CudaBK::precision() = (mode == ARAH_TRAIN_FP32) ? 1 : (mode == ARAH_TRAIN_TF32 ? 2 : 0);
    CudaBK::gemm(M, N, K, A, (long)sa_i, (long)sa_k, B, (long)sb_k, (long)sb_j, C, ldc, bias, accumulate != 0, (cudaStream_t)stream);
    CudaBK::launches() = 0;
    CU(cudaGetLastError());
    return ARAH_OK;
}
