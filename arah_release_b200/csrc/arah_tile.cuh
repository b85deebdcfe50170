// arah_tile.cuh — CTA-tile MLP machinery for sm_100a.
//
// One CTA = 256 threads = 8 warps evaluates an MLP for a tile of TM = 64 rows (points, or value+tangent rows).
// Warp w owns rows 8w..8w+7 of the activation tile A (row-major in shared memory) for the WHOLE network:
// it reads only its rows and overwrites them in place with the next layer's activations, so layers chain with
// __syncwarp() only.  The one thing the 8 warps share is the weight stream: every layer's [K][N] matrix
// (pre-transposed, K-major rows of N floats, zero padded) is staged chunk-by-chunk (KC rows = 16 KB for N=256)
// from L2 into a 2-deep shared-memory ring by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx),
// issued by one thread; consumers wait on the mbarrier phase, compute 8x8 register tiles with fp32 FFMA
// (k-sequential accumulation: the order the CPU oracle uses), and a CTA barrier recycles the slot.
//
// Thread (warp w, lane l) accumulates rows 8w+r (r<8) x columns {4l..4l+3} U {128+4l..128+4l+3} (N=256),
// {4l..4l+3} (N=128) or {l} (N=32): float4 weight reads and float4 activation stores are bank-conflict free,
// activation reads are warp-broadcast.
#pragma once
#include <cuda_runtime.h>
#include "arah_math.cuh"

namespace arah {

constexpr int TM = 64;          // rows per tile
constexpr int NTHREADS = 256;   // threads per CTA
constexpr int KC = 16;          // weight rows per staged chunk
constexpr int SDF_H = 256;
constexpr int SKIN_H = 128;
constexpr int COL_H = 256;
constexpr int COL_IN = 289;     // feat 256 + xn 3 + PE(view) 27 + normal 3   (latent folded into the biases)
constexpr int COL_IN_PAD = 304; // multiple of KC
constexpr int WBUF_FLOATS = 2 * KC * 256;

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    int spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1 << 22)) __trap();     // never hang the GPU box: a lost copy becomes a launch failure
    }
}

// weight-stream pipeline state: 2 mbarriers + per-buffer phase parity (tracked identically by every thread)
struct WPipe {
    float* buf;         // [2][KC*256]
    uint64_t* bars;     // [2]
    uint32_t par;       // bit b = parity to wait for on buffer b
};
__device__ __forceinline__ void wpipe_init(WPipe& p, float* buf, uint64_t* bars) {
    p.buf = buf; p.bars = bars; p.par = 0;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ GEMM core
template <int N> struct ColsPerThread { static constexpr int value = N / 32; };

template <int N>
__device__ __forceinline__ void compute_chunk(float (&acc)[8][N / 32], const float* __restrict__ Arow, const int lda,
                                              const float* __restrict__ wb, const int lane) {
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
        float4 a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float4*>(Arow + r * lda + kk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if constexpr (N == 256) {
                const float4 w0 = *reinterpret_cast<const float4*>(wb + (kk + j) * 256 + 4 * lane);
                const float4 w1 = *reinterpret_cast<const float4*>(wb + (kk + j) * 256 + 128 + 4 * lane);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float av = (j == 0) ? a[r].x : (j == 1) ? a[r].y : (j == 2) ? a[r].z : a[r].w;
                    acc[r][0] = fmaf(av, w0.x, acc[r][0]); acc[r][1] = fmaf(av, w0.y, acc[r][1]);
                    acc[r][2] = fmaf(av, w0.z, acc[r][2]); acc[r][3] = fmaf(av, w0.w, acc[r][3]);
                    acc[r][4] = fmaf(av, w1.x, acc[r][4]); acc[r][5] = fmaf(av, w1.y, acc[r][5]);
                    acc[r][6] = fmaf(av, w1.z, acc[r][6]); acc[r][7] = fmaf(av, w1.w, acc[r][7]);
                }
            } else if constexpr (N == 128) {
                const float4 w0 = *reinterpret_cast<const float4*>(wb + (kk + j) * 128 + 4 * lane);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float av = (j == 0) ? a[r].x : (j == 1) ? a[r].y : (j == 2) ? a[r].z : a[r].w;
                    acc[r][0] = fmaf(av, w0.x, acc[r][0]); acc[r][1] = fmaf(av, w0.y, acc[r][1]);
                    acc[r][2] = fmaf(av, w0.z, acc[r][2]); acc[r][3] = fmaf(av, w0.w, acc[r][3]);
                }
            } else {
                const float w0 = wb[(kk + j) * 32 + lane];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float av = (j == 0) ? a[r].x : (j == 1) ? a[r].y : (j == 2) ? a[r].z : a[r].w;
                    acc[r][0] = fmaf(av, w0, acc[r][0]);
                }
            }
        }
    }
}

// acc += A[8 rows of this warp][0..K) * Wg[K][N].  K % KC == 0.  ACCUMULATE=false zeroes acc first.
// Must be called by all 256 threads (contains CTA barriers); on entry nobody may still be reading the ring.
template <int N, bool ACCUMULATE = false>
__device__ __forceinline__ void tile_gemm(float (&acc)[8][N / 32], const float* __restrict__ Arow, const int lda,
                                          const int K, const float* __restrict__ Wg, WPipe& wp) {
    constexpr uint32_t CHUNK_BYTES = KC * N * 4;
    const int lane = threadIdx.x & 31;
    const int nc = K / KC;
    if constexpr (!ACCUMULATE) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < N / 32; ++c) acc[r][c] = 0.0f;
    }
    if (threadIdx.x == 0) {
        mbar_expect_tx(&wp.bars[0], CHUNK_BYTES);
        bulk_g2s(wp.buf, Wg, CHUNK_BYTES, &wp.bars[0]);
    }
    for (int c = 0; c < nc; ++c) {
        const int b = c & 1;
        if (c + 1 < nc && threadIdx.x == 0) {
            mbar_expect_tx(&wp.bars[b ^ 1], CHUNK_BYTES);
            bulk_g2s(wp.buf + (b ^ 1) * KC * 256, Wg + (size_t)(c + 1) * KC * N, CHUNK_BYTES, &wp.bars[b ^ 1]);
        }
        mbar_wait(&wp.bars[b], (wp.par >> b) & 1u);
        wp.par ^= (1u << b);
        compute_chunk<N>(acc, Arow + c * KC, lda, wp.buf + b * KC * 256, lane);
        __syncthreads();                       // slot b is free again (and refilled two chunks later)
    }
}

// column index of accumulator slot c for this lane
template <int N> __device__ __forceinline__ int acc_col(int c, int lane) {
    if constexpr (N == 256) return (c < 4) ? (4 * lane + c) : (128 + 4 * lane + (c - 4));
    else if constexpr (N == 128) return 4 * lane + c;
    else return lane;
}
template <int N> __device__ __forceinline__ void load_cols(float (&v)[N / 32], const float* __restrict__ g, int lane) {
    if constexpr (N == 256) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(g + 4 * lane));
        const float4 b = __ldg(reinterpret_cast<const float4*>(g + 128 + 4 * lane));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if constexpr (N == 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(g + 4 * lane));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
        v[0] = __ldg(g + lane);
    }
}
// same, but through the coherent path: for scratch that THIS kernel wrote earlier (ld.global.nc would be stale)
template <int N> __device__ __forceinline__ void load_cols_rw(float (&v)[N / 32], const float* g, int lane) {
    if constexpr (N == 256) {
        const float4 a = *reinterpret_cast<const float4*>(g + 4 * lane);
        const float4 b = *reinterpret_cast<const float4*>(g + 128 + 4 * lane);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else if constexpr (N == 128) {
        const float4 a = *reinterpret_cast<const float4*>(g + 4 * lane);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
        v[0] = g[lane];
    }
}
template <int N> __device__ __forceinline__ void store_row(float* __restrict__ row, const float (&v)[N / 32], int lane) {
    if constexpr (N == 256) {
        *reinterpret_cast<float4*>(row + 4 * lane) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(row + 128 + 4 * lane) = make_float4(v[4], v[5], v[6], v[7]);
    } else if constexpr (N == 128) {
        *reinterpret_cast<float4*>(row + 4 * lane) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        row[lane] = v[0];
    }
}

// y[r] = dot(A[row r][0..K), w[0..K)) for the warp's 8 rows, K multiple of 32; result valid in all lanes.
__device__ __forceinline__ void warp_rows_dot(const float* __restrict__ Arow, int lda, int K, const float* __restrict__ w,
                                              float (&y)[8], int lane) {
#pragma unroll
    for (int r = 0; r < 8; ++r) y[r] = 0.0f;
    for (int k = lane; k < K; k += 32) {
        const float wk = __ldg(w + k);
#pragma unroll
        for (int r = 0; r < 8; ++r) y[r] = fmaf(Arow[r * lda + k], wk, y[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) y[r] += __shfl_xor_sync(0xffffffffu, y[r], o);
    }
}

// ------------------------------------------------------------------------------------------------ frame constants
struct FrameParams {
    // SDF FiLM-SIREN (hyperlayers.py:391-415): Wt = [in][out] (forward), W = [out][in] (backward, reference layout)
    const float* sdf_Wt[6];      // l=0: [3][256]; l=1..5: [256][256]
    const float* sdf_W[6];       // l=0: [256][3]; l=1..5: [256][256]
    const float* sdf_b[6];
    const float* sdf_w6;         // [256]
    const float* sdf_b6;         // [1] device scalar
    const float* sdf_freq;       // [6][256]
    const float* sdf_phase;      // [6][256]
    // skinning MLP (metaavatar/models/decoder.py:201-233), weight-norm folded
    const float* skin_Wt[5];     // l=0: [3][128]; l=1..3: [128][128]; l=4: [128][32] (25 padded)
    const float* skin_b[5];      // b4 padded to 32
    // colour MLP (metaavatar_render/models/decoder.py:69-124), weight-norm + latent folded, inputs permuted to
    // [feat 256 | xn 3 | PE(view) 27 | normal 3 | 0-pad] = 304
    const float* col_Wt0;        // [304][256]
    const float* col_Wt1;        // [256][256]
    const float* col_Wt2;        // [256][128]
    const float* col_Wt3a;       // [304][256]  (skip: the network-input part)
    const float* col_Wt3b;       // [128][256]  (skip: the lin2-output part)
    const float* col_Wt4;        // [256][256]
    const float* col_W5;         // [3][256]
    const float* col_b[6];       // b0, b3 include W[:, latent] @ latent
    // body
    const float* bone_T;         // [24][16]
    const float4* verts4;        // [n_verts] (x, y, z, 0) posed + trans
    const float* smpl_w;         // [n_verts][24]
    int n_verts;
    float trans[3], cmin, cmax, center[3], cam_loc[3], pose[16], beta;
    int n_steps, near_samples, far_samples, cano_view_dirs;
    int render_last_pt;          // last interval of a ray is 1e10 instead of 1 / n_steps (implicit_differentiable_renderer.py:380-381)
};

__device__ __forceinline__ void normalize3(const FrameParams& fp, const float* p, float* q) {
#pragma unroll
    for (int k = 0; k < 3; ++k) q[k] = normalize1(p[k], fp.center[k], fp.cmin, fp.cmax);
}
__device__ __forceinline__ void unnormalize3(const FrameParams& fp, const float* p, float* q) {
#pragma unroll
    for (int k = 0; k < 3; ++k) q[k] = unnormalize1(p[k], fp.center[k], fp.cmin, fp.cmax);
}

// ------------------------------------------------------------------------------------------------ SDF tiles
// Row modes.  PLAIN: every row is a point.  DUAL: rows 4q+0 = value of point q, rows 4q+1..3 = tangents d/dx_hat_k
// (metres; seed = dn * e_k with dn = d xn / d x_hat), so a thread's 8 rows hold 2 complete points.
enum RowMode { PLAIN = 0, DUAL = 1 };

// xs: [TM][4] normalised points (DUAL: only rows 4q are read).  After the call A[row][0..255] = last hidden
// activation (value rows) / its tangents, and out[row] = raw network output (value rows: + b6) / tangent.
// cf_save (optional, PLAIN only): global [6][TM][256] gets 30*f*cos(arg) per layer for the reverse pass.
template <RowMode MODE>
__device__ __forceinline__ void sdf_tile_forward(const FrameParams& fp, const float (*xs)[4], float* A, const int lda,
                                                 WPipe& wp, float* out, float* cf_save, const float dn) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Arow = A + warp * 8 * lda;
    float acc[8][8];
    // ---- layer 0 (K = 3): direct
    {
        float w0[8], w1[8], w2[8];
        load_cols<256>(w0, fp.sdf_Wt[0], lane);
        load_cols<256>(w1, fp.sdf_Wt[0] + 256, lane);
        load_cols<256>(w2, fp.sdf_Wt[0] + 512, lane);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int vr = (MODE == DUAL) ? (r & ~3) : r;            // value row feeding this row
            const float x = xs[warp * 8 + vr][0], y = xs[warp * 8 + vr][1], z = xs[warp * 8 + vr][2];
            const int ty = r & 3;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (MODE == DUAL && ty != 0) acc[r][c] = dn * ((ty == 1) ? w0[c] : (ty == 2) ? w1[c] : w2[c]);
                else acc[r][c] = fmaf(w2[c], z, fmaf(w1[c], y, w0[c] * x));
            }
        }
    }
    for (int l = 0; l < 6; ++l) {
        if (l > 0) tile_gemm<256>(acc, Arow, lda, SDF_H, fp.sdf_Wt[l], wp);
        float b[8], f[8], ph[8];
        load_cols<256>(b, fp.sdf_b[l], lane);
        load_cols<256>(f, fp.sdf_freq + l * SDF_H, lane);
        load_cols<256>(ph, fp.sdf_phase + l * SDF_H, lane);
        if constexpr (MODE == PLAIN) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float h[8], cf[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float arg = 30.0f * (f[c] * (acc[r][c] + b[c]) + ph[c]);
                    if (cf_save) { float s, co; sincosf(arg, &s, &co); h[c] = s; cf[c] = 30.0f * f[c] * co; }
                    else h[c] = sinf(arg);
                }
                store_row<256>(Arow + r * lda, h, lane);
                if (cf_save) store_row<256>(cf_save + ((size_t)l * TM + warp * 8 + r) * SDF_H, cf, lane);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float h[8], cf[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float arg = 30.0f * (f[c] * (acc[4 * q][c] + b[c]) + ph[c]);
                    float s, co; sincosf(arg, &s, &co);
                    h[c] = s; cf[c] = 30.0f * f[c] * co;
                }
                store_row<256>(Arow + (4 * q) * lda, h, lane);
#pragma unroll
                for (int t = 1; t < 4; ++t) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) h[c] = cf[c] * acc[4 * q + t][c];
                    store_row<256>(Arow + (4 * q + t) * lda, h, lane);
                }
            }
        }
        __syncwarp();
    }
    // ---- output layer 256 -> 1
    float y[8];
    warp_rows_dot(Arow, lda, SDF_H, fp.sdf_w6, y, lane);
    if (lane < 8) {
        float v = y[0];
#pragma unroll
        for (int r = 1; r < 8; ++r) if (lane == r) v = y[r];
        const bool is_value = (MODE == PLAIN) || ((lane & 3) == 0);
        out[warp * 8 + lane] = is_value ? (v + __ldg(fp.sdf_b6)) : v;
    }
}

// reverse pass (PLAIN rows): A holds h5 on entry is NOT required; uses cf_save[l] (global) from the forward.
// On exit grad[row][0..2] = d sdf_raw / d xn.
__device__ __forceinline__ void sdf_tile_backward(const FrameParams& fp, float* A, const int lda, WPipe& wp,
                                                  const float* cf_save, float (*grad)[4]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Arow = A + warp * 8 * lda;
    float acc[8][8];
    {   // g_a5 = w6 * cf5
        float w6[8];
        load_cols<256>(w6, fp.sdf_w6, lane);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float cf[8], g[8];
            load_cols_rw<256>(cf, cf_save + ((size_t)5 * TM + warp * 8 + r) * SDF_H, lane);
#pragma unroll
            for (int c = 0; c < 8; ++c) g[c] = w6[c] * cf[c];
            store_row<256>(Arow + r * lda, g, lane);
        }
        __syncwarp();
    }
    for (int l = 5; l >= 1; --l) {
        // g_h(l-1) = g_a(l) @ W_l   (W_l is [out][in] = [K][N] as stored by the reference)
        tile_gemm<256>(acc, Arow, lda, SDF_H, fp.sdf_W[l], wp);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float cf[8], g[8];
            load_cols_rw<256>(cf, cf_save + ((size_t)(l - 1) * TM + warp * 8 + r) * SDF_H, lane);
#pragma unroll
            for (int c = 0; c < 8; ++c) g[c] = acc[r][c] * cf[c];
            store_row<256>(Arow + r * lda, g, lane);
        }
        __syncwarp();
    }
    // grad_xn[j] = sum_o g_a0[o] W0[o][j]
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float y[8];
        for (int r = 0; r < 8; ++r) y[r] = 0.0f;
        for (int k = lane; k < SDF_H; k += 32) {
            const float wk = __ldg(fp.sdf_W[0] + k * 3 + j);
#pragma unroll
            for (int r = 0; r < 8; ++r) y[r] = fmaf(Arow[r * lda + k], wk, y[r]);
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) y[r] += __shfl_xor_sync(0xffffffffu, y[r], o);
        }
        if (lane < 8) {
            float v = y[0];
#pragma unroll
            for (int r = 1; r < 8; ++r) if (lane == r) v = y[r];
            grad[warp * 8 + lane][j] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ skinning tile
// logits[row][0..24] = network output (value rows: + bias; DUAL tangent rows: d/dx_hat_k), [25..31] = 0
template <RowMode MODE>
__device__ __forceinline__ void skin_tile_forward(const FrameParams& fp, const float (*xs)[4], float* A, const int lda,
                                                  WPipe& wp, float (*logits)[32], const float dn) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Arow = A + warp * 8 * lda;
    float acc[8][4];
    {
        float w0[4], w1[4], w2[4];
        load_cols<128>(w0, fp.skin_Wt[0], lane);
        load_cols<128>(w1, fp.skin_Wt[0] + 128, lane);
        load_cols<128>(w2, fp.skin_Wt[0] + 256, lane);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int vr = (MODE == DUAL) ? (r & ~3) : r;
            const float x = xs[warp * 8 + vr][0], y = xs[warp * 8 + vr][1], z = xs[warp * 8 + vr][2];
            const int ty = r & 3;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (MODE == DUAL && ty != 0) acc[r][c] = dn * ((ty == 1) ? w0[c] : (ty == 2) ? w1[c] : w2[c]);
                else acc[r][c] = fmaf(w2[c], z, fmaf(w1[c], y, w0[c] * x));
            }
        }
    }
    for (int l = 0; l < 4; ++l) {
        if (l > 0) tile_gemm<128>(acc, Arow, lda, SKIN_H, fp.skin_Wt[l], wp);
        float b[4];
        load_cols<128>(b, fp.skin_b[l], lane);
        if constexpr (MODE == PLAIN) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float h[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) h[c] = softplus100(acc[r][c] + b[c]);
                store_row<128>(Arow + r * lda, h, lane);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float h[4], gs[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float a = acc[4 * q][c] + b[c];
                    h[c] = softplus100(a);
                    gs[c] = softplus100_grad(a);
                }
                store_row<128>(Arow + (4 * q) * lda, h, lane);
#pragma unroll
                for (int t = 1; t < 4; ++t) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) h[c] = gs[c] * acc[4 * q + t][c];
                    store_row<128>(Arow + (4 * q + t) * lda, h, lane);
                }
            }
        }
        __syncwarp();
    }
    float acc1[8][1];
    tile_gemm<32>(acc1, Arow, lda, SKIN_H, fp.skin_Wt[4], wp);
    const float b4 = __ldg(fp.skin_b[4] + lane);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const bool is_value = (MODE == PLAIN) || ((r & 3) == 0);
        logits[warp * 8 + r][lane] = is_value ? (acc1[r][0] + b4) : acc1[r][0];
    }
}

// per-point consumer of the skinning logits: w = hsoftmax(20*logits), T = sum w B, x_bar = T [x;1]
__device__ __forceinline__ void skin_point(const FrameParams& fp, const float* lg32, const float* x_hat, float* T12, float* x_bar) {
    float lg[25], w[NJ];
#pragma unroll
    for (int k = 0; k < 25; ++k) lg[k] = lg32[k] * 20.0f;
    hierarchical_softmax(lg, w);
    blend_T(w, fp.bone_T, T12, nullptr);
    apply_T(T12, x_hat, x_bar);
}

}  // namespace arah
