// arah_hyper.cu — the MetaAvatar hypernetwork forward: pose -> the frame's FiLM-SIREN SDF parameters.
//
// Replaces HyperBVPNet.forward up to the assembled decoder (metaavatar/models/siren_modules.py:280-312):
//   HierarchicalPoseEncoder                     siren_modules.py:196-244     rots [24][9], Jtrs [24][3] -> cond [144]
//   CustomMappingNetwork                        hyperlayers.py:107-139       latent [128] -> freq / phase_shift [6][256]
//   7 x HyperLinear(FiLM).hypo_params (FCBlock) hyperlayers.py:426-510, pytorch_prototyping.py:12-81
//        FCLayer(144,256) -> FCLayer(256,256) -> Linear(256, in*out + out), + hypo_params_init, split into weights / biases
// One row of batch (the reference renders one frame per call, models/__init__.py:152).
//
// 86.6 M parameters, 99.6 % of them in the seven output matrices [in*out + out][256]: the op is a batch-1 GEMV that reads
// 341 MB once — HBM-bound by construction (2 FLOP per 4 bytes).  Two launches:
//   k_hyper_head : 8 CTAs.  CTA l < 7: pose encoder (recomputed per CTA: 13 k MAC) + the two LayerNorm/ReLU layers of hypo
//                  layer l -> hidden[l][256];  CTA 7: the three LeakyReLU layers of the mapping network -> hidden[7][256].
//   k_hyper_gemv : all 333 313 output rows (7 hypo layers + the mapping network's last layer) as one row space; a warp takes 8
//                  consecutive rows, each lane keeps its 8 hidden values in registers, a row is two coalesced 512-byte streaming
//                  loads, reduced with a shuffle butterfly, and y = dot + bias (+ hypo_params_init) is written straight into
//                  the reference's weights / biases / freq / phase_shift tensors (the layout ArahFrame.sdf_* expects).
// Summation order differs from the reference's MKL/cuBLAS sgemv, so agreement is to fp32 rounding (tests: 2e-5 absolute on
// O(1) outputs), not bitwise.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/arah_b200.h"

extern "C" int arah_internal_fail(int code, const char* msg);

namespace arah_hyper {

constexpr int HID = 256, COND = 144, NJ = 24, NGROUP = 8, ROWS_PER_WARP = 8;

struct HeadArgs {
    ArahHyperWeights w;
    const float* rots; const float* Jtrs; const float* latent;
    float* hidden;          // [8][256]
};

__device__ __constant__ int c_parent[NJ] = {-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};
// joints grouped by depth in the kinematic tree: a level only needs its parents' features, so 9 steps instead of 24
__device__ __constant__ int c_level_start[10] = {0, 1, 4, 7, 10, 15, 18, 20, 22, 24};
__device__ __constant__ int c_level_joint[NJ] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23};

constexpr int HEAD_THREADS = 1024;

// y[r] = b[r] + sum_k W[r][k] x[k] for r < rows: a warp takes 4 rows at a time (coalesced row reads, 4 x cols/32 independent
// loads in flight per lane), x in shared memory
__device__ __forceinline__ void dense_rows(const float* __restrict__ W, const float* __restrict__ b, const float* x, int rows, int cols, float* y) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r0 = 4 * warp; r0 < rows; r0 += 4 * nw) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = lane; k < cols; k += 32) {
            const float xv = x[k];
#pragma unroll
            for (int i = 0; i < 4; ++i) if (r0 + i < rows) s[i] = fmaf(__ldg(W + (size_t)(r0 + i) * cols + k), xv, s[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < 4; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
        if (lane < 4 && r0 + lane < rows) {
            const float v = lane == 0 ? s[0] : lane == 1 ? s[1] : lane == 2 ? s[2] : s[3];
            y[r0 + lane] = v + __ldg(b + r0 + lane);
        }
    }
}
// nn.LayerNorm([256]) (eps 1e-5, biased variance) followed by ReLU, in place; the first 256 threads hold one element each
__device__ __forceinline__ void layernorm_relu(float* v, const float* __restrict__ g, const float* __restrict__ b, float* red) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool on = t < HID;
    const float x = on ? v[t] : 0.f;
    float s = x;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (on && lane == 0) red[warp] = s;
    __syncthreads();
    float mean = 0.f;
    for (int i = 0; i < 8; ++i) mean += red[i];
    mean *= (1.0f / HID);
    __syncthreads();
    const float d = x - mean;
    float q = on ? d * d : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (on && lane == 0) red[warp] = q;
    __syncthreads();
    float var = 0.f;
    for (int i = 0; i < 8; ++i) var += red[i];
    var *= (1.0f / HID);
    if (on) v[t] = fmaxf(d * rsqrtf(var + 1e-5f) * __ldg(g + t) + __ldg(b + t), 0.f);
    __syncthreads();
}

__global__ void __launch_bounds__(HEAD_THREADS) k_hyper_head(HeadArgs a) {
    __shared__ float in[288], feat[NJ][6], gfeat[6], cond[COND], v0[HID], v1[HID], red[8];
    __shared__ float sW1[NJ * 19 * 19], x19[NJ][20], h19[NJ][20];
    const int t = threadIdx.x, l = blockIdx.x, warp = t >> 5, lane = t & 31;
    const ArahHyperWeights& w = a.w;
    if (l == 7) {                                   // mapping network, hidden part (hyperlayers.py:112-121): LeakyReLU(0.2)
        if (t < 128) in[t] = a.latent ? a.latent[t] : 0.f;
        __syncthreads();
        dense_rows(w.map_W[0], w.map_b[0], in, HID, 128, v0);
        __syncthreads();
        if (t < HID) v0[t] = v0[t] > 0.f ? v0[t] : 0.2f * v0[t];
        __syncthreads();
        dense_rows(w.map_W[1], w.map_b[1], v0, HID, HID, v1);
        __syncthreads();
        if (t < HID) v1[t] = v1[t] > 0.f ? v1[t] : 0.2f * v1[t];
        __syncthreads();
        dense_rows(w.map_W[2], w.map_b[2], v1, HID, HID, v0);
        __syncthreads();
        if (t < HID) a.hidden[7 * HID + t] = v0[t] > 0.f ? v0[t] : 0.2f * v0[t];
        return;
    }
    // ---- HierarchicalPoseEncoder (siren_modules.py:217-244): weights staged once (coalesced), joints evaluated level by level
    for (int i = t; i < NJ * 361; i += HEAD_THREADS) sW1[i] = __ldg(w.pe_W1 + i);
    for (int i = t; i < 216; i += HEAD_THREADS) in[i] = a.rots[i];
    for (int i = t; i < 72; i += HEAD_THREADS) {
        const int j = i / 3, k = i % 3;
        float v = a.Jtrs[i];
        if (w.rel_joints && j > 0) v -= a.Jtrs[3 * c_parent[j] + k];          // :220-224
        in[216 + i] = v;
    }
    __syncthreads();
    dense_rows(w.pe_l0_W, w.pe_l0_b, in, 6, 288, gfeat);                      // global_feat (:226-227)
    __syncthreads();
    for (int lev = 0; lev < 9; ++lev) {
        const int n = c_level_start[lev + 1] - c_level_start[lev];
        const int j = (warp < n) ? c_level_joint[c_level_start[lev] + warp] : -1;     // one warp per joint of the level
        if (j >= 0 && lane < 19) {
            const int p = c_parent[j];
            float v;
            if (lane < 9) v = in[9 * j + lane];
            else if (lane < 12) v = in[216 + 3 * j + (lane - 9)];
            else if (lane == 12) {                                            // bone length (:235,239)
                float d[3];
                for (int k = 0; k < 3; ++k) {
                    d[k] = in[216 + 3 * j + k];
                    if (p >= 0 && !w.rel_joints) d[k] -= in[216 + 3 * p + k];
                }
                v = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            } else v = (p < 0) ? gfeat[lane - 13] : feat[p][lane - 13];
            x19[j][lane] = v;
        }
        __syncwarp();
        if (j >= 0 && lane < 19) {
            float s = __ldg(w.pe_b1 + 19 * j + lane);
            for (int k = 0; k < 19; ++k) s = fmaf(sW1[(19 * j + lane) * 19 + k], x19[j][k], s);
            h19[j][lane] = fmaxf(s, 0.f);
        }
        __syncwarp();
        if (j >= 0 && lane < 6) {
            float s = __ldg(w.pe_b2 + 6 * j + lane);
#pragma unroll
            for (int k = 0; k < 19; ++k) s = fmaf(__ldg(w.pe_W2 + (6 * j + lane) * 19 + k), h19[j][k], s);
            feat[j][lane] = s;
            cond[6 * j + lane] = s;
        }
        __syncthreads();
    }
    // ---- FCBlock hidden layers of hypo layer l (pytorch_prototyping.py:60-63): Linear -> LayerNorm -> ReLU, twice
    dense_rows(w.fc1_W[l], w.fc1_b[l], cond, HID, COND, v0);
    __syncthreads();
    layernorm_relu(v0, w.ln1_g[l], w.ln1_b[l], red);
    dense_rows(w.fc2_W[l], w.fc2_b[l], v0, HID, HID, v1);
    __syncthreads();
    layernorm_relu(v1, w.ln2_g[l], w.ln2_b[l], red);
    if (t < HID) a.hidden[l * HID + t] = v1[t];
}

struct GemvGroup {
    const float* W; const float* b; const float* init;     // [rows][256], [rows], [rows] or null
    float* out_a; float* out_b;                             // rows [0, n_a) -> out_a, rows [n_a, rows) -> out_b
    int rows, n_a, chunk0;                                  // chunk0: first ROWS_PER_WARP-row chunk of this group in the global chunk space
};
struct GemvArgs { GemvGroup g[NGROUP]; const float* hidden; int nchunks; };

__global__ void __launch_bounds__(256) k_hyper_gemv(GemvArgs a) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    int cur = -1;
    float4 h0 = make_float4(0, 0, 0, 0), h1 = h0;
    for (int c = gw; c < a.nchunks; c += nw) {
        int gi = 0;
#pragma unroll
        for (int i = 1; i < NGROUP; ++i) gi += (c >= a.g[i].chunk0) ? 1 : 0;
        const GemvGroup& g = a.g[gi];
        if (gi != cur) {                                     // this lane's 8 hidden values: k = 4 lane .. +3 and 128 + 4 lane .. +3
            const float4* hp = reinterpret_cast<const float4*>(a.hidden + gi * HID);
            h0 = hp[lane]; h1 = hp[32 + lane];
            cur = gi;
        }
        const int r0 = (c - g.chunk0) * ROWS_PER_WARP;
        const int nr = min(ROWS_PER_WARP, g.rows - r0);
        // 16 unconditional 16-byte streaming loads per lane issued back to back (8 KB per warp in flight); rows past the end
        // of a group re-read its last row (their results are never stored)
        float4 x0[ROWS_PER_WARP], x1[ROWS_PER_WARP];
        const int rlast = g.rows - 1;
#pragma unroll
        for (int i = 0; i < ROWS_PER_WARP; ++i) {
            const float4* rp = reinterpret_cast<const float4*>(g.W + (size_t)min(r0 + i, rlast) * HID);
            x0[i] = __ldcs(rp + lane); x1[i] = __ldcs(rp + 32 + lane);
        }
        float s[ROWS_PER_WARP];
#pragma unroll
        for (int i = 0; i < ROWS_PER_WARP; ++i) {
            float t = x0[i].x * h0.x;
            t = fmaf(x0[i].y, h0.y, t); t = fmaf(x0[i].z, h0.z, t); t = fmaf(x0[i].w, h0.w, t);
            t = fmaf(x1[i].x, h1.x, t); t = fmaf(x1[i].y, h1.y, t); t = fmaf(x1[i].z, h1.z, t); t = fmaf(x1[i].w, h1.w, t);
            s[i] = t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int i = 0; i < ROWS_PER_WARP; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
        if (lane < nr) {
            float y = 0.f;
#pragma unroll
            for (int i = 0; i < ROWS_PER_WARP; ++i) y = (lane == i) ? s[i] : y;
            const int r = r0 + lane;
            y += __ldg(g.b + r);
            if (g.init) y += __ldg(g.init + r);
            if (r < g.n_a) g.out_a[r] = y; else g.out_b[r - g.n_a] = y;
        }
    }
}

}  // namespace arah_hyper

using namespace arah_hyper;

#define HCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return arah_internal_fail(ARAH_ECUDA, (std::string(#x) + ": " + cudaGetErrorString(e_)).c_str()); } while (0)

extern "C" size_t arah_hyper_workspace(void) { return (size_t)NGROUP * HID * sizeof(float); }

extern "C" int arah_hyper_forward(const ArahHyperWeights* w, const float* rots, const float* Jtrs, const float* latent,
                                  const ArahSdfParams* out, void* workspace, void* stream) {
    if (!w || !rots || !Jtrs || !out || !workspace) return arah_internal_fail(ARAH_EINVAL, "null argument");
    static const int in_ch[7] = {3, 256, 256, 256, 256, 256, 256}, out_ch[7] = {256, 256, 256, 256, 256, 256, 1};
    if (!w->pe_l0_W || !w->pe_l0_b || !w->pe_W1 || !w->pe_b1 || !w->pe_W2 || !w->pe_b2) return arah_internal_fail(ARAH_EINVAL, "pose-encoder weights missing");
    for (int i = 0; i < 4; ++i) if (!w->map_W[i] || !w->map_b[i]) return arah_internal_fail(ARAH_EINVAL, "mapping-network weights missing");
    for (int l = 0; l < 7; ++l) {
        if (!w->fc1_W[l] || !w->fc1_b[l] || !w->ln1_g[l] || !w->ln1_b[l] || !w->fc2_W[l] || !w->fc2_b[l] || !w->ln2_g[l] || !w->ln2_b[l] ||
            !w->out_W[l] || !w->out_b[l]) return arah_internal_fail(ARAH_EINVAL, "hypo-layer weights missing");
        if (!out->sdf_W[l] || !out->sdf_b[l]) return arah_internal_fail(ARAH_EINVAL, "output buffers missing");
    }
    if (!out->sdf_freq || !out->sdf_phase) return arah_internal_fail(ARAH_EINVAL, "output buffers missing");
    cudaStream_t st = (cudaStream_t)stream;
    HeadArgs ha;
    ha.w = *w; ha.rots = rots; ha.Jtrs = Jtrs; ha.latent = latent; ha.hidden = (float*)workspace;
    k_hyper_head<<<NGROUP, HEAD_THREADS, 0, st>>>(ha);
    GemvArgs ga;
    int chunk = 0;
    for (int l = 0; l < 7; ++l) {
        GemvGroup& g = ga.g[l];
        g.W = w->out_W[l]; g.b = w->out_b[l]; g.init = w->init[l];
        g.out_a = out->sdf_W[l]; g.out_b = out->sdf_b[l];
        g.n_a = in_ch[l] * out_ch[l]; g.rows = g.n_a + out_ch[l]; g.chunk0 = chunk;
        chunk += (g.rows + ROWS_PER_WARP - 1) / ROWS_PER_WARP;
    }
    {   // mapping network's last Linear(256, 3072): first half = frequencies, second half = phase shifts (hyperlayers.py:132-139)
        GemvGroup& g = ga.g[7];
        g.W = w->map_W[3]; g.b = w->map_b[3]; g.init = nullptr;
        g.out_a = out->sdf_freq; g.out_b = out->sdf_phase;
        g.n_a = 6 * HID; g.rows = 12 * HID; g.chunk0 = chunk;
        chunk += g.rows / ROWS_PER_WARP;
    }
    ga.hidden = (const float*)workspace; ga.nchunks = chunk;
    int dev = 0, nsm = 148;
    HCU(cudaGetDevice(&dev));
    HCU(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    k_hyper_gemv<<<nsm * 8, 256, 0, st>>>(ga);        // 8 resident CTAs x 8 warps x 8 KB of rows in flight per SM
    HCU(cudaGetLastError());
    return ARAH_OK;
}
