// arah_train.h — training forward/backward of the ARAH shading path (BASELINE configs[2]).
//
// What the reference does with torch autograd (paths relative to /root/reference/im2mesh):
//   * differentiable shading of the traced samples: implicit-gradient LBS correction, FiLM-SIREN SDF forward, its
//     input gradient as "normal" (create_graph=True, i.e. differentiated a second time), colour MLP, sigma-from-SDF,
//     alpha compositing                          metaavatar_render/renderer/implicit_differentiable_renderer.py:261-396
//   * auxiliary SDF evaluations (eikonal / off-surface / inside points)                               ibid. :117-140
//   * skinning-weight prediction for the skinning loss                  ibid. :73-78, utils/root_finding_utils.py:54-113
// Here the same chain rule is written out by hand, layer by layer, as dense [points x width] matrices in HBM:
// every linear map is one call of the backend's strided GEMM (forward NT, backward-data NN, weight-gradient TN with
// split-K), everything between two GEMMs is a fused element-wise / column-reduction functor.  The file is
// backend-generic: `BK` supplies alloc / gemm / for_each / col_reduce.  The product instantiates it with the CUDA
// backend (arah_train_cuda.cuh); tests/host_train.cpp instantiates it with a plain-loop host backend so the whole
// chain rule is checked against torch.autograd on the CPU without a GPU.  There is no CPU path in the product library.
//
// Layouts: weights in the reference's own [out][in] layout (weight-norm already applied), gradients are ACCUMULATED
// into caller-zeroed buffers of the same layout.  Per-point matrices are row-major, leading dimension = width.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <math.h>
#include "arah_math.cuh"

namespace arah {
namespace train {

constexpr int SDF_HID = 256, SKIN_HID = 128, COL_HID = 256, COL_MID = 128;
constexpr int COL_XIN = 33;        // [xn 3 | PE4(view) 27 | normal 3]   (metaavatar_render/models/decoder.py:105)
constexpr int COL_XLD = 36;        // padded leading dimension of the 33-wide block
constexpr int SKIN_OUT = 25, SKIN_OLD = 28;

ARAH_HD void atomic_addf(float* p, float v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
#pragma omp atomic
    *p += v;
#endif
}

// ---- parameter / gradient views ------------------------------------------------------------------------------------
struct SdfParams { const float* W[7]; const float* b[7]; const float* freq; const float* phase; };       // [6][256] FiLM
struct SdfGrads { float* W[7]; float* b[7]; float* freq; float* phase; };
struct SkinParams { const float* W[5]; const float* b[5]; };     // [128][3], 3 x [128][128], [25][128]
struct SkinGrads { float* W[5]; float* b[5]; };
struct ColParams { const float* W[6]; const float* b[6]; const float* latent; int latent_dim; };
struct ColGrads { float* W[6]; float* b[6]; float* latent; };
struct Norm { float cmin, cmax, center[3]; };

// =====================================================================================================================
// element-wise functors (one call per matrix element i = m * N + j unless stated otherwise)
// =====================================================================================================================
// h = sin(30 (f * pre + phi))                      hyperlayers.py:412-415 + siren_modules.py:35-37
struct FilmSin {
    const float* pre; const float* f; const float* ph; float* h;
    ARAH_HD void operator()(size_t i) const { const int j = (int)(i % SDF_HID); h[i] = sinf(30.0f * (f[j] * pre[i] + ph[j])); }
};
// dt = up * 30 f cos(30 a); up = a per-point matrix, or (row_const) the single row W6
struct RevMul {
    const float* up; int row_const; const float* pre; const float* f; const float* ph; float* dt;
    ARAH_HD void operator()(size_t i) const {
        const int j = (int)(i % SDF_HID);
        const float u = row_const ? up[j] : up[i];
        dt[i] = u * (30.0f * f[j] * cosf(30.0f * (f[j] * pre[i] + ph[j])));
    }
};
// second-order step of layer l (reads gdt = d L / d dt_l, writes u_next = d L / d delta_{l+1} and the extra gradient
// ga2 on the FiLM argument a_l that the normal's dependence on a_l creates); column sum -> d L / d freq_l
struct DblStep {
    const float* gdt; const float* up; int row_const; const float* pre; const float* f; const float* ph;
    float* u_next; float* ga2;
    ARAH_HD void operator()(int m, int j, float* red) const {
        const size_t i = (size_t)m * SDF_HID + j;
        const float a30 = 30.0f * (f[j] * pre[i] + ph[j]);
        const float c = cosf(a30), s = sinf(a30);
        const float g = gdt[i], d = row_const ? up[j] : up[i];
        u_next[i] = g * (30.0f * f[j] * c);
        const float gc = g * d;                       // d L / d c'_l,  c'_l = 30 f cos(30 a)
        const float gaa = gc * (-900.0f * f[j] * s);  // through cos(30 a)
        ga2[i] = gaa;
        // c' depends on f twice: explicitly and through a = f * pre + phi
        red[0] = gc * 30.0f * c + gaa * pre[i];
        red[1] = gaa;                                 // through a -> phi
    }
};
// first-order step of layer l: g_pre = (g_h * 30 cos(30 a) + ga2) * f ; column sums -> freq, phase, bias
struct FwdBackStep {
    const float* gh; const float* ga2; const float* pre; const float* f; const float* ph; float* gpre;
    ARAH_HD void operator()(int m, int j, float* red) const {
        const size_t i = (size_t)m * SDF_HID + j;
        const float a30 = 30.0f * (f[j] * pre[i] + ph[j]);
        float ga = gh[i] * (30.0f * cosf(a30));
        const float gp1 = ga * f[j];
        red[0] = ga * pre[i]; red[1] = ga;
        // the second-order part already put its own freq/phase contribution into the sums (DblStep); it still has to
        // flow on to pre (and from there to W, b and the layer below)
        float gp = gp1;
        if (ga2) gp += ga2[i] * f[j];
        gpre[i] = gp;
        red[2] = gp;
    }
};
// g_h6 = g_F + g_s (x) W6 ; column sum -> d L / d W6 (first-order part: sum_m g_s h6)
struct HeadBack {
    const float* gF; int ldgF; const float* gs; const float* W6; const float* h6; float* gh;
    ARAH_HD void operator()(int m, int j, float* red) const {
        const size_t i = (size_t)m * SDF_HID + j;
        const float g_s = gs ? gs[m] : 0.0f;
        gh[i] = (gF ? gF[(size_t)m * ldgF + j] : 0.0f) + g_s * W6[j];
        red[0] = g_s * h6[i];
    }
};
struct ColSumVec { const float* v; ARAH_HD void operator()(int m, int j, float* red) const { (void)j; red[0] = v[m]; } };
struct ColSumMat { const float* v; int ld; ARAH_HD void operator()(int m, int j, float* red) const { red[0] = v[(size_t)m * ld + j]; } };
struct Relu { float* h; ARAH_HD void operator()(size_t i) const { h[i] = fmaxf(h[i], 0.0f); } };
// g *= (h > 0), then column sum (bias gradient)
struct ReluBack {
    float* g; const float* h; int N;
    ARAH_HD void operator()(int m, int j, float* red) const {
        const size_t i = (size_t)m * N + j;
        const float v = h[i] > 0.0f ? g[i] : 0.0f;
        g[i] = v; red[0] = v;
    }
};
struct Sigmoid4 { float* o; ARAH_HD void operator()(size_t i) const { o[i] = 1.0f / (1.0f + expf(-o[i])); } };
struct SigmoidBack {
    float* g; const float* o;
    ARAH_HD void operator()(int m, int j, float* red) const {
        const size_t i = (size_t)m * 4 + j;
        const float v = g[i] * o[i] * (1.0f - o[i]);
        g[i] = v; red[0] = v;
    }
};
// b'[o] = b[o] + W[o][col0 .. col0+L) . latent     (per-frame-constant pose feature, decoder.py:96-100)
struct FoldLatent {
    const float* W; int ld, col0; const float* lat; int L; const float* b; float* out;
    ARAH_HD void operator()(size_t o) const {
        float a = b[o];
        for (int j = 0; j < L; ++j) a += W[o * ld + col0 + j] * lat[j];
        out[o] = a;
    }
};
// gradient of the fold: gW[o][col0+j] += gb[o] lat[j]; glat[j] += sum_o W[o][col0+j] gb[o]   (i = o * L + j)
struct FoldLatentBack {
    const float* W; float* gW; int ld, col0; const float* lat; float* glat; int L; const float* gb;
    ARAH_HD void operator()(size_t i) const {
        const size_t o = i / L; const int j = (int)(i % L);
        if (gW) gW[o * ld + col0 + j] += gb[o] * lat[j];
        if (glat) atomic_addf(glat + j, W[o * ld + col0 + j] * gb[o]);
    }
};
struct Softplus { const float* pre; float* a; ARAH_HD void operator()(size_t i) const { a[i] = softplus100(pre[i]); } };
struct SoftplusBack {
    float* g; const float* pre; int N;
    ARAH_HD void operator()(int m, int j, float* red) const {
        const size_t i = (size_t)m * N + j;
        const float v = g[i] * softplus100_grad(pre[i]);
        g[i] = v; red[0] = v;
    }
};
// tangent rows: t[(3m+k)][j] *= softplus'(pre[m][j])
struct SoftplusTangent {
    float* t; const float* pre; int N;
    ARAH_HD void operator()(size_t i) const { const size_t row = i / N; const int j = (int)(i % N); t[i] *= softplus100_grad(pre[(row / 3) * N + j]); }
};
// first tangent rows of the skinning MLP: d pre0 / d xn_k = W0[:, k]
struct SkinTangent0 {
    const float* W0; const float* pre0; float* t;
    ARAH_HD void operator()(size_t i) const {
        const size_t row = i / SKIN_HID; const int j = (int)(i % SKIN_HID); const int k = (int)(row % 3);
        t[i] = W0[j * 3 + k] * softplus100_grad(pre0[(row / 3) * SKIN_HID + j]);
    }
};

// reverse-mode hierarchical softmax: x = 25 logits already scaled by 20; gp = d L / d p[24]  ->  gx[25]
// (utils/utils.py:138-181; forward order of arah_math.cuh::hierarchical_softmax, replayed backwards)
ARAH_HD void hierarchical_softmax_vjp(const float* x, const float* gp_in, float* gx) {
    // split list in forward order: (child c, parent q, gate g); the two softmax3 groups are handled separately
    float p[NJ], gp[NJ];
    hierarchical_softmax(x, p);
    for (int i = 0; i < NJ; ++i) gp[i] = gp_in[i];
    for (int i = 0; i < 25; ++i) gx[i] = 0.0f;
    // undo a split p[c] = P s, p[q] = P (1 - s) where P = p[c] + p[q] is the parent's value before the split
#define ARAH_UNSPLIT(c, q, g) { const float s_ = sigmoid_(x[g]); const float P_ = p[c] + p[q]; \
        const float gs_ = (gp[c] - gp[q]) * P_; gx[g] += gs_ * s_ * (1.0f - s_); \
        gp[q] = gp[c] * s_ + gp[q] * (1.0f - s_); gp[c] = 0.0f; p[q] = P_; p[c] = 0.0f; }
    ARAH_UNSPLIT(23, 21, 23) ARAH_UNSPLIT(22, 20, 22)
    ARAH_UNSPLIT(21, 19, 21) ARAH_UNSPLIT(20, 18, 20)
    ARAH_UNSPLIT(19, 17, 19) ARAH_UNSPLIT(18, 16, 18)
    ARAH_UNSPLIT(17, 14, 17) ARAH_UNSPLIT(16, 13, 16)
    ARAH_UNSPLIT(15, 12, 15)
    {   // p[12+k] = P9 s24 sm[k]; p[9] = P9 (1 - s24)
        float sm[3];
        softmax3(x + 12, sm);
        const float s24 = sigmoid_(x[24]);
        const float P9 = p[9] + p[12] + p[13] + p[14];
        float gsm[3], dot = 0.0f, gs24 = -gp[9] * P9, gP9 = gp[9] * (1.0f - s24);
        for (int k = 0; k < 3; ++k) { gsm[k] = gp[12 + k] * P9 * s24; gs24 += gp[12 + k] * P9 * sm[k]; gP9 += gp[12 + k] * s24 * sm[k]; dot += gsm[k] * sm[k]; }
        for (int k = 0; k < 3; ++k) gx[12 + k] += sm[k] * (gsm[k] - dot);
        gx[24] += gs24 * s24 * (1.0f - s24);
        gp[9] = gP9; p[9] = P9;
        for (int k = 0; k < 3; ++k) { gp[12 + k] = 0.0f; p[12 + k] = 0.0f; }
    }
    ARAH_UNSPLIT(11, 8, 11) ARAH_UNSPLIT(10, 7, 10)
    ARAH_UNSPLIT(9, 6, 9) ARAH_UNSPLIT(8, 5, 8) ARAH_UNSPLIT(7, 4, 7)
    ARAH_UNSPLIT(6, 3, 6) ARAH_UNSPLIT(5, 2, 5) ARAH_UNSPLIT(4, 1, 4)
#undef ARAH_UNSPLIT
    {   // p[1+k] = s0 sm[k]; p[0] = 1 - s0
        float sm[3];
        softmax3(x + 1, sm);
        const float s0 = sigmoid_(x[0]);
        float gsm[3], dot = 0.0f, gs0 = -gp[0];
        for (int k = 0; k < 3; ++k) { gsm[k] = gp[1 + k] * s0; gs0 += gp[1 + k] * sm[k]; dot += gsm[k] * sm[k]; }
        for (int k = 0; k < 3; ++k) gx[1 + k] += sm[k] * (gsm[k] - dot);
        gx[0] += gs0 * s0 * (1.0f - s0);
    }
}

// =====================================================================================================================
// SDF network (FiLM-SIREN 3 -> 256 x 6 -> 1): value, feature, input gradient, and the backward of all three
// =====================================================================================================================
template <class BK>
struct SdfNet {
    typedef typename BK::Stream Stream;
    int cap = 0, n = 0;
    bool has_normal = false;
    float* x = nullptr;           // [n][4]   normalised points
    float* pre[6] = {};           // [n][256] W h + b
    float* h[6] = {};             // [n][256] sin(30 (f pre + phi)); h[5] = feature
    float* dl[5] = {};            // [n][256] dl[l] = d sdf / d h[l]  (delta_{l+1} in the notes), l = 0..4
    float* ga2[6] = {};           // [n][256] second-order gradient on the FiLM arguments
    float* s = nullptr;           // [n]
    float* nrm = nullptr;         // [n][4]   d sdf / d x
    float* t0 = nullptr; float* t1 = nullptr; float* t2 = nullptr; float* t3 = nullptr;   // [n][256] scratch
    float* u0 = nullptr;          // [n][4] scratch

    void release() {
        BK::free(x); for (auto& p : pre) BK::free(p); for (auto& p : h) BK::free(p); for (auto& p : dl) BK::free(p);
        for (auto& p : ga2) BK::free(p);
        BK::free(s); BK::free(nrm); BK::free(t0); BK::free(t1); BK::free(t2); BK::free(t3); BK::free(u0);
        *this = SdfNet();
    }
    bool reserve(int n_) {
        if (n_ <= cap) return true;
        release();
        const size_t c = ((size_t)n_ + 1023) / 1024 * 1024, w = c * SDF_HID;
        bool ok = true;
        auto A = [&](float*& p, size_t f) { p = BK::alloc(f); ok = ok && p; };
        A(x, c * 4); for (auto& p : pre) A(p, w); for (auto& p : h) A(p, w); for (auto& p : dl) A(p, w); for (auto& p : ga2) A(p, w);
        A(s, c); A(nrm, c * 4); A(t0, w); A(t1, w); A(t2, w); A(t3, w); A(u0, c * 4);
        if (!ok) { release(); return false; }
        cap = (int)c;
        return true;
    }

    // forward of the points already stored in x[0..n): fills pre, h, s and (with_normal) dl, nrm
    void forward(const SdfParams& P, int n_, bool with_normal, Stream st) {
        n = n_; has_normal = with_normal;
        if (n == 0) return;
        const size_t w = (size_t)n * SDF_HID;
        for (int l = 0; l < 6; ++l) {
            const int K = l ? SDF_HID : 3;
            const float* in = l ? h[l - 1] : x;
            BK::gemm(n, SDF_HID, K, in, l ? SDF_HID : 4, 1, P.W[l], 1, K, pre[l], SDF_HID, P.b[l], false, st);   // pre = in W^T + b
            BK::for_each(w, FilmSin{pre[l], P.freq + l * SDF_HID, P.phase + l * SDF_HID, h[l]}, st);
        }
        BK::gemm(n, 1, SDF_HID, h[5], SDF_HID, 1, P.W[6], 1, SDF_HID, s, 1, P.b[6], false, st);
        if (!with_normal) return;
        // reverse sweep: dt_l = delta_{l+1} * 30 f cos(30 a_l);  delta_l = dt_l W_l      (autograd of hyperlayers.py:412-415)
        for (int l = 5; l >= 0; --l) {
            const float* up = (l == 5) ? P.W[6] : dl[l];
            BK::for_each(w, RevMul{up, l == 5, pre[l], P.freq + l * SDF_HID, P.phase + l * SDF_HID, t0}, st);
            if (l > 0) BK::gemm(n, SDF_HID, SDF_HID, t0, SDF_HID, 1, P.W[l], SDF_HID, 1, dl[l - 1], SDF_HID, nullptr, false, st);
            else BK::gemm(n, 3, SDF_HID, t0, SDF_HID, 1, P.W[0], 3, 1, nrm, 4, nullptr, false, st);
        }
    }

    // backward: g_s [n] (d L / d sdf), g_F [n][ldgF] (d L / d feature), g_n [n][4] (d L / d normal); any may be null.
    // Accumulates parameter gradients into G (null members are skipped) and, if g_x != null, ADDS d L / d x into g_x [n][4].
    void backward(const SdfParams& P, const SdfGrads& G, const float* g_s, const float* g_F, int ldgF, const float* g_n, float* g_x, Stream st) {
        if (n == 0) return;
        const size_t w = (size_t)n * SDF_HID;
        const bool second = g_n != nullptr && has_normal;
        if (second) {
            // u_l = d L / d delta_l (u_0 = g_n).  delta_l = dt_l W_l  =>  g_dt_l = u_l W_l^T,  g_W_l += dt_l^T u_l
            const float* u = g_n; int ldu = 4;
            for (int l = 0; l < 6; ++l) {
                const int K = l ? SDF_HID : 3;
                const float* up = (l == 5) ? P.W[6] : dl[l];
                const float* f = P.freq + l * SDF_HID; const float* ph = P.phase + l * SDF_HID;
                BK::gemm(n, SDF_HID, K, u, ldu, 1, P.W[l], 1, K, t0, SDF_HID, nullptr, false, st);                    // g_dt_l
                if (G.W[l]) {
                    BK::for_each(w, RevMul{up, l == 5, pre[l], f, ph, t1}, st);                                       // dt_l (recomputed)
                    BK::gemm(SDF_HID, K, n, t1, 1, SDF_HID, u, ldu, 1, G.W[l], K, nullptr, true, st);                 // g_W_l += dt_l^T u_l
                }
                float* un = (l & 1) ? t3 : t2;      // u_{l+1}; u_l (l >= 1) lives in the other one
                float* red[2] = {G.freq ? G.freq + l * SDF_HID : nullptr, G.phase ? G.phase + l * SDF_HID : nullptr};
                BK::template col_reduce<2>(n, SDF_HID, DblStep{t0, up, l == 5, pre[l], f, ph, un, ga2[l]}, red, st);
                u = un; ldu = SDF_HID;
            }
            // delta_6 = W6 itself: g_W6 += sum_m u_6
            if (G.W[6]) { float* red[1] = {G.W[6]}; BK::template col_reduce<1>(n, SDF_HID, ColSumMat{u, SDF_HID}, red, st); }
        }
        // ---- first-order graph: s = h6 W6^T + b6, F = h6, h_{l+1} = sin(30 (f (h_l W_l^T + b_l) + phi))
        if (g_s && G.b[6]) { float* red[1] = {G.b[6]}; BK::template col_reduce<1>(n, 1, ColSumVec{g_s}, red, st); }
        {
            float* red[1] = {(g_s && G.W[6]) ? G.W[6] : nullptr};
            BK::template col_reduce<1>(n, SDF_HID, HeadBack{g_F, ldgF, g_s, P.W[6], h[5], t0}, red, st);               // t0 = g_h6
        }
        float* gh = t0; float* gp = t1; float* nx = t2;
        for (int l = 5; l >= 0; --l) {
            const int K = l ? SDF_HID : 3;
            const float* f = P.freq + l * SDF_HID; const float* ph = P.phase + l * SDF_HID;
            float* red[3] = {G.freq ? G.freq + l * SDF_HID : nullptr, G.phase ? G.phase + l * SDF_HID : nullptr, G.b[l]};
            BK::template col_reduce<3>(n, SDF_HID, FwdBackStep{gh, second ? ga2[l] : nullptr, pre[l], f, ph, gp}, red, st);
            const float* in = l ? h[l - 1] : x;
            if (G.W[l]) BK::gemm(SDF_HID, K, n, gp, 1, SDF_HID, in, l ? SDF_HID : 4, 1, G.W[l], K, nullptr, true, st);   // g_W_l += g_pre^T in
            if (l > 0) { BK::gemm(n, SDF_HID, SDF_HID, gp, SDF_HID, 1, P.W[l], SDF_HID, 1, nx, SDF_HID, nullptr, false, st); float* sw = gh; gh = nx; nx = sw; }
            else if (g_x) BK::gemm(n, 3, SDF_HID, gp, SDF_HID, 1, P.W[0], 3, 1, g_x, 4, nullptr, true, st);
        }
    }
};


struct AddVec { float* dst; const float* src; ARAH_HD void operator()(size_t i) const { dst[i] += src[i]; } };

// =====================================================================================================================
// colour network (metaavatar_render/models/decoder.py:69-124; mode 'idr', multires_view 4, skip-concat of the input at
// lin3, pose feature = per-frame latent).  Input columns of lin0: [xin 33 | feature 256 | latent L]; lin3 sees
// [xin 33 | feature 256 | latent L | lin2 output 128].  The latent block is folded into the bias (same for every point).
// =====================================================================================================================
template <class BK>
struct ColNet {
    typedef typename BK::Stream Stream;
    int cap = 0, n = 0;
    float* xin = nullptr;                    // [n][36]
    const float* F = nullptr;                // [n][256] borrowed (SdfNet::h[5])
    float* b0f = nullptr; float* b3f = nullptr; float* gb = nullptr;     // [256]
    float* h1 = nullptr; float* h2 = nullptr; float* h4 = nullptr; float* h5 = nullptr;   // [n][256] post-ReLU
    float* h3 = nullptr;                     // [n][128]
    float* o = nullptr;                      // [n][4] rgb (sigmoid)
    float* g0 = nullptr; float* g1 = nullptr; float* g3 = nullptr;      // scratch [n][256], [n][256], [n][128]

    void release() {
        BK::free(xin); BK::free(b0f); BK::free(b3f); BK::free(gb); BK::free(h1); BK::free(h2); BK::free(h4); BK::free(h5);
        BK::free(h3); BK::free(o); BK::free(g0); BK::free(g1); BK::free(g3);
        *this = ColNet();
    }
    bool reserve(int n_) {
        if (n_ <= cap) return true;
        release();
        const size_t c = ((size_t)n_ + 1023) / 1024 * 1024, w = c * COL_HID;
        bool ok = true;
        auto A = [&](float*& p, size_t f) { p = BK::alloc(f); ok = ok && p; };
        A(xin, c * COL_XLD); A(b0f, 256); A(b3f, 256); A(gb, 256); A(h1, w); A(h2, w); A(h4, w); A(h5, w); A(h3, c * COL_MID);
        A(o, c * 4); A(g0, w); A(g1, w); A(g3, c * COL_MID);
        if (!ok) { release(); return false; }
        cap = (int)c;
        return true;
    }
    void forward(const ColParams& P, const float* feat, int n_, Stream st) {
        n = n_; F = feat;
        if (n == 0) return;
        const int L = P.latent_dim, d0 = COL_XIN + SDF_HID + L, d3 = d0 + COL_MID;
        const size_t w = (size_t)n * COL_HID;
        BK::for_each(256, FoldLatent{P.W[0], d0, COL_XIN + SDF_HID, P.latent, L, P.b[0], b0f}, st);
        BK::for_each(256, FoldLatent{P.W[3], d3, COL_XIN + SDF_HID, P.latent, L, P.b[3], b3f}, st);
        BK::gemm(n, COL_HID, COL_XIN, xin, COL_XLD, 1, P.W[0], 1, d0, h1, COL_HID, b0f, false, st);
        BK::gemm(n, COL_HID, SDF_HID, F, SDF_HID, 1, P.W[0] + COL_XIN, 1, d0, h1, COL_HID, nullptr, true, st);
        BK::for_each(w, Relu{h1}, st);
        BK::gemm(n, COL_HID, COL_HID, h1, COL_HID, 1, P.W[1], 1, COL_HID, h2, COL_HID, P.b[1], false, st);
        BK::for_each(w, Relu{h2}, st);
        BK::gemm(n, COL_MID, COL_HID, h2, COL_HID, 1, P.W[2], 1, COL_HID, h3, COL_MID, P.b[2], false, st);
        BK::for_each((size_t)n * COL_MID, Relu{h3}, st);
        BK::gemm(n, COL_HID, COL_XIN, xin, COL_XLD, 1, P.W[3], 1, d3, h4, COL_HID, b3f, false, st);
        BK::gemm(n, COL_HID, SDF_HID, F, SDF_HID, 1, P.W[3] + COL_XIN, 1, d3, h4, COL_HID, nullptr, true, st);
        BK::gemm(n, COL_HID, COL_MID, h3, COL_MID, 1, P.W[3] + d0, 1, d3, h4, COL_HID, nullptr, true, st);
        BK::for_each(w, Relu{h4}, st);
        BK::gemm(n, COL_HID, COL_HID, h4, COL_HID, 1, P.W[4], 1, COL_HID, h5, COL_HID, P.b[4], false, st);
        BK::for_each(w, Relu{h5}, st);
        BK::zero(o, (size_t)n * 4, st);
        BK::gemm(n, 3, COL_HID, h5, COL_HID, 1, P.W[5], 1, COL_HID, o, 4, P.b[5], false, st);
        BK::for_each((size_t)n * 4, Sigmoid4{o}, st);
    }
    // g_o [n][4]: d L / d rgb (overwritten).  Writes g_xin [n][36] and g_F [n][256]; accumulates into G.
    void backward(const ColParams& P, const ColGrads& G, float* g_o, float* g_xin, float* g_F, Stream st) {
        if (n == 0) return;
        const int L = P.latent_dim, d0 = COL_XIN + SDF_HID + L, d3 = d0 + COL_MID;
        float* r1[1];
        r1[0] = G.b[5]; BK::template col_reduce<1>(n, 3, SigmoidBack{g_o, o}, r1, st);
        if (G.W[5]) BK::gemm(3, COL_HID, n, g_o, 1, 4, h5, COL_HID, 1, G.W[5], COL_HID, nullptr, true, st);
        BK::gemm(n, COL_HID, 3, g_o, 4, 1, P.W[5], COL_HID, 1, g0, COL_HID, nullptr, false, st);
        r1[0] = G.b[4]; BK::template col_reduce<1>(n, COL_HID, ReluBack{g0, h5, COL_HID}, r1, st);
        if (G.W[4]) BK::gemm(COL_HID, COL_HID, n, g0, 1, COL_HID, h4, COL_HID, 1, G.W[4], COL_HID, nullptr, true, st);
        BK::gemm(n, COL_HID, COL_HID, g0, COL_HID, 1, P.W[4], COL_HID, 1, g1, COL_HID, nullptr, false, st);
        BK::zero(gb, 256, st);
        r1[0] = gb; BK::template col_reduce<1>(n, COL_HID, ReluBack{g1, h4, COL_HID}, r1, st);
        if (G.b[3]) BK::for_each(256, AddVec{G.b[3], gb}, st);
        if (L > 0) BK::for_each((size_t)256 * L, FoldLatentBack{P.W[3], G.W[3], d3, COL_XIN + SDF_HID, P.latent, G.latent, L, gb}, st);
        if (G.W[3]) {
            BK::gemm(COL_HID, COL_XIN, n, g1, 1, COL_HID, xin, COL_XLD, 1, G.W[3], d3, nullptr, true, st);
            BK::gemm(COL_HID, SDF_HID, n, g1, 1, COL_HID, F, SDF_HID, 1, G.W[3] + COL_XIN, d3, nullptr, true, st);
            BK::gemm(COL_HID, COL_MID, n, g1, 1, COL_HID, h3, COL_MID, 1, G.W[3] + d0, d3, nullptr, true, st);
        }
        BK::zero(g_xin, (size_t)n * COL_XLD, st);
        BK::gemm(n, COL_XIN, COL_HID, g1, COL_HID, 1, P.W[3], d3, 1, g_xin, COL_XLD, nullptr, false, st);
        BK::gemm(n, SDF_HID, COL_HID, g1, COL_HID, 1, P.W[3] + COL_XIN, d3, 1, g_F, SDF_HID, nullptr, false, st);
        BK::gemm(n, COL_MID, COL_HID, g1, COL_HID, 1, P.W[3] + d0, d3, 1, g3, COL_MID, nullptr, false, st);
        r1[0] = G.b[2]; BK::template col_reduce<1>(n, COL_MID, ReluBack{g3, h3, COL_MID}, r1, st);
        if (G.W[2]) BK::gemm(COL_MID, COL_HID, n, g3, 1, COL_MID, h2, COL_HID, 1, G.W[2], COL_HID, nullptr, true, st);
        BK::gemm(n, COL_HID, COL_MID, g3, COL_MID, 1, P.W[2], COL_HID, 1, g0, COL_HID, nullptr, false, st);
        r1[0] = G.b[1]; BK::template col_reduce<1>(n, COL_HID, ReluBack{g0, h2, COL_HID}, r1, st);
        if (G.W[1]) BK::gemm(COL_HID, COL_HID, n, g0, 1, COL_HID, h1, COL_HID, 1, G.W[1], COL_HID, nullptr, true, st);
        BK::gemm(n, COL_HID, COL_HID, g0, COL_HID, 1, P.W[1], COL_HID, 1, g1, COL_HID, nullptr, false, st);
        BK::zero(gb, 256, st);
        r1[0] = gb; BK::template col_reduce<1>(n, COL_HID, ReluBack{g1, h1, COL_HID}, r1, st);
        if (G.b[0]) BK::for_each(256, AddVec{G.b[0], gb}, st);
        if (L > 0) BK::for_each((size_t)256 * L, FoldLatentBack{P.W[0], G.W[0], d0, COL_XIN + SDF_HID, P.latent, G.latent, L, gb}, st);
        if (G.W[0]) {
            BK::gemm(COL_HID, COL_XIN, n, g1, 1, COL_HID, xin, COL_XLD, 1, G.W[0], d0, nullptr, true, st);
            BK::gemm(COL_HID, SDF_HID, n, g1, 1, COL_HID, F, SDF_HID, 1, G.W[0] + COL_XIN, d0, nullptr, true, st);
        }
        BK::gemm(n, COL_XIN, COL_HID, g1, COL_HID, 1, P.W[0], d0, 1, g_xin, COL_XLD, nullptr, true, st);
        BK::gemm(n, SDF_HID, COL_HID, g1, COL_HID, 1, P.W[0] + COL_XIN, d0, 1, g_F, SDF_HID, nullptr, true, st);
    }
};

// =====================================================================================================================
// skinning network (metaavatar/models/decoder.py:201-233 Softplus(beta=100) MLP 3 -> 128 x 4 -> 25, then
// hierarchical_softmax(20 * logits), utils/root_finding_utils.py:98-99)
// =====================================================================================================================
struct SkinWeights {       // per point: w = hierarchical_softmax(20 logits)
    const float* logits; float* w;
    ARAH_HD void operator()(size_t m) const {
        float x[25];
        for (int j = 0; j < 25; ++j) x[j] = logits[m * SKIN_OLD + j] * 20.0f;
        hierarchical_softmax(x, w + m * NJ);
    }
};
struct SkinWeightsBack {   // per point: g_logits = 20 * VJP
    const float* logits; const float* gw; float* gl;
    ARAH_HD void operator()(size_t m) const {
        float x[25], g[25];
        for (int j = 0; j < 25; ++j) x[j] = logits[m * SKIN_OLD + j] * 20.0f;
        hierarchical_softmax_vjp(x, gw + m * NJ, g);
        for (int j = 0; j < 25; ++j) gl[m * SKIN_OLD + j] = g[j] * 20.0f;
        for (int j = 25; j < SKIN_OLD; ++j) gl[m * SKIN_OLD + j] = 0.0f;
    }
};

template <class BK>
struct SkinNet {
    typedef typename BK::Stream Stream;
    int cap = 0, n = 0;
    float* xn = nullptr;          // [n][4] normalised points
    float* pre[4] = {}; float* act[4] = {};     // [n][128]
    float* logits = nullptr;      // [n][28]
    float* w = nullptr;           // [n][24]
    float* tA = nullptr; float* tB = nullptr;   // [3n][128] tangent rows (d/d xn_k of point m at row 3m+k)
    float* tl = nullptr;          // [3n][28]  tangent logits
    float* g0 = nullptr; float* g1 = nullptr; float* gl = nullptr; float* gw = nullptr;   // [n][128] x2, [n][28], [n][24]

    void release() {
        BK::free(xn); for (auto& p : pre) BK::free(p); for (auto& p : act) BK::free(p);
        BK::free(logits); BK::free(w); BK::free(tA); BK::free(tB); BK::free(tl); BK::free(g0); BK::free(g1); BK::free(gl); BK::free(gw);
        *this = SkinNet();
    }
    bool reserve(int n_, bool tangents) {
        if (n_ <= cap && (!tangents || tA)) return true;
        const int keep = n_ > cap ? n_ : cap;
        release();
        const size_t c = ((size_t)keep + 1023) / 1024 * 1024, wd = c * SKIN_HID;
        bool ok = true;
        auto A = [&](float*& p, size_t f) { p = BK::alloc(f); ok = ok && p; };
        A(xn, c * 4); for (auto& p : pre) A(p, wd); for (auto& p : act) A(p, wd); A(logits, c * SKIN_OLD); A(w, c * NJ);
        if (tangents) { A(tA, 3 * wd); A(tB, 3 * wd); A(tl, 3 * c * SKIN_OLD); }
        A(g0, wd); A(g1, wd); A(gl, c * SKIN_OLD); A(gw, c * NJ);
        if (!ok) { release(); return false; }
        cap = (int)c;
        return true;
    }
    void forward(const SkinParams& P, int n_, Stream st) {
        n = n_;
        if (n == 0) return;
        for (int l = 0; l < 4; ++l) {
            const int K = l ? SKIN_HID : 3;
            BK::gemm(n, SKIN_HID, K, l ? act[l - 1] : xn, l ? SKIN_HID : 4, 1, P.W[l], 1, K, pre[l], SKIN_HID, P.b[l], false, st);
            BK::for_each((size_t)n * SKIN_HID, Softplus{pre[l], act[l]}, st);
        }
        BK::zero(logits, (size_t)n * SKIN_OLD, st);
        BK::gemm(n, SKIN_OUT, SKIN_HID, act[3], SKIN_HID, 1, P.W[4], 1, SKIN_HID, logits, SKIN_OLD, P.b[4], false, st);
        BK::for_each(n, SkinWeights{logits, w}, st);
    }
    // forward-mode Jacobian rows of the logits w.r.t. the three normalised coordinates -> tl [3n][28]
    void tangents(const SkinParams& P, Stream st) {
        if (n == 0) return;
        const size_t rows = (size_t)3 * n;
        BK::for_each(rows * SKIN_HID, SkinTangent0{P.W[0], pre[0], tA}, st);
        float* a = tA; float* b = tB;
        for (int l = 1; l < 4; ++l) {
            BK::gemm((int)rows, SKIN_HID, SKIN_HID, a, SKIN_HID, 1, P.W[l], 1, SKIN_HID, b, SKIN_HID, nullptr, false, st);
            BK::for_each(rows * SKIN_HID, SoftplusTangent{b, pre[l], SKIN_HID}, st);
            float* sw = a; a = b; b = sw;
        }
        BK::zero(tl, rows * SKIN_OLD, st);
        BK::gemm((int)rows, SKIN_OUT, SKIN_HID, a, SKIN_HID, 1, P.W[4], 1, SKIN_HID, tl, SKIN_OLD, nullptr, false, st);
    }
    // gw_ [n][24] = d L / d weights
    void backward(const SkinParams& P, const SkinGrads& G, const float* gw_, Stream st) {
        if (n == 0) return;
        BK::for_each(n, SkinWeightsBack{logits, gw_, gl}, st);
        float* r1[1];
        r1[0] = G.b[4]; BK::template col_reduce<1>(n, SKIN_OUT, ColSumMat{gl, SKIN_OLD}, r1, st);
        if (G.W[4]) BK::gemm(SKIN_OUT, SKIN_HID, n, gl, 1, SKIN_OLD, act[3], SKIN_HID, 1, G.W[4], SKIN_HID, nullptr, true, st);
        float* g = g0; float* gn = g1;
        BK::gemm(n, SKIN_HID, SKIN_OUT, gl, SKIN_OLD, 1, P.W[4], SKIN_HID, 1, g, SKIN_HID, nullptr, false, st);
        for (int l = 3; l >= 0; --l) {
            const int K = l ? SKIN_HID : 3;
            r1[0] = G.b[l]; BK::template col_reduce<1>(n, SKIN_HID, SoftplusBack{g, pre[l], SKIN_HID}, r1, st);
            if (G.W[l]) BK::gemm(SKIN_HID, K, n, g, 1, SKIN_HID, l ? act[l - 1] : xn, l ? SKIN_HID : 4, 1, G.W[l], K, nullptr, true, st);
            if (l > 0) { BK::gemm(n, SKIN_HID, SKIN_HID, g, SKIN_HID, 1, P.W[l], SKIN_HID, 1, gn, SKIN_HID, nullptr, false, st); float* sw = g; g = gn; gn = sw; }
        }
    }
};

// =====================================================================================================================
// per-sample / per-ray glue of the shading pass
// =====================================================================================================================
struct ShadeGeom {
    int P, S, cano_view_dirs, ray_augm;
    const int* list;              // [M] sample slot (ray * S + i) of the m-th converged sample
    const float* smp_xn;          // [P*S][3]  normalised canonical points (tracer output)
    const float* smp_T;           // [P*S][12] forward transforms (3x4)
    const float* z_vals;          // [P*S]
    const uint8_t* smp_conv;      // [P*S]
    const float* view;            // [P][3] ray directions after view augmentation (the kernel negates them)
    const float* view_orig;       // [P][3] un-augmented ray directions
    float sdf_scale;              // cmax - cmin; raw sdf -> metres is s / 2 * 1.1 * (cmax - cmin)   (:359)
    float beta_raw;               // ||variance||
};

struct GatherPoints {   // x[m] = smp_xn[list[m]]
    ShadeGeom g; float* x;
    ARAH_HD void operator()(size_t m) const {
        const size_t sl = (size_t)g.list[m];
        x[4 * m] = g.smp_xn[3 * sl]; x[4 * m + 1] = g.smp_xn[3 * sl + 1]; x[4 * m + 2] = g.smp_xn[3 * sl + 2]; x[4 * m + 3] = 0.0f;
    }
};
// colour-net inputs [xn | PE4(view) | normal]   (implicit_differentiable_renderer.py:293-301,336-350; embedder.py)
struct ColInputs {
    ShadeGeom g; const float* x; const float* nrm; float* xin;
    ARAH_HD void operator()(size_t m) const {
        const size_t sl = (size_t)g.list[m];
        const int r = (int)(sl / g.S);
        const float* T = g.smp_T + 12 * sl;
        float v[3], vo[3], nn[3];
        const float n0 = nrm[4 * m], n1 = nrm[4 * m + 1], n2 = nrm[4 * m + 2];
        if (g.cano_view_dirs) {
            float A[9], Ai[9];
            for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) A[a * 3 + c] = T[a * 4 + c];
            invert3(A, Ai);
            for (int a = 0; a < 3; ++a) {
                v[a] = Ai[a * 3] * -g.view[3 * r] + Ai[a * 3 + 1] * -g.view[3 * r + 1] + Ai[a * 3 + 2] * -g.view[3 * r + 2];
                vo[a] = Ai[a * 3] * -g.view_orig[3 * r] + Ai[a * 3 + 1] * -g.view_orig[3 * r + 1] + Ai[a * 3 + 2] * -g.view_orig[3 * r + 2];
            }
            nn[0] = n0; nn[1] = n1; nn[2] = n2;
        } else {
            for (int a = 0; a < 3; ++a) { v[a] = -g.view[3 * r + a]; vo[a] = -g.view_orig[3 * r + a]; nn[a] = T[a * 4] * n0 + T[a * 4 + 1] * n1 + T[a * 4 + 2] * n2; }
        }
        if (g.ray_augm) {      // :342-350 views that end up behind the surface fall back to the un-augmented direction
            const float nl = sqrtf(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
            const float d = (nn[0] / nl) * v[0] + (nn[1] / nl) * v[1] + (nn[2] / nl) * v[2];
            if (acosf(d) >= 1.5707963267948966f) { v[0] = vo[0]; v[1] = vo[1]; v[2] = vo[2]; }
        }
        float* o = xin + m * COL_XLD;
        o[0] = x[4 * m]; o[1] = x[4 * m + 1]; o[2] = x[4 * m + 2];
        o[3] = v[0]; o[4] = v[1]; o[5] = v[2];
        float fr = 1.0f;
        for (int k = 0; k < 4; ++k) {
            for (int a = 0; a < 3; ++a) { o[6 + 6 * k + a] = sinf(v[a] * fr); o[9 + 6 * k + a] = cosf(v[a] * fr); }
            fr *= 2.0f;
        }
        o[30] = nn[0]; o[31] = nn[1]; o[32] = nn[2]; o[33] = 0.0f; o[34] = 0.0f; o[35] = 0.0f;
    }
};
struct ScatterShaded {   // per-slot colour and sdf in metres
    ShadeGeom g; const float* o; const float* s; float* smp_rgb; float* smp_sdf;
    ARAH_HD void operator()(size_t m) const {
        const size_t sl = (size_t)g.list[m];
        smp_rgb[3 * sl] = o[4 * m]; smp_rgb[3 * sl + 1] = o[4 * m + 1]; smp_rgb[3 * sl + 2] = o[4 * m + 2];
        smp_sdf[sl] = s[m] / 2.0f * 1.1f * g.sdf_scale;
    }
};
// one ray: sigma-from-SDF, alpha compositing (implicit_differentiable_renderer.py:366-394); keeps alpha / transmittance / dz
// per slot for the backward sweep
struct CompositeFwd {
    ShadeGeom g; const float* smp_rgb; const float* smp_sdf; float* alpha; float* trans; float* dzs; float* rgb; float* wsum; float* wraw;
    ARAH_HD void operator()(size_t r) const {
        const int S = g.S;
        float beta = fminf(fmaxf(fabsf(g.beta_raw), 1e-6f), 1e6f);
        const float ib = 1.0f / beta;
        float Tr = 1.0f, a0 = 0.f, a1 = 0.f, a2 = 0.f, aw = 0.f;
        int prev = -1;
        float prev_sig = 0.f;
        // a sample's interval ends at the next converged sample; the last one gets 1 / n_steps
        for (int i = 0; i <= S; ++i) {
            const bool valid = (i < S) && g.smp_conv[r * S + i];
            if (!valid && i < S) continue;
            if (prev >= 0) {
                const size_t sp = r * S + prev;
                const float dz = (i < S) ? (g.z_vals[r * S + i] - g.z_vals[sp]) : (1.0f / (float)S);
                const float al = 1.0f - expf(-prev_sig * dz);
                const float wgt = al * Tr;
                alpha[sp] = al; trans[sp] = Tr; dzs[sp] = dz;
                a0 += wgt * smp_rgb[3 * sp]; a1 += wgt * smp_rgb[3 * sp + 1]; a2 += wgt * smp_rgb[3 * sp + 2]; aw += wgt;
                Tr = Tr * (1.0f - al + 1e-7f);
            }
            if (i < S) { prev = i; prev_sig = laplace_density(smp_sdf[r * S + i], ib); }
        }
        rgb[3 * r] = a0; rgb[3 * r + 1] = a1; rgb[3 * r + 2] = a2;
        wraw[r] = aw;
        wsum[r] = fminf(fmaxf(aw, 0.0f), 1.0f);
    }
};
// one ray, reverse sweep: d L / d (per-slot colour), d L / d (per-slot raw sdf), d L / d beta (atomic)
struct CompositeBwd {
    ShadeGeom g; const float* smp_rgb; const float* smp_sdf; const float* alpha; const float* trans; const float* dzs; const float* wraw;
    const float* g_rgb; const float* g_wsum; float* g_smp_rgb; float* g_smp_s; float* g_beta;
    ARAH_HD void operator()(size_t r) const {
        const int S = g.S;
        const float braw = fabsf(g.beta_raw);
        const float beta = fminf(fmaxf(braw, 1e-6f), 1e6f);
        const bool beta_free = braw >= 1e-6f && braw <= 1e6f;
        const float ib = 1.0f / beta;
        const float gr0 = g_rgb[3 * r], gr1 = g_rgb[3 * r + 1], gr2 = g_rgb[3 * r + 2];
        const float gws = (g_wsum && wraw[r] >= 0.0f && wraw[r] <= 1.0f) ? g_wsum[r] : 0.0f;
        float gT_next = 0.0f, g_ib = 0.0f;       // d L / d T_{k+1}
        for (int i = S - 1; i >= 0; --i) {
            const size_t sl = r * S + i;
            if (!g.smp_conv[sl]) continue;
            const float al = alpha[sl], Tk = trans[sl], dz = dzs[sl];
            const float c0 = smp_rgb[3 * sl], c1 = smp_rgb[3 * sl + 1], c2 = smp_rgb[3 * sl + 2];
            const float gw = gr0 * c0 + gr1 * c1 + gr2 * c2 + gws;
            const float wgt = al * Tk;
            g_smp_rgb[3 * sl] = wgt * gr0; g_smp_rgb[3 * sl + 1] = wgt * gr1; g_smp_rgb[3 * sl + 2] = wgt * gr2;
            const float g_al = gw * Tk - gT_next * Tk;
            gT_next = gw * al + gT_next * (1.0f - al + 1e-7f);
            // alpha = 1 - exp(-sigma dz)
            const float sm = smp_sdf[sl];
            const float sig = laplace_density(sm, ib);
            const float g_sig = g_al * dz * expf(-sig * dz);
            // sigma = relu(ib (0.5 + 0.5 sign(-s) (1 - exp(-|s| ib))))
            const float e = expf(-fabsf(sm) * ib);
            const float sg = (-sm > 0.0f) ? 1.0f : ((-sm < 0.0f) ? -1.0f : 0.0f);
            const float q = 0.5f + 0.5f * sg * (1.0f - e);
            float g_sm = 0.0f;
            if (ib * q > 0.0f) {
                g_sm = g_sig * (-0.5f * ib * ib * e) * ((sm != 0.0f) ? 1.0f : 0.0f);
                g_ib += g_sig * (q + ib * 0.5f * sg * e * fabsf(sm));
            }
            g_smp_s[sl] = g_sm / 2.0f * 1.1f * g.sdf_scale;
        }
        if (g_beta && beta_free && g_ib != 0.0f) atomic_addf(g_beta, -g_ib * ib * ib);
    }
};
struct GatherShadeGrads {    // per converged sample: g_o [m][4], g_s [m]
    ShadeGeom g; const float* g_smp_rgb; const float* g_smp_s; float* go; float* gs;
    ARAH_HD void operator()(size_t m) const {
        const size_t sl = (size_t)g.list[m];
        go[4 * m] = g_smp_rgb[3 * sl]; go[4 * m + 1] = g_smp_rgb[3 * sl + 1]; go[4 * m + 2] = g_smp_rgb[3 * sl + 2]; go[4 * m + 3] = 0.0f;
        gs[m] = g_smp_s[sl];
    }
};
struct SplitXinGrad {        // g_xin -> g_normal (canonical space), g_x (direct colour-net dependence on the point)
    ShadeGeom g; const float* gxin; float* gn; float* gx;
    ARAH_HD void operator()(size_t m) const {
        const float* q = gxin + m * COL_XLD;
        const float a0 = q[30], a1 = q[31], a2 = q[32];
        if (g.cano_view_dirs) { gn[4 * m] = a0; gn[4 * m + 1] = a1; gn[4 * m + 2] = a2; }
        else {
            const float* T = g.smp_T + 12 * (size_t)g.list[m];
            for (int c = 0; c < 3; ++c) gn[4 * m + c] = T[c] * a0 + T[4 + c] * a1 + T[8 + c] * a2;      // A^T g
        }
        gn[4 * m + 3] = 0.0f;
        gx[4 * m] = q[0]; gx[4 * m + 1] = q[1]; gx[4 * m + 2] = q[2]; gx[4 * m + 3] = 0.0f;
    }
};
// skinning-net input of the implicit-gradient correction: normalise(unnormalise(pi))  (:315-329)
struct SkinInputs {
    Norm nm; const float* x; float* xn;
    ARAH_HD void operator()(size_t m) const {
        for (int k = 0; k < 3; ++k) {
            const float xh = unnormalize1(x[4 * m + k], nm.center[k], nm.cmin, nm.cmax);
            xn[4 * m + k] = normalize1(xh, nm.center[k], nm.cmin, nm.cmax);
        }
        xn[4 * m + 3] = 0.0f;
    }
};
struct SkinInputsMetres {    // query_weights on points given in metres (root_finding_utils.py:79)
    Norm nm; const float* xh; float* xn;
    ARAH_HD void operator()(size_t m) const {
        for (int k = 0; k < 3; ++k) xn[4 * m + k] = normalize1(xh[3 * m + k], nm.center[k], nm.cmin, nm.cmax);
        xn[4 * m + 3] = 0.0f;
    }
};
// pi' = pi - J^-1 (lbs(pi) - lbs(pi).detach())  =>  d L / d lbs = -J^-T g_pi ;  d L / d w_j = g_lbs . (B_j [x_hat;1])
// with J = d lbs / d pi (full Jacobian including d w / d pi), implicit_differentiable_renderer.py:315-334
struct ImplicitSkinGrad {
    Norm nm; const float* bone_T; const float* x; const float* logits; const float* tl; const float* gx; float* gw;
    ARAH_HD void operator()(size_t m) const {
        float xh[3];
        for (int k = 0; k < 3; ++k) xh[k] = unnormalize1(x[4 * m + k], nm.center[k], nm.cmin, nm.cmax);
        const float su = 1.1f * (nm.cmax - nm.cmin) / 2.0f;        // d x_hat / d pi
        Dual3 lg[25], wd[NJ];
        for (int j = 0; j < 25; ++j) {
            lg[j].v = logits[m * SKIN_OLD + j] * 20.0f;
            for (int k = 0; k < 3; ++k) lg[j].d[k] = tl[(3 * m + k) * SKIN_OLD + j] * 20.0f;
        }
        hierarchical_softmax_dual(lg, wd);
        float J[9], Ji[9], Bx[NJ][3];
        for (int e = 0; e < 9; ++e) J[e] = 0.0f;
        for (int j = 0; j < NJ; ++j) {
            const float* B = bone_T + 16 * j;
            for (int a = 0; a < 3; ++a) Bx[j][a] = B[a * 4] * xh[0] + B[a * 4 + 1] * xh[1] + B[a * 4 + 2] * xh[2] + B[a * 4 + 3];
            for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) J[a * 3 + c] += Bx[j][a] * wd[j].d[c] + wd[j].v * B[a * 4 + c] * su;
        }
        float gl[3] = {0.f, 0.f, 0.f};
        if (invert3(J, Ji)) for (int a = 0; a < 3; ++a) gl[a] = -(Ji[0 * 3 + a] * gx[4 * m] + Ji[1 * 3 + a] * gx[4 * m + 1] + Ji[2 * 3 + a] * gx[4 * m + 2]);
        for (int j = 0; j < NJ; ++j) gw[m * NJ + j] = gl[0] * Bx[j][0] + gl[1] * Bx[j][1] + gl[2] * Bx[j][2];
    }
};
struct Pad3to4 { const float* src; float* dst; ARAH_HD void operator()(size_t m) const { dst[4 * m] = src[3 * m]; dst[4 * m + 1] = src[3 * m + 1]; dst[4 * m + 2] = src[3 * m + 2]; dst[4 * m + 3] = 0.0f; } };
struct Unpad4to3 { const float* src; float* dst; ARAH_HD void operator()(size_t m) const { dst[3 * m] = src[4 * m]; dst[3 * m + 1] = src[4 * m + 1]; dst[3 * m + 2] = src[4 * m + 2]; } };
struct CopyF { const float* src; float* dst; ARAH_HD void operator()(size_t i) const { dst[i] = src[i]; } };

// =====================================================================================================================
// one training step's worth of state: shading pass + auxiliary evaluations
// =====================================================================================================================
struct AllParams { SdfParams sdf; SkinParams skin; ColParams col; const float* bone_T; Norm nm; };
struct AllGrads { SdfGrads sdf; SkinGrads skin; ColGrads col; float* beta; };

template <class BK>
struct Session {
    typedef typename BK::Stream Stream;
    static constexpr int N_AUX = 3;
    SdfNet<BK> sdf;               // shading samples
    ColNet<BK> col;
    SkinNet<BK> skin;             // implicit-gradient correction (filled during backward)
    SdfNet<BK> aux_sdf[N_AUX];    // eikonal + off-surface points, inside points, spare
    SkinNet<BK> aux_skin;         // points_skinning
    // per-slot buffers [P*S]
    size_t slot_cap = 0;
    float* smp_rgb = nullptr; float* smp_sdf = nullptr; float* alpha = nullptr; float* trans = nullptr; float* dzs = nullptr;
    float* g_smp_rgb = nullptr; float* g_smp_s = nullptr;
    float* wraw = nullptr;        // [P]
    // per-sample scratch
    int samp_cap = 0;
    float* go = nullptr; float* gs = nullptr; float* gxin = nullptr; float* gF = nullptr; float* gn = nullptr; float* gx = nullptr;
    ShadeGeom geom{};
    int M = 0;
    bool train_skin = false;

    void release() {
        sdf.release(); col.release(); skin.release(); for (auto& a : aux_sdf) a.release(); aux_skin.release();
        BK::free(smp_rgb); BK::free(smp_sdf); BK::free(alpha); BK::free(trans); BK::free(dzs); BK::free(g_smp_rgb); BK::free(g_smp_s); BK::free(wraw);
        BK::free(go); BK::free(gs); BK::free(gxin); BK::free(gF); BK::free(gn); BK::free(gx);
        smp_rgb = smp_sdf = alpha = trans = dzs = g_smp_rgb = g_smp_s = wraw = go = gs = gxin = gF = gn = gx = nullptr;
        slot_cap = 0; samp_cap = 0;
    }
    bool reserve(size_t slots, int P, int M_) {
        bool ok = true;
        auto A = [&](float*& p, size_t f) { p = BK::alloc(f); ok = ok && p; };
        if (slots > slot_cap) {
            BK::free(smp_rgb); BK::free(smp_sdf); BK::free(alpha); BK::free(trans); BK::free(dzs); BK::free(g_smp_rgb); BK::free(g_smp_s); BK::free(wraw);
            A(smp_rgb, slots * 3); A(smp_sdf, slots); A(alpha, slots); A(trans, slots); A(dzs, slots); A(g_smp_rgb, slots * 3); A(g_smp_s, slots);
            A(wraw, (size_t)P + 1);
            slot_cap = ok ? slots : 0;
        }
        if (M_ > samp_cap) {
            BK::free(go); BK::free(gs); BK::free(gxin); BK::free(gF); BK::free(gn); BK::free(gx);
            const size_t c = ((size_t)M_ + 1023) / 1024 * 1024;
            A(go, c * 4); A(gs, c); A(gxin, c * COL_XLD); A(gF, c * SDF_HID); A(gn, c * 4); A(gx, c * 4);
            samp_cap = ok ? (int)c : 0;
        }
        return ok && sdf.reserve(M_) && col.reserve(M_);
    }

    // ---- IDHRNetwork.forward, training branch, the differentiable part (implicit_differentiable_renderer.py:150-178)
    // rgb [P][3], wsum [P] (0 for rays without a converged sample)
    int shade_forward(const AllParams& A, const ShadeGeom& g, int M_, bool train_skinning_net, float* rgb, float* wsum, Stream st) {
        geom = g; M = M_; train_skin = train_skinning_net;
        const size_t slots = (size_t)g.P * g.S;
        if (!reserve(slots, g.P, M)) return -1;
        if (M > 0) {
            BK::for_each(M, GatherPoints{g, sdf.x}, st);
            sdf.forward(A.sdf, M, true, st);
            BK::for_each(M, ColInputs{g, sdf.x, sdf.nrm, col.xin}, st);
            col.forward(A.col, sdf.h[5], M, st);
            BK::for_each(M, ScatterShaded{g, col.o, sdf.s, smp_rgb, smp_sdf}, st);
        } else { sdf.n = 0; col.n = 0; }
        BK::for_each(g.P, CompositeFwd{g, smp_rgb, smp_sdf, alpha, trans, dzs, rgb, wsum, wraw}, st);
        return 0;
    }
    int shade_backward(const AllParams& A, const AllGrads& G, const float* g_rgb, const float* g_wsum, Stream st) {
        const ShadeGeom& g = geom;
        if (M == 0) return 0;
        BK::for_each(g.P, CompositeBwd{g, smp_rgb, smp_sdf, alpha, trans, dzs, wraw, g_rgb, g_wsum, g_smp_rgb, g_smp_s, G.beta}, st);
        BK::for_each(M, GatherShadeGrads{g, g_smp_rgb, g_smp_s, go, gs}, st);
        col.backward(A.col, G.col, go, gxin, gF, st);
        BK::for_each(M, SplitXinGrad{g, gxin, gn, gx}, st);
        sdf.backward(A.sdf, G.sdf, gs, gF, SDF_HID, gn, gx, st);
        if (train_skin) {
            if (!skin.reserve(M, true)) return -1;
            BK::for_each(M, SkinInputs{A.nm, sdf.x, skin.xn}, st);
            skin.forward(A.skin, M, st);
            skin.tangents(A.skin, st);
            BK::for_each(M, ImplicitSkinGrad{A.nm, A.bone_T, sdf.x, skin.logits, skin.tl, gx, skin.gw}, st);
            skin.backward(A.skin, G.skin, skin.gw, st);
        }
        return 0;
    }
    // ---- auxiliary SDF evaluation: points [n][3] normalised -> sdf [n] (raw network output), grad [n][3] (optional)
    int sdf_forward(const AllParams& A, int slot, const float* pts, int n, bool with_grad, float* out_sdf, float* out_grad, Stream st) {
        if (slot < 0 || slot >= N_AUX) return -2;
        SdfNet<BK>& s = aux_sdf[slot];
        if (!s.reserve(n)) return -1;
        BK::for_each(n, Pad3to4{pts, s.x}, st);
        s.forward(A.sdf, n, with_grad, st);
        if (n == 0) return 0;
        BK::for_each(n, CopyF{s.s, out_sdf}, st);
        if (with_grad && out_grad) BK::for_each(n, Unpad4to3{s.nrm, out_grad}, st);
        return 0;
    }
    int sdf_backward(const AllParams& A, const AllGrads& G, int slot, const float* g_sdf, const float* g_grad, Stream st) {
        if (slot < 0 || slot >= N_AUX) return -2;
        SdfNet<BK>& s = aux_sdf[slot];
        if (s.n == 0) return 0;
        const float* gnp = nullptr;
        if (g_grad && s.has_normal) { BK::for_each(s.n, Pad3to4{g_grad, s.u0}, st); gnp = s.u0; }
        s.backward(A.sdf, G.sdf, g_sdf, nullptr, 0, gnp, nullptr, st);
        return 0;
    }
    // ---- query_weights(points in metres) -> [n][24]
    int skin_forward(const AllParams& A, const float* pts_m, int n, float* out_w, Stream st) {
        if (!aux_skin.reserve(n, false)) return -1;
        BK::for_each(n, SkinInputsMetres{A.nm, pts_m, aux_skin.xn}, st);
        aux_skin.forward(A.skin, n, st);
        if (n) BK::for_each((size_t)n * NJ, CopyF{aux_skin.w, out_w}, st);
        return 0;
    }
    int skin_backward(const AllParams& A, const AllGrads& G, const float* g_w, Stream st) {
        aux_skin.backward(A.skin, G.skin, g_w, st);
        return 0;
    }
};

}  // namespace train
}  // namespace arah
