// arah_math.cuh — per-point scalar math of the ARAH hot path (host+device).
//
// Everything here is fp32 and branch-compatible with the reference's batched-mask code:
//   normalisation        /root/reference/im2mesh/utils/root_finding_utils.py:37-51
//   hierarchical softmax /root/reference/im2mesh/utils/utils.py:138-181
//   LBS blend / apply    /root/reference/im2mesh/utils/root_finding_utils.py:13-33
//   Broyden step         /root/reference/im2mesh/utils/broyden.py:47-76
// The functions are __host__ __device__ so that tests/test_host_math.py can exercise the exact code the
// kernels run on the CPU (compiled with g++ through csrc/host_math_test.cpp) without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ARAH_HD __host__ __device__ __forceinline__
#else
#define ARAH_HD inline
#endif

namespace arah {

constexpr int NJ = 24;            // SMPL joints
constexpr float CVG_THRESH = 1e-5f;
constexpr float DVG_THRESH = 1.0f;
constexpr float BROYDEN_EPS = 1e-6f;

struct NormParams { float cmin, cmax, center[3]; };

ARAH_HD float normalize1(float p, float center, float cmin, float cmax) {
    const float d = cmax - cmin, padding = d * 0.05f;
    float v = p - center;
    v = (v - cmin + padding) / d / 1.1f;
    v = v - 0.5f;
    return v * 2.0f;
}
ARAH_HD float unnormalize1(float p, float center, float cmin, float cmax) {
    const float d = cmax - cmin, padding = d * 0.05f;
    return (p / 2.0f + 0.5f) * 1.1f * d + cmin - padding + center;
}
ARAH_HD float sdf_to_metres(float s, float cmin, float cmax) { return s / 2.0f * 1.1f * (cmax - cmin); }

ARAH_HD float sigmoid_(float x) { return 1.0f / (1.0f + expf(-x)); }
// torch.nn.Softplus(beta=100, threshold=20) and its derivative
ARAH_HD float softplus100(float x) {
    const float bx = x * 100.0f;
    return bx > 20.0f ? x : log1pf(expf(bx)) / 100.0f;
}
ARAH_HD float softplus100_grad(float x) {
    const float bx = x * 100.0f;
    return bx > 20.0f ? 1.0f : sigmoid_(bx);
}

ARAH_HD void softmax3(const float* x, float* y) {
    const float m = fmaxf(x[0], fmaxf(x[1], x[2]));
    const float e0 = expf(x[0] - m), e1 = expf(x[1] - m), e2 = expf(x[2] - m), s = e0 + e1 + e2;
    y[0] = e0 / s; y[1] = e1 / s; y[2] = e2 / s;
}

// x: 25 logits ALREADY multiplied by 20 (root_finding_utils.py:99) -> p: 24 weights
ARAH_HD void hierarchical_softmax(const float* x, float* p) {
    float sm[3];
    softmax3(x + 1, sm);
    const float s0 = sigmoid_(x[0]);
    p[0] = 1.0f;
    for (int k = 0; k < 3; ++k) p[1 + k] = p[0] * s0 * sm[k];
    p[0] = p[0] * (1.0f - s0);
#define ARAH_SPLIT(c, q, g) { const float s_ = sigmoid_(x[g]); p[c] = p[q] * s_; p[q] = p[q] * (1.0f - s_); }
    ARAH_SPLIT(4, 1, 4) ARAH_SPLIT(5, 2, 5) ARAH_SPLIT(6, 3, 6)
    ARAH_SPLIT(7, 4, 7) ARAH_SPLIT(8, 5, 8) ARAH_SPLIT(9, 6, 9)
    ARAH_SPLIT(10, 7, 10) ARAH_SPLIT(11, 8, 11)
    softmax3(x + 12, sm);
    {
        const float s24 = sigmoid_(x[24]), p9 = p[9];
        for (int k = 0; k < 3; ++k) p[12 + k] = p9 * s24 * sm[k];
        p[9] = p9 * (1.0f - s24);
    }
    ARAH_SPLIT(15, 12, 15)
    ARAH_SPLIT(16, 13, 16) ARAH_SPLIT(17, 14, 17)
    ARAH_SPLIT(18, 16, 18) ARAH_SPLIT(19, 17, 19)
    ARAH_SPLIT(20, 18, 20) ARAH_SPLIT(21, 19, 21)
    ARAH_SPLIT(22, 20, 22) ARAH_SPLIT(23, 21, 23)
#undef ARAH_SPLIT
}

// forward-mode (3 tangents) version, used once per ray for the full LBS Jacobian
// (forward_skinning_jac, root_finding_utils.py:170-230; the reference uses 3 autograd VJPs)
struct Dual3 { float v, d[3]; };
ARAH_HD Dual3 dmul(const Dual3& a, const Dual3& b) {
    Dual3 r; r.v = a.v * b.v;
    for (int k = 0; k < 3; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
    return r;
}
ARAH_HD Dual3 dcompl(const Dual3& a) { Dual3 r; r.v = 1.0f - a.v; for (int k = 0; k < 3; ++k) r.d[k] = -a.d[k]; return r; }
ARAH_HD Dual3 dsigmoid(const Dual3& a) {
    Dual3 r; r.v = sigmoid_(a.v);
    const float s = r.v * (1.0f - r.v);
    for (int k = 0; k < 3; ++k) r.d[k] = s * a.d[k];
    return r;
}
ARAH_HD void dsoftmax3(const Dual3* x, Dual3* y) {
    const float v[3] = {x[0].v, x[1].v, x[2].v};
    float s[3];
    softmax3(v, s);
    for (int i = 0; i < 3; ++i) {
        y[i].v = s[i];
        for (int k = 0; k < 3; ++k) {
            const float dot = s[0] * x[0].d[k] + s[1] * x[1].d[k] + s[2] * x[2].d[k];
            y[i].d[k] = s[i] * (x[i].d[k] - dot);
        }
    }
}
ARAH_HD void hierarchical_softmax_dual(const Dual3* x, Dual3* p) {
    Dual3 sm[3];
    dsoftmax3(x + 1, sm);
    const Dual3 s0 = dsigmoid(x[0]);
    Dual3 one; one.v = 1.0f; one.d[0] = one.d[1] = one.d[2] = 0.0f;
    p[0] = one;
    for (int k = 0; k < 3; ++k) p[1 + k] = dmul(dmul(p[0], s0), sm[k]);
    p[0] = dmul(p[0], dcompl(s0));
#define ARAH_SPLIT(c, q, g) { const Dual3 s_ = dsigmoid(x[g]); p[c] = dmul(p[q], s_); p[q] = dmul(p[q], dcompl(s_)); }
    ARAH_SPLIT(4, 1, 4) ARAH_SPLIT(5, 2, 5) ARAH_SPLIT(6, 3, 6)
    ARAH_SPLIT(7, 4, 7) ARAH_SPLIT(8, 5, 8) ARAH_SPLIT(9, 6, 9)
    ARAH_SPLIT(10, 7, 10) ARAH_SPLIT(11, 8, 11)
    dsoftmax3(x + 12, sm);
    {
        const Dual3 s24 = dsigmoid(x[24]), p9 = p[9];
        for (int k = 0; k < 3; ++k) p[12 + k] = dmul(dmul(p9, s24), sm[k]);
        p[9] = dmul(p9, dcompl(s24));
    }
    ARAH_SPLIT(15, 12, 15)
    ARAH_SPLIT(16, 13, 16) ARAH_SPLIT(17, 14, 17)
    ARAH_SPLIT(18, 16, 18) ARAH_SPLIT(19, 17, 19)
    ARAH_SPLIT(20, 18, 20) ARAH_SPLIT(21, 19, 21)
    ARAH_SPLIT(22, 20, 22) ARAH_SPLIT(23, 21, 23)
#undef ARAH_SPLIT
}

// T (3x4 affine, row-major 12 floats) = sum_j w_j B_j[:3,:]; B is [24][16] row-major 4x4.
// s = sum_j w_j B_j[3][3] (the homogeneous entry; 1 up to rounding).
ARAH_HD void blend_T(const float* w, const float* B, float* T12, float* s) {
    float acc[12];
    for (int e = 0; e < 12; ++e) acc[e] = 0.0f;
    float ss = 0.0f;
    for (int j = 0; j < NJ; ++j) {
        const float wj = w[j];
        for (int e = 0; e < 12; ++e) acc[e] += wj * B[j * 16 + e];
        ss += wj * B[j * 16 + 15];
    }
    for (int e = 0; e < 12; ++e) T12[e] = acc[e];
    if (s) *s = ss;
}
ARAH_HD void apply_T(const float* T12, const float* x, float* y) {
    for (int r = 0; r < 3; ++r) y[r] = T12[r * 4 + 0] * x[0] + T12[r * 4 + 1] * x[1] + T12[r * 4 + 2] * x[2] + T12[r * 4 + 3];
}

// 3x3 inverse by adjugate (row-major). Returns false if singular.
ARAH_HD bool invert3(const float* A, float* Ai) {
    const float a = A[0], b = A[1], c = A[2], d = A[3], e = A[4], f = A[5], g = A[6], h = A[7], i = A[8];
    const float c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
    const float det = a * c00 + b * c01 + c * c02;
    if (det == 0.0f) { for (int k = 0; k < 9; ++k) Ai[k] = 0.0f; return false; }
    const float inv = 1.0f / det;
    Ai[0] = c00 * inv; Ai[1] = (c * h - b * i) * inv; Ai[2] = (b * f - c * e) * inv;
    Ai[3] = c01 * inv; Ai[4] = (a * i - c * g) * inv; Ai[5] = (c * d - a * f) * inv;
    Ai[6] = c02 * inv; Ai[7] = (b * g - a * h) * inv; Ai[8] = (a * e - b * d) * inv;
    return true;
}
// x_hat = (T^-1 [x;1])[:3] for the 4x4 T = [[A, t],[0 0 0 s]]  ->  A^-1 x - A^-1 t / s
// (role of torch.inverse + matmul at ray_tracing.py:393-397, 416-420)
ARAH_HD void affine_inverse_apply(const float* T12, float s, const float* x, float* xh) {
    float A[9], Ai[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A[r * 3 + c] = T12[r * 4 + c];
    invert3(A, Ai);
    const float is = 1.0f / s;
    const float u[3] = {x[0] - T12[3] * is, x[1] - T12[7] * is, x[2] - T12[11] * is};
    for (int r = 0; r < 3; ++r) xh[r] = Ai[r * 3] * u[0] + Ai[r * 3 + 1] * u[1] + Ai[r * 3 + 2] * u[2];
}

// general small inverse (Gauss-Jordan, partial pivoting), N = 3 or 4; used for the 4x4 iso-search Jacobian
template <int N>
ARAH_HD bool invert_gj(const float* A, float* Ai) {
    float M[N][2 * N];
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) { M[r][c] = A[r * N + c]; M[r][N + c] = (r == c) ? 1.0f : 0.0f; }
    for (int c = 0; c < N; ++c) {
        int piv = c;
        for (int r = c + 1; r < N; ++r) if (fabsf(M[r][c]) > fabsf(M[piv][c])) piv = r;
        if (M[piv][c] == 0.0f) { for (int k = 0; k < N * N; ++k) Ai[k] = 0.0f; return false; }
        if (piv != c) for (int k = 0; k < 2 * N; ++k) { const float t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
        const float inv = 1.0f / M[c][c];
        for (int k = 0; k < 2 * N; ++k) M[c][k] *= inv;
        for (int r = 0; r < N; ++r) if (r != c) {
            const float f = M[r][c];
            for (int k = 0; k < 2 * N; ++k) M[r][k] -= f * M[c][k];
        }
    }
    for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) Ai[r * N + c] = M[r][N + c];
    return true;
}

// ---------------------------------------------------------------------------------------------
// Broyden state of one point (D = 3: canonical correspondence, D = 4: joint iso-surface search).
// Lives in HBM between the per-iteration launches (AoS record, 16-byte aligned), in registers inside one.
template <int D>
struct alignas(16) BroydenState {
    float x[D];
    float Jinv[D * D];
    float gx[D];
    float upd[D];
    float best_x[D];
    float best_T[12];
    float best_n;
    int32_t owner;          // ray index (D=4) or sample slot index ray*S+slot (D=3)
    float tgt[3];           // D=3: posed target (x - trans); D=4: unused
    int32_t g_evals;
};

// bookkeeping right after the initial evaluation g(x0) (broyden.py:36-45); T_init is the kNN transform.
template <int D>
ARAH_HD void broyden_begin(BroydenState<D>& s, const float* x0, const float* gx0, const float* Jinv0, const float* T_init12) {
    for (int i = 0; i < D; ++i) { s.x[i] = x0[i]; s.gx[i] = gx0[i]; s.best_x[i] = x0[i]; }
    for (int i = 0; i < D * D; ++i) s.Jinv[i] = Jinv0[i];
    for (int i = 0; i < 12; ++i) s.best_T[i] = T_init12[i];
    float n = 0.0f;
    for (int i = 0; i < D; ++i) n += gx0[i] * gx0[i];
    s.best_n = sqrtf(n);
    for (int r = 0; r < D; ++r) { float a = 0.0f; for (int c = 0; c < D; ++c) a += s.Jinv[r * D + c] * s.gx[c]; s.upd[r] = -a; }
    s.g_evals = 1;
}
// first half of an iteration: x += update (broyden.py:50-51).  dx is kept by the caller for the second half.
template <int D>
ARAH_HD void broyden_advance(BroydenState<D>& s, float* dx) {
    for (int i = 0; i < D; ++i) { dx[i] = s.upd[i]; s.x[i] += dx[i]; }
}
// second half given g(x) and its transform (broyden.py:52-76).  Returns true if the point stays active.
template <int D>
ARAH_HD bool broyden_update(BroydenState<D>& s, const float* dx, const float* g_new, const float* T12) {
    float dg[D];
    for (int i = 0; i < D; ++i) { dg[i] = g_new[i] - s.gx[i]; s.gx[i] += dg[i]; }
    float n = 0.0f;
    for (int i = 0; i < D; ++i) n += s.gx[i] * s.gx[i];
    const float cur = sqrtf(n);
    s.g_evals += 1;
    if (cur < s.best_n) {
        s.best_n = cur;
        for (int i = 0; i < D; ++i) s.best_x[i] = s.x[i];
        for (int i = 0; i < 12; ++i) s.best_T[i] = T12[i];
    }
    if (!(s.best_n > CVG_THRESH && cur < DVG_THRESH)) return false;
    float vT[D], a[D], b = 0.0f;
    for (int c = 0; c < D; ++c) { float t = 0.0f; for (int r = 0; r < D; ++r) t += dx[r] * s.Jinv[r * D + c]; vT[c] = t; }
    for (int r = 0; r < D; ++r) { float t = 0.0f; for (int c = 0; c < D; ++c) t += s.Jinv[r * D + c] * dg[c]; a[r] = dx[r] - t; }
    for (int c = 0; c < D; ++c) b += vT[c] * dg[c];
    if (b >= 0.0f) b += BROYDEN_EPS; else b -= BROYDEN_EPS;
    for (int r = 0; r < D; ++r) { const float u = a[r] / b; for (int c = 0; c < D; ++c) s.Jinv[r * D + c] += u * vT[c]; }
    for (int r = 0; r < D; ++r) { float t = 0.0f; for (int c = 0; c < D; ++c) t += s.Jinv[r * D + c] * s.gx[c]; s.upd[r] = -t; }
    return true;
}

// torch.linspace(0, 1, n)[i] in float32 (ATen evaluates symmetrically from both ends)
ARAH_HD float linspace01(int i, int n) {
    if (n == 1) return 0.0f;
    const float step = 1.0f / (float)(n - 1);
    return (i < n / 2) ? (step * (float)i) : (1.0f - step * (float)(n - 1 - i));
}

// sigma-from-SDF (VolSDF Laplace CDF), implicit_differentiable_renderer.py:366-368
ARAH_HD float laplace_density(float sdf_m, float inv_beta) {
    const float ms = -sdf_m;
    const float sg = (ms > 0.0f) ? 1.0f : ((ms < 0.0f) ? -1.0f : 0.0f);
    const float den = inv_beta * (0.5f + 0.5f * sg * (1.0f - expf(-fabsf(ms) * inv_beta)));
    return fmaxf(den, 0.0f);
}

// sin(x) to ~1-2 ulp for |x| < ~1e4 without libdevice's large-argument slow path: the 32-way unrolled epilogues would
// otherwise inline 64 copies of the Payne-Hanek fallback (local-memory tables, divergent regions) and thrash the
// instruction cache.  3-term Cody-Waite reduction by pi, odd Taylor polynomial to r^11 on [-pi/2, pi/2].
ARAH_HD float sin_cw(float x) {
    const float k = rintf(x * 0.318309886183790672f);
    float r = fmaf(-k, 3.140625f, x);
    r = fmaf(-k, 9.67502593994140625e-4f, r);
    r = fmaf(-k, 1.509957990978376e-7f, r);
    const float s = r * r;
    float p = fmaf(s, -2.5052108385441718775e-8f, 2.7557319223985890653e-6f);
    p = fmaf(s, p, -1.9841269841269841270e-4f);
    p = fmaf(s, p, 8.3333333333333333333e-3f);
    p = fmaf(s, p, -1.6666666666666666667e-1f);
    const float res = fmaf(r * s, p, r);
    return (((int)k) & 1) ? -res : res;
}

// sin and cos with one shared 3-term Cody-Waite reduction by pi (|x| < ~1e4), odd / even polynomials on [-pi/2, pi/2]
ARAH_HD void sincos_cw(float x, float& sn, float& cs) {
    const float k = rintf(x * 0.318309886183790672f);
    float r = fmaf(-k, 3.140625f, x);
    r = fmaf(-k, 9.67502593994140625e-4f, r);
    r = fmaf(-k, 1.509957990978376e-7f, r);
    const float s = r * r;
    float p = fmaf(s, -2.5052108385441718775e-8f, 2.7557319223985890653e-6f);
    p = fmaf(s, p, -1.9841269841269841270e-4f);
    p = fmaf(s, p, 8.3333333333333333333e-3f);
    p = fmaf(s, p, -1.6666666666666666667e-1f);
    const float rs = fmaf(r * s, p, r);
    float q = fmaf(s, 2.0876756987868098979e-9f, -2.7557319223985890653e-7f);
    q = fmaf(s, q, 2.4801587301587301587e-5f);
    q = fmaf(s, q, -1.3888888888888888889e-3f);
    q = fmaf(s, q, 4.1666666666666666667e-2f);
    q = fmaf(s, q, -0.5f);
    const float rc = fmaf(s, q, 1.0f);
    const bool odd = ((int)k) & 1;
    sn = odd ? -rs : rs;
    cs = odd ? -rc : rc;
}

}  // namespace arah
