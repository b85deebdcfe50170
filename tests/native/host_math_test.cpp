// host_math_test.cpp — compiles arah_math.cuh (the per-point math the kernels run) for the HOST so that
// tests/test_host_math.py can check it on CPU against independent implementations (torch / the reference's broyden()).
#include <cstring>
#include "../../arah_release_b200/csrc/arah_math.cuh"

using namespace arah;

extern "C" {
void hm_hsoftmax(const float* logits25_times20, int n, float* out24) {
    for (int i = 0; i < n; ++i) hierarchical_softmax(logits25_times20 + 25 * i, out24 + 24 * i);
}
void hm_hsoftmax_dual(const float* x25, const float* dx25x3, float* w24, float* dw24x3) {
    Dual3 x[25], p[24];
    for (int c = 0; c < 25; ++c) { x[c].v = x25[c]; for (int k = 0; k < 3; ++k) x[c].d[k] = dx25x3[c * 3 + k]; }
    hierarchical_softmax_dual(x, p);
    for (int j = 0; j < 24; ++j) { w24[j] = p[j].v; for (int k = 0; k < 3; ++k) dw24x3[j * 3 + k] = p[j].d[k]; }
}
// Broyden on g(x) = A x + c + eps * sin(3 x) (component-wise), D = 3, same driver structure as k_corr_step
int hm_broyden3(const float* A9, const float* c3, float eps, const float* x0, const float* Jinv0, const float* Tinit12,
                float* x_out, float* T_out12, float* diff, int* valid, int max_steps) {
    auto g = [&](const float* x, float* gx, float* T12) {
        for (int r = 0; r < 3; ++r) {
            float s = c3[r];
            for (int k = 0; k < 3; ++k) s += A9[r * 3 + k] * x[k];
            gx[r] = s + eps * sinf(3.0f * x[r]);
        }
        for (int e = 0; e < 12; ++e) T12[e] = x[e % 3] * (float)(e + 1);      // a transform that depends on x (bookkeeping check)
    };
    BroydenState<3> st;
    float g0[3], T[12];
    g(x0, g0, T);
    broyden_begin<3>(st, x0, g0, Jinv0, Tinit12);
    int evals = 1;
    for (int it = 0; it < max_steps; ++it) {
        float dx[3], gn[3];
        broyden_advance<3>(st, dx);
        g(st.x, gn, T);
        ++evals;
        if (!broyden_update<3>(st, dx, gn, T)) break;
    }
    memcpy(x_out, st.best_x, sizeof(float) * 3);
    memcpy(T_out12, st.best_T, sizeof(float) * 12);
    *diff = st.best_n;
    *valid = st.best_n < CVG_THRESH;
    return evals;
}
// the Cody-Waite sine / sine+cosine of the tensor-core epilogues (FiLM-SIREN arguments reach |x| ~ 100)
void hm_sincos(const float* x, int n, float* s_only, float* s, float* c) {
    for (int i = 0; i < n; ++i) { s_only[i] = sin_cw(x[i]); sincos_cw(x[i], s[i], c[i]); }
}
void hm_misc(float* out) {
    for (int i = 0; i < 17; ++i) out[i] = linspace01(i, 17);
    for (int i = 0; i < 16; ++i) out[17 + i] = linspace01(i, 16);
    out[33] = laplace_density(0.0f, 200.0f);
    out[34] = laplace_density(0.01f, 200.0f);
    out[35] = laplace_density(-0.01f, 200.0f);
    out[36] = softplus100(0.003f);
    out[37] = softplus100(0.5f);
    float A[9] = {2, 0.1f, 0, 0.3f, 1.5f, 0.2f, 0, 0.4f, 1.1f}, Ai[9];
    invert3(A, Ai);
    for (int i = 0; i < 9; ++i) out[38 + i] = Ai[i];
    float B[16] = {1.2f, 0.1f, 0, 0.3f, 0.2f, 0.9f, 0.1f, 0, 0, 0.3f, 1.1f, 0.5f, 0.1f, 0, 0.2f, 1.0f}, Bi[16];
    invert_gj<4>(B, Bi);
    for (int i = 0; i < 16; ++i) out[47 + i] = Bi[i];
}
}
