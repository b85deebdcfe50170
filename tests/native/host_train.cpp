// TEST INFRASTRUCTURE — host instantiation of the training engine (arah_release_b200/csrc/arah_train.h).
//
// The product instantiates arah::train::Session with the CUDA backend only.  This file instantiates the SAME templates
// with a plain-loop host backend so that tests/test_train_host.py can check the hand-written chain rule against
// torch.autograd on the CPU (no GPU in the build container).  It is compiled by the test on demand with g++ and is never
// loaded by the product.
#include <stdlib.h>
#include <string.h>
#include "../../arah_release_b200/csrc/arah_train.h"

using namespace arah;
using namespace arah::train;

struct HostBK {
    typedef int Stream;
    static float* alloc(size_t floats) { return static_cast<float*>(calloc(floats ? floats : 1, sizeof(float))); }
    static void free(float* p) { ::free(p); }
    static void zero(float* p, size_t n, Stream) { memset(p, 0, n * sizeof(float)); }
    template <class F>
    static void for_each(size_t n, F f, Stream) {
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)n; ++i) f((size_t)i);
    }
    template <int NR, class F>
    static void col_reduce(int M, int N, F f, float* const* outs, Stream) {
#pragma omp parallel for schedule(static)
        for (int j = 0; j < N; ++j) {
            double acc[NR];
            for (int r = 0; r < NR; ++r) acc[r] = 0.0;
            for (int m = 0; m < M; ++m) {
                float red[NR];
                f(m, j, red);
                for (int r = 0; r < NR; ++r) acc[r] += red[r];
            }
            for (int r = 0; r < NR; ++r) if (outs[r]) outs[r][j] += (float)acc[r];
        }
    }
    static void gemm(int M, int N, int K, const float* A, long sa_i, long sa_k, const float* B, long sb_k, long sb_j, float* C, int ldc,
                     const float* bias, bool accumulate, Stream) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < N; ++j) {
                float acc = 0.0f;
                for (int k = 0; k < K; ++k) acc += A[i * sa_i + k * sa_k] * B[k * sb_k + j * sb_j];
                if (bias) acc += bias[j];
                float* d = C + (long)i * ldc + j;
                *d = accumulate ? (*d + acc) : acc;
            }
    }
};

struct HostTrainParams {      // mirrors the pointer part of ArahFrame / ArahTrainGrads (include/arah_b200.h)
    const float* sdf_W[7]; const float* sdf_b[7]; const float* sdf_freq; const float* sdf_phase;
    const float* skin_W[5]; const float* skin_b[5];
    const float* col_W[6]; const float* col_b[6]; const float* latent; int32_t latent_dim;
    const float* bone_T; float cmin, cmax, center[3];
};
struct HostTrainGrads {
    float* sdf_W[7]; float* sdf_b[7]; float* sdf_freq; float* sdf_phase;
    float* skin_W[5]; float* skin_b[5];
    float* col_W[6]; float* col_b[6]; float* latent; float* beta;
};

static AllParams to_params(const HostTrainParams* p) {
    AllParams A;
    for (int i = 0; i < 7; ++i) { A.sdf.W[i] = p->sdf_W[i]; A.sdf.b[i] = p->sdf_b[i]; }
    A.sdf.freq = p->sdf_freq; A.sdf.phase = p->sdf_phase;
    for (int i = 0; i < 5; ++i) { A.skin.W[i] = p->skin_W[i]; A.skin.b[i] = p->skin_b[i]; }
    for (int i = 0; i < 6; ++i) { A.col.W[i] = p->col_W[i]; A.col.b[i] = p->col_b[i]; }
    A.col.latent = p->latent; A.col.latent_dim = p->latent_dim;
    A.bone_T = p->bone_T;
    A.nm.cmin = p->cmin; A.nm.cmax = p->cmax;
    for (int k = 0; k < 3; ++k) A.nm.center[k] = p->center[k];
    return A;
}
static AllGrads to_grads(const HostTrainGrads* g) {
    AllGrads G;
    for (int i = 0; i < 7; ++i) { G.sdf.W[i] = g->sdf_W[i]; G.sdf.b[i] = g->sdf_b[i]; }
    G.sdf.freq = g->sdf_freq; G.sdf.phase = g->sdf_phase;
    for (int i = 0; i < 5; ++i) { G.skin.W[i] = g->skin_W[i]; G.skin.b[i] = g->skin_b[i]; }
    for (int i = 0; i < 6; ++i) { G.col.W[i] = g->col_W[i]; G.col.b[i] = g->col_b[i]; }
    G.col.latent = g->latent; G.beta = g->beta;
    return G;
}

static Session<HostBK> g_sess;
static int* g_list = nullptr;

extern "C" {

int host_train_shade_forward(const HostTrainParams* p, int P, int S, int cano_view_dirs, int ray_augm, int train_skinning_net, float beta,
                             const float* smp_xn, const float* smp_T12, const float* z_vals, const uint8_t* smp_conv,
                             const float* view, const float* view_orig, float* rgb, float* wsum) {
    ::free(g_list);
    g_list = static_cast<int*>(malloc(sizeof(int) * (size_t)P * S + 4));
    int M = 0;
    for (int sl = 0; sl < P * S; ++sl) if (smp_conv[sl]) g_list[M++] = sl;
    ShadeGeom g;
    g.P = P; g.S = S; g.cano_view_dirs = cano_view_dirs; g.ray_augm = ray_augm; g.list = g_list; g.smp_xn = smp_xn; g.smp_T = smp_T12;
    g.z_vals = z_vals; g.smp_conv = smp_conv; g.view = view; g.view_orig = view_orig ? view_orig : view;
    g.sdf_scale = p->cmax - p->cmin; g.beta_raw = beta;
    const int rc = g_sess.shade_forward(to_params(p), g, M, train_skinning_net != 0, rgb, wsum, 0);
    return rc == 0 ? M : rc;
}
int host_train_shade_backward(const HostTrainParams* p, const HostTrainGrads* gr, const float* g_rgb, const float* g_wsum) {
    return g_sess.shade_backward(to_params(p), to_grads(gr), g_rgb, g_wsum, 0);
}
int host_train_sdf_forward(const HostTrainParams* p, int slot, const float* pts, int n, int with_grad, float* sdf, float* grad) {
    return g_sess.sdf_forward(to_params(p), slot, pts, n, with_grad != 0, sdf, grad, 0);
}
int host_train_sdf_backward(const HostTrainParams* p, const HostTrainGrads* gr, int slot, const float* g_sdf, const float* g_grad) {
    return g_sess.sdf_backward(to_params(p), to_grads(gr), slot, g_sdf, g_grad, 0);
}
int host_train_skin_forward(const HostTrainParams* p, const float* pts_m, int n, float* w) {
    return g_sess.skin_forward(to_params(p), pts_m, n, w, 0);
}
int host_train_skin_backward(const HostTrainParams* p, const HostTrainGrads* gr, const float* g_w) {
    return g_sess.skin_backward(to_params(p), to_grads(gr), g_w, 0);
}
void host_train_hsoftmax_vjp(const float* x25, const float* gp24, float* gx25) { hierarchical_softmax_vjp(x25, gp24, gx25); }
void host_train_release(void) { g_sess.release(); ::free(g_list); g_list = nullptr; }

}  // extern "C"
