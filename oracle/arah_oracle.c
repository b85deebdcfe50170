/*
 * arah_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, fp32, one ray / one point at a time) of the reference's hot path
 *   /root/reference/im2mesh/metaavatar_render/renderer/{ray_tracing,implicit_differentiable_renderer}.py
 *   /root/reference/im2mesh/utils/{root_finding_utils,broyden}.py
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker / the reported CPU baseline.  The product path (arah_release_b200) never
 * links, imports or executes it.
 *
 * Parity status: PINNED against outputs of the reference itself run in the build container
 * (oracle/gen_golden.py -> tests/golden/ *.npz; the reference has no tests or golden vectors of its own for
 * this path).  Floating point: the reference is batched torch (MKL sgemm); this file is
 * sequential-k fp32 FMA, so results agree to rounding, not bitwise; tolerances live in the tests.
 *
 * The batched-mask control flow of the reference is restated per ray / per point; equivalences relied upon:
 *   - broyden(): global early exit has no per-point effect, every point takes >= 1 step, best-iterate
 *     bookkeeping starts from T_init (utils/broyden.py:38-40,45,57-65).
 *   - sphere_tracing(): a ray leaves the active set for good once converged/diverged (ray_tracing.py:238-241).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NJ 24
#define SDF_H 256
#define SKIN_H 128
#define COL_H 256
#define MAX_STEPS 256
#define BROYDEN_MAX 50

typedef struct {
    /* SDF FiLM-SIREN, reference layout [out][in] (hyperlayers.py:391-415) */
    const float *sdf_W[7], *sdf_b[7];
    const float *sdf_freq, *sdf_phase;            /* [6][256] */
    /* skinning MLP, weight-norm already folded, [out][in] (metaavatar/models/decoder.py:201-233) */
    const float *skin_W[5], *skin_b[5];
    /* colour MLP, weight-norm folded, latent NOT folded: lin0 [256][d_in], lin3 [256][d_in+128] */
    const float *col_W[6], *col_b[6];
    const float *latent;                           /* [latent_dim] */
    int32_t latent_dim;                            /* 128 */
    float beta;                                    /* SingleVarianceNetwork.variance */
    const float *bone_T;                           /* [24][16] */
    const float *smpl_verts;                       /* [n_verts][3] posed + trans */
    const float *smpl_w;                           /* [n_verts][24] */
    int32_t n_verts;
    float trans[3], cmin, cmax, center[3], cam_loc[3], pose[16];
    int32_t n_steps, near_samples, far_samples, cano_view_dirs;
    int32_t render_last_pt;                        /* implicit_differentiable_renderer.py:380-381: last interval 1e10 instead of 1 / n_steps */
} OracleFrame;

typedef struct {
    /* all optional (NULL to skip) */
    float *points_hat_norm;      /* [P][3]   BodyRayTracing.forward tuple[0] */
    uint8_t *trace_mask;         /* [P]      tuple[1] network_body_mask (ray tracer) */
    float *dists;                /* [P]      tuple[2] */
    float *sampled_pts;          /* [P][S][3] */
    float *sampled_dists;        /* [P][S] */
    float *sampled_T;            /* [P][S][16] */
    uint8_t *sampled_conv;       /* [P][S] */
    float *rgb;                  /* [P][3] */
    uint8_t *vol_mask;           /* [P] */
    float *points_cam;           /* [P][3] */
    float *weights_sum;          /* [P] */
    /* counters: per ray */
    int32_t *n_trace_evals;      /* sphere tracing SDF evals */
    int32_t *n_iso_evals;        /* iso-search g evals (0 if not searched) */
    int32_t *n_corr_evals;       /* sum over samples of correspondence g evals */
    int32_t *n_shaded;           /* shaded samples */
} OracleOut;

/* ---------------------------------------------------------------- packed weights */
typedef struct {
    float *sdf_Wt[7];     /* [in][out] */
    float *skin_Wt[5];
    float *col_Wt[6];
    int col_in;           /* 3 + 27 + 3 + 256 + latent */
} Packed;

static float *transpose(const float *W, int out, int in) {
    float *t = (float *)malloc(sizeof(float) * (size_t)out * in);
    for (int o = 0; o < out; ++o)
        for (int i = 0; i < in; ++i) t[(size_t)i * out + o] = W[(size_t)o * in + i];
    return t;
}

static void pack(const OracleFrame *f, Packed *p) {
    const int sd_in[7] = {3, 256, 256, 256, 256, 256, 256}, sd_out[7] = {256, 256, 256, 256, 256, 256, 1};
    for (int l = 0; l < 7; ++l) p->sdf_Wt[l] = transpose(f->sdf_W[l], sd_out[l], sd_in[l]);
    const int sk_in[5] = {3, 128, 128, 128, 128}, sk_out[5] = {128, 128, 128, 128, 25};
    for (int l = 0; l < 5; ++l) p->skin_Wt[l] = transpose(f->skin_W[l], sk_out[l], sk_in[l]);
    p->col_in = 3 + 27 + 3 + 256 + f->latent_dim;
    const int c_in[6] = {p->col_in, 256, 256, p->col_in + 128, 256, 256}, c_out[6] = {256, 256, 128, 256, 256, 3};
    for (int l = 0; l < 6; ++l) p->col_Wt[l] = transpose(f->col_W[l], c_out[l], c_in[l]);
}
static void unpack(Packed *p) {
    for (int l = 0; l < 7; ++l) free(p->sdf_Wt[l]);
    for (int l = 0; l < 5; ++l) free(p->skin_Wt[l]);
    for (int l = 0; l < 6; ++l) free(p->col_Wt[l]);
}

/* y[o] = b[o] + sum_i Wt[i][o] x[i]   (k-sequential per output, vectorises over o) */
__attribute__((target_clones("avx512f","default")))
static void matvec_t(const float *restrict Wt, const float *restrict b, const float *restrict x,
                            float *restrict y, int in, int out) {
    for (int o = 0; o < out; ++o) y[o] = 0.0f;
    for (int i = 0; i < in; ++i) {
        const float xi = x[i];
        const float *restrict w = Wt + (size_t)i * out;
        for (int o = 0; o < out; ++o) y[o] += w[o] * xi;
    }
    if (b) for (int o = 0; o < out; ++o) y[o] += b[o];     /* reference: matmul then += bias */
}
/* gx[i] = sum_o W[o][i] g[o]  (reverse-mode through a linear layer, W in [out][in]) */
__attribute__((target_clones("avx512f","default")))
static void matvec_bwd(const float *restrict W, const float *restrict g, float *restrict gx, int in, int out) {
    for (int i = 0; i < in; ++i) gx[i] = 0.0f;
    for (int o = 0; o < out; ++o) {
        const float go = g[o];
        const float *restrict w = W + (size_t)o * in;
        for (int i = 0; i < in; ++i) gx[i] += w[i] * go;
    }
}

/* ---------------------------------------------------------------- normalisation (root_finding_utils.py:37-51) */
static inline void normalize_pts(const OracleFrame *f, const float *p, float *q) {
    const float d = f->cmax - f->cmin, padding = d * 0.05f;
    for (int k = 0; k < 3; ++k) {
        float v = p[k] - f->center[k];
        v = (v - f->cmin + padding) / d / 1.1f;
        v = v - 0.5f;
        q[k] = v * 2.0f;
    }
}
static inline void unnormalize_pts(const OracleFrame *f, const float *p, float *q) {
    const float d = f->cmax - f->cmin, padding = d * 0.05f;
    for (int k = 0; k < 3; ++k) q[k] = (p[k] / 2.0f + 0.5f) * 1.1f * d + f->cmin - padding + f->center[k];
}
static inline float sdf_to_metres(const OracleFrame *f, float s) { return s / 2.0f * 1.1f * (f->cmax - f->cmin); }

/* ---------------------------------------------------------------- SDF network (hyperlayers.py:412-415, siren_modules.py:35-37) */
/* forward; optionally keeps the sine arguments (for the gradient) and returns the last hidden feature */
static float sdf_forward(const OracleFrame *f, const Packed *p, const float *xn, float *feat /*256 or NULL*/,
                         float *args /*6*256 or NULL*/) {
    float h[SDF_H], a[SDF_H];
    const float *in = xn;
    int nin = 3;
    for (int l = 0; l < 6; ++l) {
        matvec_t(p->sdf_Wt[l], f->sdf_b[l], in, a, nin, SDF_H);
        const float *fr = f->sdf_freq + l * SDF_H, *ph = f->sdf_phase + l * SDF_H;
        for (int o = 0; o < SDF_H; ++o) {
            const float t = 30.0f * (fr[o] * a[o] + ph[o]);
            if (args) args[l * SDF_H + o] = t;
            h[o] = sinf(t);
        }
        in = h;
        nin = SDF_H;
    }
    if (feat) memcpy(feat, h, sizeof(float) * SDF_H);
    float y;
    matvec_t(p->sdf_Wt[6], f->sdf_b[6], h, &y, SDF_H, 1);
    return y;
}
/* d sdf / d xn (reverse mode == what autograd does, diff_operators.py:39-50) */
static void sdf_gradient(const OracleFrame *f, const float *args, float *grad3) {
    float g[SDF_H], gh[SDF_H];
    for (int o = 0; o < SDF_H; ++o) gh[o] = f->sdf_W[6][o];           /* d y / d h5 */
    for (int l = 5; l >= 0; --l) {
        const float *fr = f->sdf_freq + l * SDF_H;
        for (int o = 0; o < SDF_H; ++o) g[o] = gh[o] * cosf(args[l * SDF_H + o]) * 30.0f * fr[o];
        if (l > 0) matvec_bwd(f->sdf_W[l], g, gh, SDF_H, SDF_H);
        else matvec_bwd(f->sdf_W[0], g, grad3, 3, SDF_H);
    }
}

/* ---------------------------------------------------------------- hierarchical softmax (utils/utils.py:138-181) */
static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
static inline void softmax3(const float *x, float *y) {
    float m = fmaxf(x[0], fmaxf(x[1], x[2]));
    float e0 = expf(x[0] - m), e1 = expf(x[1] - m), e2 = expf(x[2] - m), s = e0 + e1 + e2;
    y[0] = e0 / s; y[1] = e1 / s; y[2] = e2 / s;
}
static void hierarchical_softmax(const float *x /*25*/, float *p /*24*/) {
    float sm[3];
    for (int j = 0; j < NJ; ++j) p[j] = 1.0f;
    softmax3(x + 1, sm);
    float s0 = sigmoidf_(x[0]);
    for (int k = 0; k < 3; ++k) p[1 + k] = p[0] * s0 * sm[k];
    p[0] = p[0] * (1.0f - s0);
#define SPLIT(c, q, g) do { float s_ = sigmoidf_(x[g]); p[c] = p[q] * s_; p[q] = p[q] * (1.0f - s_); } while (0)
    SPLIT(4, 1, 4); SPLIT(5, 2, 5); SPLIT(6, 3, 6);
    SPLIT(7, 4, 7); SPLIT(8, 5, 8); SPLIT(9, 6, 9);
    SPLIT(10, 7, 10); SPLIT(11, 8, 11);
    softmax3(x + 12, sm);
    { float s24 = sigmoidf_(x[24]), p9 = p[9];
      for (int k = 0; k < 3; ++k) p[12 + k] = p9 * s24 * sm[k];
      p[9] = p9 * (1.0f - s24); }
    SPLIT(15, 12, 15);
    SPLIT(16, 13, 16); SPLIT(17, 14, 17);
    SPLIT(18, 16, 18); SPLIT(19, 17, 19);
    SPLIT(20, 18, 20); SPLIT(21, 19, 21);
    SPLIT(22, 20, 22); SPLIT(23, 21, 23);
#undef SPLIT
}

/* dual numbers with 3 tangents for the full LBS Jacobian (forward_skinning_jac, root_finding_utils.py:170-230) */
typedef struct { float v, d[3]; } Dual;
static inline Dual d_mul(Dual a, Dual b) { Dual r; r.v = a.v * b.v; for (int k = 0; k < 3; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k]; return r; }
static inline Dual d_one_minus(Dual a) { Dual r; r.v = 1.0f - a.v; for (int k = 0; k < 3; ++k) r.d[k] = -a.d[k]; return r; }
static inline Dual d_sigmoid(Dual a) { Dual r; r.v = sigmoidf_(a.v); float s = r.v * (1.0f - r.v); for (int k = 0; k < 3; ++k) r.d[k] = s * a.d[k]; return r; }
static void d_softmax3(const Dual *x, Dual *y) {
    float v[3] = {x[0].v, x[1].v, x[2].v}, s[3];
    softmax3(v, s);
    for (int i = 0; i < 3; ++i) {
        y[i].v = s[i];
        for (int k = 0; k < 3; ++k) {
            float dot = s[0] * x[0].d[k] + s[1] * x[1].d[k] + s[2] * x[2].d[k];
            y[i].d[k] = s[i] * (x[i].d[k] - dot);
        }
    }
}
static void hierarchical_softmax_dual(const Dual *x, Dual *p) {
    Dual sm[3], one = {1.0f, {0, 0, 0}};
    for (int j = 0; j < NJ; ++j) p[j] = one;
    d_softmax3(x + 1, sm);
    Dual s0 = d_sigmoid(x[0]);
    for (int k = 0; k < 3; ++k) p[1 + k] = d_mul(d_mul(p[0], s0), sm[k]);
    p[0] = d_mul(p[0], d_one_minus(s0));
#define SPLIT(c, q, g) do { Dual s_ = d_sigmoid(x[g]); p[c] = d_mul(p[q], s_); p[q] = d_mul(p[q], d_one_minus(s_)); } while (0)
    SPLIT(4, 1, 4); SPLIT(5, 2, 5); SPLIT(6, 3, 6);
    SPLIT(7, 4, 7); SPLIT(8, 5, 8); SPLIT(9, 6, 9);
    SPLIT(10, 7, 10); SPLIT(11, 8, 11);
    d_softmax3(x + 12, sm);
    { Dual s24 = d_sigmoid(x[24]), p9 = p[9];
      for (int k = 0; k < 3; ++k) p[12 + k] = d_mul(d_mul(p9, s24), sm[k]);
      p[9] = d_mul(p9, d_one_minus(s24)); }
    SPLIT(15, 12, 15);
    SPLIT(16, 13, 16); SPLIT(17, 14, 17);
    SPLIT(18, 16, 18); SPLIT(19, 17, 19);
    SPLIT(20, 18, 20); SPLIT(21, 19, 21);
    SPLIT(22, 20, 22); SPLIT(23, 21, 23);
#undef SPLIT
}

/* ---------------------------------------------------------------- skinning net (query_weights, root_finding_utils.py:54-113) */
static inline float softplus100(float x) {       /* torch.nn.Softplus(beta=100, threshold=20) */
    const float bx = x * 100.0f;
    return bx > 20.0f ? x : log1pf(expf(bx)) / 100.0f;
}
static void skin_logits(const OracleFrame *f, const Packed *p, const float *xn, float *logits /*25*/,
                        float *pre /*4*128 or NULL: pre-activations for the tangent pass*/) {
    float h[SKIN_H], a[SKIN_H];
    const float *in = xn;
    int nin = 3;
    for (int l = 0; l < 4; ++l) {
        matvec_t(p->skin_Wt[l], f->skin_b[l], in, a, nin, SKIN_H);
        if (pre) memcpy(pre + l * SKIN_H, a, sizeof(a));
        for (int o = 0; o < SKIN_H; ++o) h[o] = softplus100(a[o]);
        in = h;
        nin = SKIN_H;
    }
    matvec_t(p->skin_Wt[4], f->skin_b[4], h, logits, SKIN_H, 25);
}
static void query_weights(const OracleFrame *f, const Packed *p, const float *x_hat /*metres*/, float *w /*24*/) {
    float xn[3], lg[25];
    normalize_pts(f, x_hat, xn);
    skin_logits(f, p, xn, lg, NULL);
    for (int k = 0; k < 25; ++k) lg[k] *= 20.0f;
    hierarchical_softmax(lg, w);
}
/* T = sum_j w_j B_j  (skinning(), root_finding_utils.py:13-33) ; x_bar = (T [x;1])[:3] */
static inline void blend(const float *w, const float *B, float *T) {
    for (int e = 0; e < 16; ++e) T[e] = 0.0f;
    for (int j = 0; j < NJ; ++j) {
        const float wj = w[j];
        for (int e = 0; e < 16; ++e) T[e] += wj * B[j * 16 + e];
    }
}
static inline void apply_T(const float *T, const float *x, float *y) {
    for (int r = 0; r < 3; ++r) y[r] = T[r * 4 + 0] * x[0] + T[r * 4 + 1] * x[1] + T[r * 4 + 2] * x[2] + T[r * 4 + 3];
}
static void forward_skinning(const OracleFrame *f, const Packed *p, const float *x_hat, float *x_bar, float *T) {
    float w[NJ];
    query_weights(f, p, x_hat, w);
    blend(w, f->bone_T, T);
    apply_T(T, x_hat, x_bar);
}
/* full Jacobian d LBS(x_hat) / d x_hat, including d w / d x_hat (forward-mode; autograd in the reference) */
static void forward_skinning_jac(const OracleFrame *f, const Packed *p, const float *x_hat, float *J /*3x3*/) {
    float xn[3], lg[25], pre[4 * SKIN_H];
    normalize_pts(f, x_hat, xn);
    skin_logits(f, p, xn, lg, pre);
    const float dn = 2.0f / (f->cmax - f->cmin) / 1.1f;          /* d xn / d x_hat (diagonal) */
    float t[3][SKIN_H], tn[3][SKIN_H], tl[3][25];
    for (int k = 0; k < 3; ++k) {                                   /* layer 0: tangent of W0 xn is column k * dn */
        for (int o = 0; o < SKIN_H; ++o) {
            const float a = pre[o], s = (a * 100.0f > 20.0f) ? 1.0f : sigmoidf_(a * 100.0f);
            t[k][o] = s * (f->skin_W[0][o * 3 + k] * dn);
        }
    }
    for (int l = 1; l < 4; ++l) {
        for (int k = 0; k < 3; ++k) {
            matvec_t(p->skin_Wt[l], NULL, t[k], tn[k], SKIN_H, SKIN_H);
            for (int o = 0; o < SKIN_H; ++o) {
                const float a = pre[l * SKIN_H + o], s = (a * 100.0f > 20.0f) ? 1.0f : sigmoidf_(a * 100.0f);
                tn[k][o] *= s;
            }
        }
        memcpy(t, tn, sizeof(t));
    }
    for (int k = 0; k < 3; ++k) matvec_t(p->skin_Wt[4], NULL, t[k], tl[k], SKIN_H, 25);
    Dual x[25], w[NJ];
    for (int c = 0; c < 25; ++c) {
        x[c].v = lg[c] * 20.0f;
        for (int k = 0; k < 3; ++k) x[c].d[k] = tl[k][c] * 20.0f;
    }
    hierarchical_softmax_dual(x, w);
    for (int e = 0; e < 9; ++e) J[e] = 0.0f;
    for (int j = 0; j < NJ; ++j) {
        const float *B = f->bone_T + j * 16;
        float bx[3];
        apply_T(B, x_hat, bx);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) J[r * 3 + c] += w[j].v * B[r * 4 + c] + bx[r] * w[j].d[c];
    }
}

/* ---------------------------------------------------------------- small dense algebra */
static int invert_n(const float *A, float *Ainv, int n) {   /* Gauss-Jordan with partial pivoting, n <= 4 */
    float M[4][8];
    for (int r = 0; r < n; ++r) {
        for (int c = 0; c < n; ++c) { M[r][c] = A[r * n + c]; M[r][n + c] = (r == c) ? 1.0f : 0.0f; }
    }
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r) if (fabsf(M[r][c]) > fabsf(M[piv][c])) piv = r;
        if (M[piv][c] == 0.0f) return -1;
        if (piv != c) for (int k = 0; k < 2 * n; ++k) { float t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
        const float inv = 1.0f / M[c][c];
        for (int k = 0; k < 2 * n; ++k) M[c][k] *= inv;
        for (int r = 0; r < n; ++r) if (r != c) {
            const float fct = M[r][c];
            if (fct != 0.0f) for (int k = 0; k < 2 * n; ++k) M[r][k] -= fct * M[c][k];
        }
    }
    for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) Ainv[r * n + c] = M[r][n + c];
    return 0;
}

/* ---------------------------------------------------------------- kNN inverse skinning (ray_tracing.py:382-400, 403-421) */
static int knn1(const OracleFrame *f, const float *x) {
    int best = 0;
    float bd = INFINITY;
    for (int v = 0; v < f->n_verts; ++v) {
        const float dx = x[0] - f->smpl_verts[v * 3], dy = x[1] - f->smpl_verts[v * 3 + 1], dz = x[2] - f->smpl_verts[v * 3 + 2];
        const float d = dx * dx + dy * dy + dz * dz;
        if (d < bd) { bd = d; best = v; }
    }
    return best;
}
/* x: posed point incl. trans.  Returns x_hat in metres (T^-1 (x - trans)) and the kNN-blend transform T. */
static void knn_inverse_skinning(const OracleFrame *f, const float *x, float *x_hat, float *T) {
    const int v = knn1(f, x);
    blend(f->smpl_w + (size_t)v * NJ, f->bone_T, T);
    float Ti[16];
    if (invert_n(T, Ti, 4) != 0) memset(Ti, 0, sizeof(Ti));
    const float xl[3] = {x[0] - f->trans[0], x[1] - f->trans[1], x[2] - f->trans[2]};
    for (int r = 0; r < 3; ++r) x_hat[r] = Ti[r * 4] * xl[0] + Ti[r * 4 + 1] * xl[1] + Ti[r * 4 + 2] * xl[2] + Ti[r * 4 + 3] * 1.0f;
}

/* ---------------------------------------------------------------- Broyden (utils/broyden.py:4-78), per point */
typedef void (*gfun)(void *ctx, const float *x, float *gx, float *T);
typedef struct { float x[4], T[16], diff; int valid, g_evals; } BroydenResult;

static void broyden(gfun g, void *ctx, int D, const float *x0, const float *T0, const float *Jinv0, BroydenResult *res) {
    const float cvg = 1e-5f, dvg = 1.0f, eps = 1e-6f;
    float x[4], T[16], Ji[16], gx[4], upd[4], dx[4] = {0, 0, 0, 0}, dg[4] = {0, 0, 0, 0}, Tdummy[16];
    memcpy(x, x0, sizeof(float) * D);
    memcpy(T, T0, sizeof(T));
    memcpy(Ji, Jinv0, sizeof(float) * D * D);
    g(ctx, x, gx, Tdummy);                                   /* :36  (its T is discarded) */
    int evals = 1;
    for (int r = 0; r < D; ++r) { float s = 0; for (int c = 0; c < D; ++c) s += Ji[r * D + c] * gx[c]; upd[r] = -s; }
    memcpy(res->x, x, sizeof(float) * D);
    memcpy(res->T, T, sizeof(T));                            /* :39-40 best-so-far starts at T_init */
    float nrm = 0; for (int c = 0; c < D; ++c) nrm += gx[c] * gx[c];
    float best = sqrtf(nrm);
    for (int it = 0; it < BROYDEN_MAX; ++it) {
        for (int c = 0; c < D; ++c) { dx[c] = upd[c]; x[c] += dx[c]; }
        float gn[4];
        g(ctx, x, gn, T);
        ++evals;
        for (int c = 0; c < D; ++c) { dg[c] = gn[c] - gx[c]; gx[c] += dg[c]; }   /* :53-54 */
        nrm = 0; for (int c = 0; c < D; ++c) nrm += gx[c] * gx[c];
        const float cur = sqrtf(nrm);
        if (cur < best) { best = cur; memcpy(res->x, x, sizeof(float) * D); memcpy(res->T, T, sizeof(T)); }
        if (!(best > cvg && cur < dvg)) break;               /* :64 */
        float vT[4], a[4], b = 0;
        for (int c = 0; c < D; ++c) { float s = 0; for (int r = 0; r < D; ++r) s += dx[r] * Ji[r * D + c]; vT[c] = s; }
        for (int r = 0; r < D; ++r) { float s = 0; for (int c = 0; c < D; ++c) s += Ji[r * D + c] * dg[c]; a[r] = dx[r] - s; }
        for (int c = 0; c < D; ++c) b += vT[c] * dg[c];
        if (b >= 0) b += eps; else b -= eps;
        for (int r = 0; r < D; ++r) { const float u = a[r] / b; for (int c = 0; c < D; ++c) Ji[r * D + c] += u * vT[c]; }
        for (int r = 0; r < D; ++r) { float s = 0; for (int c = 0; c < D; ++c) s += Ji[r * D + c] * gx[c]; upd[r] = -s; }
    }
    res->diff = best;
    res->valid = best < cvg;
    res->g_evals = evals;
}

typedef struct { const OracleFrame *f; const Packed *p; float tgt[3]; } CorrCtx;
static void g_corr(void *c_, const float *x, float *gx, float *T) {        /* root_finding_utils.py:337-345 */
    CorrCtx *c = (CorrCtx *)c_;
    float xb[3];
    forward_skinning(c->f, c->p, x, xb, T);
    for (int k = 0; k < 3; ++k) gx[k] = xb[k] - c->tgt[k];
}
typedef struct { const OracleFrame *f; const Packed *p; float o[3], d[3]; } IsoCtx;
static void g_iso(void *c_, const float *u, float *gx, float *T) {         /* root_finding_utils.py:426-457 */
    IsoCtx *c = (IsoCtx *)c_;
    float xb[3], xn[3];
    forward_skinning(c->f, c->p, u, xb, T);
    for (int k = 0; k < 3; ++k) {
        const float xbar = c->d[k] * u[3] + c->o[k];
        gx[1 + k] = xb[k] - (xbar - c->f->trans[k]);
    }
    normalize_pts(c->f, u, xn);
    gx[0] = sdf_to_metres(c->f, sdf_forward(c->f, c->p, xn, NULL, NULL));
}

/* torch.linspace(0,1,n)[i] in float32 (symmetric evaluation, ATen RangeFactories) */
static inline float linspace01(int i, int n) {
    if (n == 1) return 0.0f;
    const float step = (1.0f - 0.0f) / (float)(n - 1);
    return (i < n / 2) ? (0.0f + step * (float)i) : (1.0f - step * (float)(n - 1 - i));
}
static int cmp_float(const void *a, const void *b) { float x = *(const float *)a, y = *(const float *)b; return (x > y) - (x < y); }

/* ---------------------------------------------------------------- colour network (metaavatar_render/models/decoder.py:69-124) */
static void color_forward(const OracleFrame *f, const Packed *p, const float *xn, const float *normal, const float *view,
                          const float *feat, float *rgb) {
    const int din = p->col_in;
    float in[3 + 27 + 3 + 256 + 512], h0[COL_H], h1[COL_H], cat[3 + 27 + 3 + 256 + 512 + 128];
    int q = 0;
    for (int k = 0; k < 3; ++k) in[q++] = xn[k];
    for (int k = 0; k < 3; ++k) in[q++] = view[k];                       /* embedder.py:11-36: x, sin(2^l x), cos(2^l x) */
    for (int l = 0; l < 4; ++l) {
        const float fr = (float)(1 << l);
        for (int k = 0; k < 3; ++k) in[q++] = sinf(view[k] * fr);
        for (int k = 0; k < 3; ++k) in[q++] = cosf(view[k] * fr);
    }
    for (int k = 0; k < 3; ++k) in[q++] = normal[k];
    for (int k = 0; k < 256; ++k) in[q++] = feat[k];
    for (int k = 0; k < f->latent_dim; ++k) in[q++] = f->latent[k];
    matvec_t(p->col_Wt[0], f->col_b[0], in, h0, din, 256);
    for (int o = 0; o < 256; ++o) h0[o] = fmaxf(h0[o], 0.0f);
    matvec_t(p->col_Wt[1], f->col_b[1], h0, h1, 256, 256);
    for (int o = 0; o < 256; ++o) h1[o] = fmaxf(h1[o], 0.0f);
    matvec_t(p->col_Wt[2], f->col_b[2], h1, h0, 256, 128);
    for (int o = 0; o < 128; ++o) h0[o] = fmaxf(h0[o], 0.0f);
    memcpy(cat, in, sizeof(float) * din);                                 /* skip: cat([rendering_input, x]) :113-115 */
    memcpy(cat + din, h0, sizeof(float) * 128);
    matvec_t(p->col_Wt[3], f->col_b[3], cat, h1, din + 128, 256);
    for (int o = 0; o < 256; ++o) h1[o] = fmaxf(h1[o], 0.0f);
    matvec_t(p->col_Wt[4], f->col_b[4], h1, h0, 256, 256);
    for (int o = 0; o < 256; ++o) h0[o] = fmaxf(h0[o], 0.0f);
    float out[3];
    matvec_t(p->col_Wt[5], f->col_b[5], h0, out, 256, 3);
    for (int k = 0; k < 3; ++k) rgb[k] = sigmoidf_(out[k]);
}

/* ---------------------------------------------------------------- one ray, end to end */
/* training-mode stochastic inputs (ray_tracing.py:298-311): the three torch.rand draws of ray_sampler, per ray */
typedef struct { const float *u_all /*[P][S]*/, *u_near /*[P][near+1]*/, *u_far /*[P][far]*/; } TrainNoise;

/* perturb_z_vals (ray_tracing.py:298-311): stratified jitter inside the mid-point intervals; slot `fix` keeps t = 0.5 */
static void perturb_z(float *z, int n, const float *u, int fix) {
    float lo[MAX_STEPS], up[MAX_STEPS];
    for (int i = 0; i < n; ++i) {
        lo[i] = (i == 0) ? z[0] : 0.5f * (z[i] + z[i - 1]);
        up[i] = (i == n - 1) ? z[n - 1] : 0.5f * (z[i + 1] + z[i]);
    }
    for (int i = 0; i < n; ++i) { const float t = (i == fix) ? 0.5f : u[i]; z[i] = lo[i] + (up[i] - lo[i]) * t; }
}

static void render_ray(const OracleFrame *f, const Packed *p, const float *d, float near, float far, int ray, const OracleOut *out,
                       const TrainNoise *tn /* NULL = eval mode */) {
    const int S = f->n_steps;
    const float thr = 1e-5f;
    /* ---- sphere tracing (ray_tracing.py:174-241) */
    float t = near;
    int unfinished = near < far, diverge = near >= far;
    float cur_xn[3] = {0, 0, 0}, cur_T[16];
    memset(cur_T, 0, sizeof(cur_T));
    int n_trace = 0;
    for (int it = 0; it < 50 && unfinished; ++it) {
        float x[3], xh[3];
        for (int k = 0; k < 3; ++k) x[k] = d[k] * t + f->cam_loc[k];
        knn_inverse_skinning(f, x, xh, cur_T);
        normalize_pts(f, xh, cur_xn);
        const float sdf = sdf_to_metres(f, sdf_forward(f, p, cur_xn, NULL, NULL));
        ++n_trace;
        const float sm = fminf(fmaxf(sdf, -0.1f), 0.1f);
        if (fabsf(sm) > thr && fabsf(sdf) < 1e6f) { t = t + sm; diverge = t >= far; }
        if (fabsf(sdf) <= thr || diverge) unfinished = 0;
    }
    /* ---- joint iso-surface / correspondence search (root_finding_utils.py:365-484), eval: non-diverged rays */
    float x0[3], xopt[3], zopt = t, Topt[16];
    unnormalize_pts(f, cur_xn, x0);
    memcpy(xopt, x0, sizeof(x0));
    memcpy(Topt, cur_T, sizeof(Topt));
    int conv = 0, n_iso = 0;
    if (!diverge || tn) {                                      /* training: all rays (ray_tracing.py:249) */
        float J[16], Jl[9], xn[3], args[6 * SDF_H], gs[3], Ji[16];
        forward_skinning_jac(f, p, x0, Jl);
        normalize_pts(f, x0, xn);
        sdf_forward(f, p, xn, NULL, args);
        sdf_gradient(f, args, gs);
        /* chain rule exactly as autograd composes it: d(sdf/2*1.1*D)/d sdf, d xn / d x */
        const float D = f->cmax - f->cmin, so = 1.0f / 2.0f * 1.1f * D, si = 2.0f / D / 1.1f;
        for (int c = 0; c < 3; ++c) J[c] = gs[c] * so * si;
        J[3] = 0.0f;
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) J[(r + 1) * 4 + c] = Jl[r * 3 + c]; J[(r + 1) * 4 + 3] = -d[r]; }
        if (invert_n(J, Ji, 4) != 0) memset(Ji, 0, sizeof(Ji));
        IsoCtx c = {f, p, {f->cam_loc[0], f->cam_loc[1], f->cam_loc[2]}, {d[0], d[1], d[2]}};
        const float u0[4] = {x0[0], x0[1], x0[2], t};
        BroydenResult r;
        broyden(g_iso, &c, 4, u0, cur_T, Ji, &r);
        n_iso = r.g_evals;
        memcpy(xopt, r.x, sizeof(xopt));
        zopt = r.x[3];
        memcpy(Topt, r.T, sizeof(Topt));
        conv = r.valid;
    }
    conv = conv && (zopt >= near) && (zopt <= far);            /* ray_tracing.py:266 */
    float pnorm[3];
    normalize_pts(f, xopt, pnorm);
    const float dist = conv ? zopt : near;                     /* :274-278 */
    if (out->points_hat_norm) memcpy(out->points_hat_norm + (size_t)ray * 3, pnorm, sizeof(pnorm));
    if (out->trace_mask) out->trace_mask[ray] = (uint8_t)conv;
    if (out->dists) out->dists[ray] = dist;
    if (out->n_trace_evals) out->n_trace_evals[ray] = n_trace;
    if (out->n_iso_evals) out->n_iso_evals[ray] = n_iso;

    /* ---- sample placement (ray_sampler, ray_tracing.py:313-350) */
    float z[MAX_STEPS];
    uint8_t on[MAX_STEPS];
    for (int i = 0; i < S; ++i) { z[i] = dist + (far - dist) * linspace01(i, S); on[i] = 1; }
    const int nn = f->near_samples + 1, nf = f->far_samples;
    if (tn) perturb_z(z, S, tn->u_all + (size_t)ray * S, -1);
    if (conv) {
        for (int i = nn; i < S; ++i) on[i] = 0;
        for (int i = 0; i < nn; ++i) z[i] = dist - 0.05f + (0.05f * 2) * linspace01(i, nn);
        if (tn) perturb_z(z, nn, tn->u_near + (size_t)ray * nn, f->near_samples / 2);
        if (nf > 0) {
            for (int i = 0; i < nf; ++i) { on[nn + i] = 1; z[nn + i] = near + fmaxf(dist - 0.05f - near, 1e-5f) * linspace01(i, nf); }
            if (tn) perturb_z(z + nn, nf, tn->u_far + (size_t)ray * nf, -1);
            qsort(z, nn + nf, sizeof(float), cmp_float);
        }
    }
    /* ---- canonical correspondences for every sample (inv_transform_points_opt + search_canonical_corr) */
    float spts[MAX_STEPS][3], sT[MAX_STEPS][16];
    uint8_t sconv[MAX_STEPS];
    int n_corr = 0;
    for (int i = 0; i < S; ++i) {
        memset(spts[i], 0, sizeof(spts[i]));
        memset(sT[i], 0, sizeof(sT[i]));
        sconv[i] = 0;
        if (!on[i]) continue;
        float x[3], xh0[3], T0[16], w[NJ], Tn[16], A[9], Ai[9];
        for (int k = 0; k < 3; ++k) x[k] = d[k] * z[i] + f->cam_loc[k];
        knn_inverse_skinning(f, x, xh0, T0);
        query_weights(f, p, xh0, w);                                          /* root_finding_utils.py:327-328 */
        blend(w, f->bone_T, Tn);
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A[r * 3 + c] = Tn[r * 4 + c];
        if (invert_n(A, Ai, 3) != 0) memset(Ai, 0, sizeof(Ai));
        CorrCtx c = {f, p, {x[0] - f->trans[0], x[1] - f->trans[1], x[2] - f->trans[2]}};
        BroydenResult r;
        broyden(g_corr, &c, 3, xh0, T0, Ai, &r);
        n_corr += r.g_evals + 1;                                              /* + the J-init skin eval */
        normalize_pts(f, r.x, spts[i]);
        memcpy(sT[i], r.T, sizeof(sT[i]));
        sconv[i] = (uint8_t)r.valid;
    }
    if (out->n_corr_evals) out->n_corr_evals[ray] = n_corr;
    if (out->sampled_pts) memcpy(out->sampled_pts + (size_t)ray * S * 3, spts, sizeof(float) * S * 3);
    if (out->sampled_dists) memcpy(out->sampled_dists + (size_t)ray * S, z, sizeof(float) * S);
    if (out->sampled_T) memcpy(out->sampled_T + (size_t)ray * S * 16, sT, sizeof(float) * S * 16);
    if (out->sampled_conv) memcpy(out->sampled_conv + (size_t)ray * S, sconv, S);

    /* ---- shading + compositing (get_rbg_value_vol_sdf, implicit_differentiable_renderer.py:261-396) */
    int len = 0;
    float cz[MAX_STEPS], crgb[MAX_STEPS][3], cden[MAX_STEPS];
    float beta = fabsf(f->beta);
    beta = fminf(fmaxf(beta, 1e-6f), 1e6f);
    const float inv_beta = 1.0f / beta;
    for (int i = 0; i < S; ++i) {
        if (!sconv[i]) continue;
        float feat[SDF_H], args[6 * SDF_H], n[3], view[3], nrm[3];
        const float s = sdf_forward(f, p, spts[i], feat, args);
        sdf_gradient(f, args, n);
        const float *T = sT[i];
        if (f->cano_view_dirs) {
            float A[9], Ai[9];
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A[r * 3 + c] = T[r * 4 + c];
            if (invert_n(A, Ai, 3) != 0) memset(Ai, 0, sizeof(Ai));
            for (int r = 0; r < 3; ++r) view[r] = Ai[r * 3] * -d[0] + Ai[r * 3 + 1] * -d[1] + Ai[r * 3 + 2] * -d[2];
            memcpy(nrm, n, sizeof(n));
        } else {
            for (int r = 0; r < 3; ++r) { view[r] = -d[r]; nrm[r] = T[r * 4] * n[0] + T[r * 4 + 1] * n[1] + T[r * 4 + 2] * n[2]; }
        }
        const float sm = sdf_to_metres(f, s);
        color_forward(f, p, spts[i], nrm, view, feat, crgb[len]);
        const float sg = (-sm > 0.0f) - (-sm < 0.0f);
        float den = inv_beta * (0.5f + 0.5f * sg * (1.0f - expf(-fabsf(-sm) * inv_beta)));
        cden[len] = fmaxf(den, 0.0f);
        cz[len] = z[i];
        ++len;
    }
    float rgb[3] = {0, 0, 0}, wsum = 0.0f, Tr = 1.0f;
    for (int k = 0; k < len; ++k) {
        float dz = (k + 1 < len) ? (cz[k + 1] - cz[k]) : (f->render_last_pt ? 1e10f : 1.0f / (float)S);     /* :379-385 */
        const float alpha = 1.0f - expf(-cden[k] * dz);
        const float w = alpha * Tr;
        Tr = Tr * (1.0f - alpha + 1e-7f);
        wsum += w;
        for (int c = 0; c < 3; ++c) rgb[c] += crgb[k][c] * w;
    }
    wsum = fminf(fmaxf(wsum, 0.0f), 1.0f);
    if (out->rgb) memcpy(out->rgb + (size_t)ray * 3, rgb, sizeof(rgb));
    if (out->vol_mask) out->vol_mask[ray] = (uint8_t)(len > 0);
    if (out->weights_sum) out->weights_sum[ray] = wsum;
    if (out->n_shaded) out->n_shaded[ray] = len;
    if (out->points_cam) {                                       /* implicit_differentiable_renderer.py:114-115,142-143,251 */
        const int surf = conv && fabsf(pnorm[0]) <= 1.0f && fabsf(pnorm[1]) <= 1.0f && fabsf(pnorm[2]) <= 1.0f;
        float pw[3], pc[3] = {0, 0, 0};
        for (int k = 0; k < 3; ++k) pw[k] = ((f->cam_loc[k] + dist * d[k]) - f->trans[k]) + f->trans[k];
        if (surf) for (int r = 0; r < 3; ++r) pc[r] = pw[0] * f->pose[r * 4] + pw[1] * f->pose[r * 4 + 1] + pw[2] * f->pose[r * 4 + 2] + f->pose[r * 4 + 3];
        memcpy(out->points_cam + (size_t)ray * 3, pc, sizeof(pc));
    }
}

/* ================================================================ exported entry points */
int arah_oracle_render(const OracleFrame *f, const float *ray_dirs, const float *near_far, int P, const OracleOut *out, int n_threads) {
    if (f->n_steps > MAX_STEPS || f->near_samples + 1 + f->far_samples > f->n_steps) return -1;
    Packed p;
    pack(f, &p);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int r = 0; r < P; ++r) render_ray(f, &p, ray_dirs + (size_t)r * 3, near_far[r * 2], near_far[r * 2 + 1], r, out, NULL);
    unpack(&p);
    return 0;
}

/* BodyRayTracing.forward(eval_mode=False) + the forward VALUES of the training render (the implicit-gradient correction of
 * implicit_differentiable_renderer.py:315-334 does not change values); u_* are the reference's three torch.rand draws. */
int arah_oracle_render_train(const OracleFrame *f, const float *ray_dirs, const float *near_far, int P, const OracleOut *out, int n_threads,
                             const float *u_all, const float *u_near, const float *u_far) {
    if (f->n_steps > MAX_STEPS || f->near_samples + 1 + f->far_samples > f->n_steps) return -1;
    if (!u_all || !u_near || (f->far_samples > 0 && !u_far)) return -2;
    Packed p;
    pack(f, &p);
    const TrainNoise tn = {u_all, u_near, u_far};
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int r = 0; r < P; ++r) render_ray(f, &p, ray_dirs + (size_t)r * 3, near_far[r * 2], near_far[r * 2 + 1], r, out, &tn);
    unpack(&p);
    return 0;
}

/* unit-level entry points (tests compare CUDA device functions against these) */
int arah_oracle_sdf(const OracleFrame *f, const float *xn, int n, float *sdf, float *grad /*n*3 or NULL*/, float *feat /*n*256 or NULL*/) {
    Packed p;
    pack(f, &p);
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        float args[6 * SDF_H];
        sdf[i] = sdf_forward(f, &p, xn + i * 3, feat ? feat + (size_t)i * SDF_H : NULL, args);
        if (grad) sdf_gradient(f, args, grad + i * 3);
    }
    unpack(&p);
    return 0;
}
int arah_oracle_skin(const OracleFrame *f, const float *x_hat, int n, float *w /*n*24*/, float *x_bar /*n*3*/, float *J /*n*9 or NULL*/) {
    Packed p;
    pack(f, &p);
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        float T[16];
        query_weights(f, &p, x_hat + i * 3, w + (size_t)i * NJ);
        forward_skinning(f, &p, x_hat + i * 3, x_bar + i * 3, T);
        if (J) forward_skinning_jac(f, &p, x_hat + i * 3, J + i * 9);
    }
    unpack(&p);
    return 0;
}
int arah_oracle_color(const OracleFrame *f, const float *xn, const float *normal, const float *view, const float *feat, int n, float *rgb) {
    Packed p;
    pack(f, &p);
#pragma omp parallel for
    for (int i = 0; i < n; ++i) color_forward(f, &p, xn + i * 3, normal + i * 3, view + i * 3, feat + (size_t)i * SDF_H, rgb + i * 3);
    unpack(&p);
    return 0;
}
int arah_oracle_knn(const OracleFrame *f, const float *x, int n, int32_t *idx, float *x_hat, float *T) {
#pragma omp parallel for
    for (int i = 0; i < n; ++i) {
        idx[i] = knn1(f, x + i * 3);
        knn_inverse_skinning(f, x + i * 3, x_hat + i * 3, T + (size_t)i * 16);
    }
    return 0;
}
int arah_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
