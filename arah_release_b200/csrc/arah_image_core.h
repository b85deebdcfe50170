// arah_image_core.h — per-element arithmetic of the image-space tail (arah_image.cu), written once for device and host.
//
// The kernels of arah_image.cu are thin index wrappers around these functions.  tests/native/host_image.cpp compiles the SAME
// functions with g++ (test infrastructure, like tests/native/host_train.cpp) so that their arithmetic is checked against the
// numpy oracle / the reference's golden images in the build container, which has no GPU.  libarah_b200.so has no host path.
// Every fp32 operation is rounded once (explicit _rn intrinsics on the device: no FMA contraction; -ffp-contract=off on the
// host), in the order the oracle (oracle/images_oracle.py) uses.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ARAH_HD __host__ __device__ __forceinline__
#else
#define ARAH_HD inline
#endif

namespace arah_img {

#if defined(__CUDA_ARCH__)
ARAH_HD float mul(float a, float b) { return __fmul_rn(a, b); }
ARAH_HD float add(float a, float b) { return __fadd_rn(a, b); }
ARAH_HD float sub(float a, float b) { return __fsub_rn(a, b); }
ARAH_HD float dvd(float a, float b) { return __fdiv_rn(a, b); }
ARAH_HD float sqr(float a) { return __fsqrt_rn(a); }
ARAH_HD uint32_t f2u(float f) { return __float_as_uint(f); }
#else
ARAH_HD float mul(float a, float b) { return a * b; }
ARAH_HD float add(float a, float b) { return a + b; }
ARAH_HD float sub(float a, float b) { return a - b; }
ARAH_HD float dvd(float a, float b) { return a / b; }
ARAH_HD float sqr(float a) { return sqrtf(a); }
ARAH_HD uint32_t f2u(float f) { union { float f; uint32_t u; } c; c.f = f; return c.u; }
#endif
ARAH_HD float fmin3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
ARAH_HD float fmax3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }

constexpr float K_EPS = 1e-8f;                                       // pytorch3d kEpsilon
constexpr unsigned long long KEY_EMPTY = 0xffffffffffffffffull;

// (n + 1) / 2 clipped to [0, 1] with NaN -> -1 first (lightning_model.py:204-205; models/__init__.py:263,286)
ARAH_HD float to_unit(float v) {
    if (v != v) v = -1.0f;
    return fminf(fmaxf(dvd(add(v, 1.0f), 2.0f), 0.0f), 1.0f);
}

// Finite-difference normal of the camera-space point image at pixel (y, x)          lightning_model.py:190-205
// pts [H][W][3]; out[3] in [0, 1].
ARAH_HD void depth_normal(const float* pts, int H, int W, int y, int x, float* out) {
    const float* p = pts + ((size_t)y * W + x) * 3;
    float nx = 0.0f, ny = 0.0f;
    if (x < W - 1) nx = -dvd(sub(p[3 + 2], p[2]), sub(p[3 + 0], p[0]));                       // -zx
    if (y < H - 1) { const float* q = p + (size_t)W * 3; ny = -dvd(sub(q[2], p[2]), sub(q[1], p[1])); }   // -zy
    const float n = sqr(add(add(mul(nx, nx), mul(ny, ny)), 1.0f));
    out[0] = to_unit(dvd(nx, n));
    out[1] = to_unit(dvd(ny, n));
    out[2] = to_unit(dvd(1.0f, n));
}

// pytorch3d camera (row vectors): view = X R + T, ndc = (fx x + px z, fy y + py z) / z, z stays the view depth.
struct Camera { float R[9]; float T[3]; float fx, fy, px, py; };

ARAH_HD void project(const Camera& c, const float* v, float* out) {
    float view[3];
    for (int k = 0; k < 3; ++k)
        view[k] = add(add(add(mul(v[0], c.R[0 * 3 + k]), mul(v[1], c.R[1 * 3 + k])), mul(v[2], c.R[2 * 3 + k])), c.T[k]);
    out[0] = dvd(add(mul(c.fx, view[0]), mul(c.px, view[2])), view[2]);
    out[1] = dvd(add(mul(c.fy, view[1]), mul(c.py, view[2])), view[2]);
    out[2] = view[2];
}

// centre of pixel i (counted from the +ndc end: i = S1 - 1 - index) along an axis of S1 pixels; S2 = the other axis
ARAH_HD float pix_to_ndc(int i, int S1, int S2) {
    const float rng = S1 > S2 ? dvd(mul(2.0f, (float)S1), (float)S2) : 2.0f;
    const float off = dvd(rng, 2.0f);
    return add(-off, dvd(add(mul(rng, (float)i), off), (float)S1));
}

ARAH_HD float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return sub(mul(sub(px, ax), sub(by, ay)), mul(sub(py, ay), sub(bx, ax)));
}

struct FaceSetup {                    // per-face constants of the pixel test
    float x0, y0, z0, x1, y1, z1, x2, y2, z2, barea;
    float xmin, xmax, ymin, ymax;
    bool drawable;
};

ARAH_HD bool finite3(const float* a) { return fabsf(a[0]) <= 3.0e38f && fabsf(a[1]) <= 3.0e38f && fabsf(a[2]) <= 3.0e38f; }

ARAH_HD FaceSetup face_setup(const float* a, const float* b, const float* c) {
    FaceSetup s;
    s.x0 = a[0]; s.y0 = a[1]; s.z0 = a[2]; s.x1 = b[0]; s.y1 = b[1]; s.z1 = b[2]; s.x2 = c[0]; s.y2 = c[1]; s.z2 = c[2];
    s.drawable = finite3(a) && finite3(b) && finite3(c);
    if (s.drawable && fmin3(s.z0, s.z1, s.z2) < K_EPS) s.drawable = false;           // touches the camera plane: not drawn
    const float area = edge_fn(s.x0, s.y0, s.x1, s.y1, s.x2, s.y2);
    if (s.drawable && area <= K_EPS && area >= -K_EPS) s.drawable = false;            // zero-area face
    s.barea = add(edge_fn(s.x2, s.y2, s.x0, s.y0, s.x1, s.y1), K_EPS);
    s.xmin = fmin3(s.x0, s.x1, s.x2); s.xmax = fmax3(s.x0, s.x1, s.x2);
    s.ymin = fmin3(s.y0, s.y1, s.y2); s.ymax = fmax3(s.y0, s.y1, s.y2);
    return s;
}

// CheckPixelInsideFace for blur_radius 0, perspective-correct depth; true -> *pz is the depth of the face at the pixel centre
ARAH_HD bool face_covers(const FaceSetup& s, float px, float py, float* pz) {
    if (px < s.xmin || px > s.xmax || py < s.ymin || py > s.ymax) return false;
    const float w0 = dvd(edge_fn(px, py, s.x1, s.y1, s.x2, s.y2), s.barea);
    const float w1 = dvd(edge_fn(px, py, s.x2, s.y2, s.x0, s.y0), s.barea);
    const float w2 = dvd(edge_fn(px, py, s.x0, s.y0, s.x1, s.y1), s.barea);
    if (!(w0 > 0.0f && w1 > 0.0f && w2 > 0.0f)) return false;
    const float t0 = mul(mul(w0, s.z1), s.z2), t1 = mul(mul(s.z0, w1), s.z2), t2 = mul(mul(s.z0, s.z1), w2);
    const float den = fmaxf(add(add(t0, t1), t2), K_EPS);
    const float z = add(add(mul(dvd(t0, den), s.z0), mul(dvd(t1, den), s.z1)), mul(dvd(t2, den), s.z2));
    if (!(z >= 0.0f)) return false;
    *pz = z;
    return true;
}

// depth-then-face-index key: atomicMin over it = nearest face, lowest index on ties (depth >= 0: its bits are monotonic)
ARAH_HD unsigned long long raster_key(float pz, int face) { return ((unsigned long long)f2u(pz) << 32) | (unsigned)face; }

// Conservative pixel-index range [lo, hi] whose centres can lie in [vmin, vmax] along an axis of S1 pixels (index k <-> pix_to_ndc(S1-1-k)).
ARAH_HD void pixel_range(float vmin, float vmax, int S1, int S2, int* lo, int* hi) {
    const float rng = S1 > S2 ? 2.0f * (float)S1 / (float)S2 : 2.0f;
    const float off = 0.5f * rng;
    // ndc(k) = -off + (rng (S1-1-k) + off) / S1  ->  k = S1 - 1 - ((ndc + off) S1 - off) / rng
    const float k_hi = (float)(S1 - 1) - (((vmin + off) * (float)S1 - off) / rng);
    const float k_lo = (float)(S1 - 1) - (((vmax + off) * (float)S1 - off) / rng);
    float a = floorf(k_lo) - 1.0f, b = ceilf(k_hi) + 1.0f;
    if (!(a > 0.0f)) a = 0.0f;
    if (!(b < (float)(S1 - 1))) b = (float)(S1 - 1);
    *lo = (int)a; *hi = (int)b;                      // empty when lo > hi (face outside the image)
}

// Meshes.faces_normals_packed(): cross(v1 - v0, v2 - v0) / max(|.|, 1e-6); then sign, optional rotation (rows of rot), to_unit
ARAH_HD void face_normal_pixel(const float* a, const float* b, const float* c, float sign, const float* rot, float* out) {
    const float e0 = sub(b[0], a[0]), e1 = sub(b[1], a[1]), e2 = sub(b[2], a[2]);
    const float g0 = sub(c[0], a[0]), g1 = sub(c[1], a[1]), g2 = sub(c[2], a[2]);
    const float c0 = sub(mul(e1, g2), mul(e2, g1)), c1 = sub(mul(e2, g0), mul(e0, g2)), c2 = sub(mul(e0, g1), mul(e1, g0));
    const float n = fmaxf(sqr(add(add(mul(c0, c0), mul(c1, c1)), mul(c2, c2))), 1e-6f);
    float v[3] = {mul(sign, dvd(c0, n)), mul(sign, dvd(c1, n)), mul(sign, dvd(c2, n))};
    if (rot) {
        float r[3];
        for (int i = 0; i < 3; ++i) r[i] = add(add(mul(rot[i * 3 + 0], v[0]), mul(rot[i * 3 + 1], v[1])), mul(rot[i * 3 + 2], v[2]));
        v[0] = r[0]; v[1] = r[1]; v[2] = r[2];
    }
    out[0] = to_unit(v[0]); out[1] = to_unit(v[1]); out[2] = to_unit(v[2]);
}

// SSIM of one channel at one pixel whose 7x7 window lies inside the image: skimage.metrics.structural_similarity with its defaults
// as im2mesh/utils/eval.py:11-19 calls it (uniform window, sample covariance, float64 arithmetic, data_range 2 for float images,
// K1 0.01, K2 0.03).  X, Y: [H][W][3] float32 images, (y, x) the window centre.
constexpr int SSIM_WIN = 7, SSIM_PAD = 3;
ARAH_HD double ssim_pixel(const float* X, const float* Y, int W, int y, int x, int ch) {
    double sx = 0.0, sy = 0.0, sxx = 0.0, syy = 0.0, sxy = 0.0;
    for (int dy = -SSIM_PAD; dy <= SSIM_PAD; ++dy)
        for (int dx = -SSIM_PAD; dx <= SSIM_PAD; ++dx) {
            const size_t i = ((size_t)(y + dy) * W + (x + dx)) * 3 + ch;
            const double a = (double)X[i], b = (double)Y[i];
            sx += a; sy += b; sxx += a * a; syy += b * b; sxy += a * b;
        }
    const double n = (double)(SSIM_WIN * SSIM_WIN), cov = n / (n - 1.0);
    const double ux = sx / n, uy = sy / n;
    const double vx = cov * (sxx / n - ux * ux), vy = cov * (syy / n - uy * uy), vxy = cov * (sxy / n - ux * uy);
    const double C1 = (0.01 * 2.0) * (0.01 * 2.0), C2 = (0.03 * 2.0) * (0.03 * 2.0);
    return ((2.0 * ux * uy + C1) * (2.0 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
}

}  // namespace arah_img
