// arah_loss_core.h — per-element value and derivative of every term of the reference's training loss
// (im2mesh/metaavatar_render/renderer/loss.py:46-120), written once for device and host.
//
// arah_loss.cu's kernels are index wrappers around these functions; tests/native/host_loss.cpp compiles the same functions with
// g++ (test infrastructure) so that the arithmetic is checked against the reference's autograd gradients in the build container,
// which has no GPU.  Elements are evaluated in fp32 like the reference's tensors; sums are accumulated in fp64 by the callers.
// Derivative conventions are torch autograd's: d|x| = sign(x) with sign(0) = 0, d||g|| = g / ||g|| with 0 at the origin.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ARAH_LHD __host__ __device__ __forceinline__
#else
#define ARAH_LHD inline
#endif

namespace arah_loss {

enum { RGB_L1 = 0, RGB_MSE = 1, RGB_SMOOTH_L1 = 2 };
constexpr float SMOOTH_BETA = 0.1f;                      // loss.py:40

ARAH_LHD float sgn(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }

// nn.L1Loss / nn.MSELoss / nn.SmoothL1Loss(beta = 0.1), reduction 'sum' (loss.py:34-41): one colour channel
ARAH_LHD void rgb_elem(float pred, float gt, int type, float* v, float* dv) {
    const float d = pred - gt;
    if (type == RGB_MSE) { *v = d * d; *dv = 2.0f * d; }
    else if (type == RGB_SMOOTH_L1 && fabsf(d) < SMOOTH_BETA) { *v = 0.5f * d * d / SMOOTH_BETA; *dv = d / SMOOTH_BETA; }
    else if (type == RGB_SMOOTH_L1) { *v = fabsf(d) - 0.5f * SMOOTH_BETA; *dv = sgn(d); }
    else { *v = fabsf(d); *dv = sgn(d); }
}

// get_eikonal_loss (loss.py:88-94): | ||g|| - 1 |
ARAH_LHD void eik_point(const float* g, float* v, float* dv) {
    const float n = sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const float s = sgn(n - 1.0f);
    *v = fabsf(n - 1.0f);
    const float inv = n > 0.0f ? s / n : 0.0f;
    dv[0] = g[0] * inv; dv[1] = g[1] * inv; dv[2] = g[2] * inv;
}

// get_mask_loss_vol_sdf (loss.py:96-105).  model_outputs['sdf_output'] is [1, P] (implicit_differentiable_renderer.py:229-237), so
// `weights_output[off_surface_mask] - gt` is a 1-D vector and torch.norm(dim=-1) is the 2-norm of the WHOLE vector: the term is
// sqrt(sum_off (w_i - gt_i)^2) / N — one norm, not a per-ray sum.  Element: value (w - gt)^2, derivative (w - gt) (the caller
// divides by the norm; 0 at a zero norm, as torch's norm backward).
ARAH_LHD void mask_elem(float w, uint8_t body, float* v, float* dv) { const float d = w - (float)body; *v = d * d; *dv = d; }

// get_off_surface_loss (loss.py:107-109): exp(-100 s)
ARAH_LHD void off_point(float s, float* v, float* dv) { const float e = expf(-1e2f * s); *v = e; *dv = -1e2f * e; }

// get_inside_loss (loss.py:119-120): sigmoid(5000 s); torch's backward is y (1 - y) in fp32
ARAH_LHD void inside_point(float s, float* v, float* dv) {
    const float y = 1.0f / (1.0f + expf(-5e3f * s));
    *v = y; *dv = 5e3f * ((1.0f - y) * y);
}

// get_skinning_loss (loss.py:117-118): |pred - target|
ARAH_LHD void skin_elem(float p, float t, float* v, float* dv) { const float d = p - t; *v = fabsf(d); *dv = sgn(d); }

// the rgb term only counts rays the network hit and — when the body mask carries patch labels (max > 1) — not the border label 100
ARAH_LHD bool rgb_ray_counts(uint8_t net_mask, uint8_t body, unsigned body_max) { return net_mask != 0 && !(body_max > 1u && body == 100u); }

}  // namespace arah_loss
