// arah_loss.cu — the reference's training loss and its gradient as one fused pass ("fused loss reductions").
//
// `IDHRLoss.forward` (im2mesh/metaavatar_render/renderer/loss.py:122-200) is seven small masked reductions over the training
// step's outputs; in torch that is ~60 tiny kernels forward + backward and three host synchronisations
// (`network_body_mask.sum() == 0`, `body_mask.max() > 1`, `off_surface_mask.sum() == 0`).  arah_idhr_loss computes the nine terms
// AND d loss / d input for every differentiable input (each term is a sum of element-wise functions, so the derivative is
// element-wise too) in four launches without touching the host:
//   k_loss_pre     : max of the body mask (integer atomic: order-independent) — decides whether label 100 marks patch borders
//   k_loss_partial : fixed grid-stride partition over the rays / eikonal points / off-surface points / inside points / skinning
//                    weights / SDF parameters; fp32 elements, fp64 accumulation, shuffle + shared-memory tree per CTA
//   k_loss_finish  : one CTA reduces the partials in fixed order -> terms[9], the parameter-norm coefficient
//   k_loss_grads   : the same partition writes the weighted element-wise derivatives
// HBM-bound: every input read twice (value pass, derivative pass), every gradient written once; bit-reproducible.
// The arithmetic is in arah_loss_core.h (shared with the host test harness).
#ifndef ARAH_CUDA_EMU                      // tests/native/cuda_emu.h runs this file's source on the CPU (test infrastructure)
#include <cuda_runtime.h>
#define ARAH_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif
#include <stddef.h>
#include <stdint.h>
#include <string>

#include "../../include/arah_b200.h"
#include "arah_loss_core.h"

extern "C" int arah_internal_fail(int code, const char* msg);

namespace arah_loss {

constexpr int BLK = 256, NB = 148, NACC = 7;            // one CTA per SM: the whole problem is a few MB
enum { A_RGB = 0, A_EIK, A_MASK, A_OFF, A_INSIDE, A_SKIN, A_PARAMS };

struct Scratch {                                         // head of the caller's workspace
    unsigned body_max;
    float params_coef;                                   // params_weight / (||p|| n_params), 0 for a zero norm
    float mask_coef;                                     // mask_weight / (||w - gt|| N), 0 for a zero norm
    float pad_;
    double partial[NB * NACC];
};

struct Args {                                            // by-value kernel argument
    ArahLossConfig cfg;
    ArahLossInputs in;
    ArahLossGrads g;
    long long n_params_total;
};

__global__ void __launch_bounds__(BLK) k_loss_pre(Args a, Scratch* s) {
    const int i = blockIdx.x * BLK + threadIdx.x;
    if (i >= a.in.n_rays) return;
    atomicMax(&s->body_max, (unsigned)a.in.body_mask[i]);
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                                     // sh is reused across calls
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = threadIdx.x < BLK / 32 ? sh[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;                                            // valid in thread 0
}

__global__ void __launch_bounds__(BLK) k_loss_partial(Args a, Scratch* s) {
    __shared__ double sh[BLK / 32];
    const unsigned body_max = s->body_max;
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
    const long long t0 = (long long)blockIdx.x * BLK + threadIdx.x, stride = (long long)NB * BLK;
    const ArahLossInputs& in = a.in;
    float v, dv, dv3[3];
    for (long long i = t0; i < in.n_rays; i += stride) {
        if (a.cfg.rgb_weight > 0.0f && rgb_ray_counts(in.network_body_mask[i], in.body_mask[i], body_max))
            for (int c = 0; c < 3; ++c) { rgb_elem(in.rgb_values[i * 3 + c], in.rgb_gt[i * 3 + c], a.cfg.rgb_loss_type, &v, &dv); acc[A_RGB] += (double)v; }
        if (a.cfg.mask_weight > 0.0f && in.off_surface_mask[i]) { mask_elem(in.sdf_output[i], in.body_mask[i], &v, &dv); acc[A_MASK] += (double)v; }
    }
    if (a.cfg.eikonal_weight > 0.0f)
        for (long long i = t0; i < in.n_eikonal; i += stride) { eik_point(in.grad_theta + i * 3, &v, dv3); acc[A_EIK] += (double)v; }
    if (a.cfg.off_surface_weight > 0.0f)
        for (long long i = t0; i < in.n_off; i += stride) { off_point(in.off_surface_sdf[i], &v, &dv); acc[A_OFF] += (double)v; }
    if (a.cfg.inside_weight > 0.0f)
        for (long long i = t0; i < in.n_inside; i += stride) { inside_point(in.inside_sdf[i], &v, &dv); acc[A_INSIDE] += (double)v; }
    if (a.cfg.skinning_weight > 0.0f)
        for (long long i = t0; i < (long long)in.n_skin * in.n_joints; i += stride) { skin_elem(in.pred_weights[i], in.sampled_weights[i], &v, &dv); acc[A_SKIN] += (double)v; }
    if (a.cfg.params_weight > 0.0f)
        for (int t = 0; t < in.n_param_tensors; ++t)
            for (long long i = t0; i < in.sdf_params_count[t]; i += stride) { const float p = __ldg(in.sdf_params[t] + i); acc[A_PARAMS] += (double)p * (double)p; }
    for (int k = 0; k < NACC; ++k) {
        const double t = block_sum(acc[k], sh);
        if (threadIdx.x == 0) s->partial[blockIdx.x * NACC + k] = t;
    }
}

__global__ void __launch_bounds__(BLK) k_loss_finish(Args a, Scratch* s, float* terms) {
    __shared__ double sh[BLK / 32];
    __shared__ double tot[NACC];
    for (int k = 0; k < NACC; ++k) {
        double v = 0.0;
        for (int b = threadIdx.x; b < NB; b += BLK) v += s->partial[b * NACC + k];
        const double t = block_sum(v, sh);
        if (threadIdx.x == 0) tot[k] = t;
    }
    if (threadIdx.x != 0) return;
    const double N = (double)a.in.n_rays;                 // float(body_mask.numel()) == float(network_body_mask.numel())
    const ArahLossConfig& c = a.cfg;
    const double rgb = c.rgb_weight > 0.0f && N > 0 ? tot[A_RGB] / N : 0.0;
    const double eik = c.eikonal_weight > 0.0f && a.in.n_eikonal > 0 ? tot[A_EIK] / N : 0.0;
    const double mnorm = sqrt(tot[A_MASK]);
    const double msk = c.mask_weight > 0.0f ? mnorm / N : 0.0;
    s->mask_coef = (c.mask_weight > 0.0f && mnorm > 0.0) ? (float)((double)c.mask_weight / (mnorm * N)) : 0.0f;
    const double off = c.off_surface_weight > 0.0f ? tot[A_OFF] / N : 0.0;
    const double ins = c.inside_weight > 0.0f ? tot[A_INSIDE] / N : 0.0;
    const double skn = c.skinning_weight > 0.0f ? tot[A_SKIN] / (double)a.in.n_skin : 0.0;
    const double norm = sqrt(tot[A_PARAMS]);
    const double prm = c.params_weight > 0.0f ? norm / (double)a.n_params_total : 0.0;
    s->params_coef = (c.params_weight > 0.0f && norm > 0.0) ? (float)((double)c.params_weight / (norm * (double)a.n_params_total)) : 0.0f;
    terms[1] = (float)rgb; terms[2] = 0.0f; terms[3] = (float)eik; terms[4] = (float)msk; terms[5] = (float)off; terms[6] = (float)ins;
    terms[7] = (float)prm; terms[8] = (float)skn;
    terms[0] = (float)((double)c.rgb_weight * rgb + (double)c.eikonal_weight * eik + (double)c.mask_weight * msk + (double)c.off_surface_weight * off +
                       (double)c.inside_weight * ins + (double)c.params_weight * prm + (double)c.skinning_weight * skn);
}

__global__ void __launch_bounds__(BLK) k_loss_grads(Args a, const Scratch* s) {
    const unsigned body_max = s->body_max;
    const float mask_coef = s->mask_coef;
    const long long t0 = (long long)blockIdx.x * BLK + threadIdx.x, stride = (long long)NB * BLK;
    const ArahLossInputs& in = a.in;
    const ArahLossConfig& c = a.cfg;
    const float inv_n = in.n_rays > 0 ? 1.0f / (float)in.n_rays : 0.0f;
    float v, dv, dv3[3];
    for (long long i = t0; i < in.n_rays; i += stride) {
        if (a.g.rgb_values) {
            const bool on = c.rgb_weight > 0.0f && rgb_ray_counts(in.network_body_mask[i], in.body_mask[i], body_max);
            for (int ch = 0; ch < 3; ++ch) {
                dv = 0.0f;
                if (on) rgb_elem(in.rgb_values[i * 3 + ch], in.rgb_gt[i * 3 + ch], c.rgb_loss_type, &v, &dv);
                a.g.rgb_values[i * 3 + ch] = on ? c.rgb_weight * dv * inv_n : 0.0f;
            }
        }
        if (a.g.sdf_output) {
            dv = 0.0f;
            if (c.mask_weight > 0.0f && in.off_surface_mask[i]) mask_elem(in.sdf_output[i], in.body_mask[i], &v, &dv);
            a.g.sdf_output[i] = mask_coef * dv;
        }
    }
    if (a.g.grad_theta)
        for (long long i = t0; i < in.n_eikonal; i += stride) {
            dv3[0] = dv3[1] = dv3[2] = 0.0f;
            if (c.eikonal_weight > 0.0f) eik_point(in.grad_theta + i * 3, &v, dv3);
            for (int k = 0; k < 3; ++k) a.g.grad_theta[i * 3 + k] = c.eikonal_weight * dv3[k] * inv_n;
        }
    if (a.g.off_surface_sdf)
        for (long long i = t0; i < in.n_off; i += stride) {
            dv = 0.0f;
            if (c.off_surface_weight > 0.0f) off_point(in.off_surface_sdf[i], &v, &dv);
            a.g.off_surface_sdf[i] = c.off_surface_weight * dv * inv_n;
        }
    if (a.g.inside_sdf)
        for (long long i = t0; i < in.n_inside; i += stride) {
            dv = 0.0f;
            if (c.inside_weight > 0.0f) inside_point(in.inside_sdf[i], &v, &dv);
            a.g.inside_sdf[i] = c.inside_weight * dv * inv_n;
        }
    if (a.g.pred_weights) {
        const float inv_s = in.n_skin > 0 ? 1.0f / (float)in.n_skin : 0.0f;
        for (long long i = t0; i < (long long)in.n_skin * in.n_joints; i += stride) {
            dv = 0.0f;
            if (c.skinning_weight > 0.0f) skin_elem(in.pred_weights[i], in.sampled_weights[i], &v, &dv);
            a.g.pred_weights[i] = c.skinning_weight * dv * inv_s;
        }
    }
    const float coef = s->params_coef;
    for (int t = 0; t < in.n_param_tensors; ++t)
        if (a.g.sdf_params[t])
            for (long long i = t0; i < in.sdf_params_count[t]; i += stride) a.g.sdf_params[t][i] = coef * __ldg(in.sdf_params[t] + i);
}

}  // namespace arah_loss

using namespace arah_loss;

#define LCU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return arah_internal_fail(ARAH_ECUDA, (std::string(#x) + ": " + cudaGetErrorString(e_)).c_str()); } while (0)

extern "C" size_t arah_idhr_loss_workspace(void) { return (sizeof(Scratch) + 255) & ~(size_t)255; }

extern "C" int arah_idhr_loss(const ArahLossConfig* cfg, const ArahLossInputs* in, float* terms, const ArahLossGrads* grads, void* workspace,
                              size_t workspace_bytes, void* stream) {
    if (!cfg || !in || !terms || !workspace) return arah_internal_fail(ARAH_EINVAL, "null argument");
    if (workspace_bytes < arah_idhr_loss_workspace()) return arah_internal_fail(ARAH_EINVAL, "workspace smaller than arah_idhr_loss_workspace()");
    if (cfg->perceptual_weight > 0.0f) return arah_internal_fail(ARAH_EINVAL, "the perceptual (LPIPS) term is not part of this path: perceptual_weight must be 0");
    if (cfg->rgb_loss_type < RGB_L1 || cfg->rgb_loss_type > RGB_SMOOTH_L1) return arah_internal_fail(ARAH_EINVAL, "rgb_loss_type must be 0 (l1), 1 (mse) or 2 (smoothed_l1)");
    if (in->n_rays <= 0 || !in->body_mask) return arah_internal_fail(ARAH_EINVAL, "n_rays <= 0 or no body_mask (every term is divided by body_mask.numel())");
    if (in->n_eikonal < 0 || in->n_off < 0 || in->n_inside < 0 || in->n_skin < 0 || in->n_param_tensors < 0 || in->n_param_tensors > ARAH_LOSS_MAX_PARAM_TENSORS)
        return arah_internal_fail(ARAH_EINVAL, "bad count");
    if (cfg->rgb_weight > 0.0f && (!in->rgb_values || !in->rgb_gt || !in->network_body_mask)) return arah_internal_fail(ARAH_EINVAL, "rgb term: missing input");
    if (cfg->mask_weight > 0.0f && (!in->sdf_output || !in->off_surface_mask)) return arah_internal_fail(ARAH_EINVAL, "mask term: missing input");
    if (cfg->eikonal_weight > 0.0f && in->n_eikonal > 0 && !in->grad_theta) return arah_internal_fail(ARAH_EINVAL, "eikonal term: missing input");
    if (cfg->off_surface_weight > 0.0f && in->n_off > 0 && !in->off_surface_sdf) return arah_internal_fail(ARAH_EINVAL, "off-surface term: missing input");
    if (cfg->inside_weight > 0.0f && in->n_inside > 0 && !in->inside_sdf) return arah_internal_fail(ARAH_EINVAL, "inside term: missing input");
    if (cfg->skinning_weight > 0.0f && (in->n_skin <= 0 || in->n_joints <= 0 || !in->pred_weights || !in->sampled_weights))
        return arah_internal_fail(ARAH_EINVAL, "skinning term: missing input");
    Args a;
    a.cfg = *cfg; a.in = *in;
    a.n_params_total = 0;
    for (int t = 0; t < in->n_param_tensors; ++t) {
        if (in->sdf_params_count[t] < 0 || (in->sdf_params_count[t] > 0 && !in->sdf_params[t])) return arah_internal_fail(ARAH_EINVAL, "params term: bad tensor");
        a.n_params_total += in->sdf_params_count[t];
    }
    if (cfg->params_weight > 0.0f && a.n_params_total <= 0) return arah_internal_fail(ARAH_EINVAL, "params term: no parameters");
    if (grads) a.g = *grads; else { ArahLossGrads z = {}; a.g = z; }
    // terms whose weight is 0 never read their inputs; make the loops empty rather than trusting unused counts
    if (!(cfg->eikonal_weight > 0.0f) && !a.g.grad_theta) a.in.n_eikonal = 0;
    if (!(cfg->skinning_weight > 0.0f) && !a.g.pred_weights) a.in.n_skin = 0;
    cudaStream_t st = (cudaStream_t)stream;
    Scratch* s = (Scratch*)workspace;
    LCU(cudaMemsetAsync(s, 0, offsetof(Scratch, partial), st));
    ARAH_LAUNCH(k_loss_pre, (unsigned)((in->n_rays + BLK - 1) / BLK), BLK, st, a, s);
    ARAH_LAUNCH(k_loss_partial, NB, BLK, st, a, s);
    ARAH_LAUNCH(k_loss_finish, 1, BLK, st, a, s, terms);
    if (grads) ARAH_LAUNCH(k_loss_grads, NB, BLK, st, a, s);
    LCU(cudaGetLastError());
    return ARAH_OK;
}
