// TEST INFRASTRUCTURE — host instantiation of csrc/arah_loss_core.h (the per-element arithmetic of arah_loss.cu).
// Serial counterpart of k_loss_pre / k_loss_partial / k_loss_finish / k_loss_grads: same core calls, same conditions, loops instead
// of grids.  tests/test_loss_host.py compares it with the unmodified reference's terms and autograd gradients.  Never loaded by the product.
//   g++ -O2 -std=c++17 -shared -fPIC -o libarah_loss_host.so host_loss.cpp
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/arah_b200.h"
#include "../../arah_release_b200/csrc/arah_loss_core.h"

using namespace arah_loss;

extern "C" int host_idhr_loss(const ArahLossConfig* cfg, const ArahLossInputs* inp, float* terms, const ArahLossGrads* grads) {
    const ArahLossConfig& c = *cfg;
    ArahLossInputs in = *inp;
    ArahLossGrads g;
    memset(&g, 0, sizeof(g));
    if (grads) g = *grads;
    if (c.perceptual_weight > 0.0f || in.n_rays <= 0 || !in.body_mask) return -1;
    if (!(c.eikonal_weight > 0.0f) && !g.grad_theta) in.n_eikonal = 0;
    if (!(c.skinning_weight > 0.0f) && !g.pred_weights) in.n_skin = 0;
    // k_loss_pre
    unsigned body_max = 0;
    for (int i = 0; i < in.n_rays; ++i) if (in.body_mask[i] > body_max) body_max = in.body_mask[i];
    // k_loss_partial
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    long long n_params_total = 0;
    for (int t = 0; t < in.n_param_tensors; ++t) n_params_total += in.sdf_params_count[t];
    float v, dv, dv3[3];
    for (long long i = 0; i < in.n_rays; ++i) {
        if (c.rgb_weight > 0.0f && rgb_ray_counts(in.network_body_mask[i], in.body_mask[i], body_max))
            for (int ch = 0; ch < 3; ++ch) { rgb_elem(in.rgb_values[i * 3 + ch], in.rgb_gt[i * 3 + ch], c.rgb_loss_type, &v, &dv); acc[0] += (double)v; }
        if (c.mask_weight > 0.0f && in.off_surface_mask[i]) { mask_elem(in.sdf_output[i], in.body_mask[i], &v, &dv); acc[2] += (double)v; }
    }
    if (c.eikonal_weight > 0.0f) for (long long i = 0; i < in.n_eikonal; ++i) { eik_point(in.grad_theta + i * 3, &v, dv3); acc[1] += (double)v; }
    if (c.off_surface_weight > 0.0f) for (long long i = 0; i < in.n_off; ++i) { off_point(in.off_surface_sdf[i], &v, &dv); acc[3] += (double)v; }
    if (c.inside_weight > 0.0f) for (long long i = 0; i < in.n_inside; ++i) { inside_point(in.inside_sdf[i], &v, &dv); acc[4] += (double)v; }
    if (c.skinning_weight > 0.0f)
        for (long long i = 0; i < (long long)in.n_skin * in.n_joints; ++i) { skin_elem(in.pred_weights[i], in.sampled_weights[i], &v, &dv); acc[5] += (double)v; }
    if (c.params_weight > 0.0f)
        for (int t = 0; t < in.n_param_tensors; ++t)
            for (long long i = 0; i < in.sdf_params_count[t]; ++i) { const float p = in.sdf_params[t][i]; acc[6] += (double)p * (double)p; }
    // k_loss_finish
    const double N = (double)in.n_rays;
    const double rgb = c.rgb_weight > 0.0f && N > 0 ? acc[0] / N : 0.0;
    const double eik = c.eikonal_weight > 0.0f && in.n_eikonal > 0 ? acc[1] / N : 0.0;
    const double mnorm = sqrt(acc[2]);
    const double msk = c.mask_weight > 0.0f ? mnorm / N : 0.0;
    const float mask_coef = (c.mask_weight > 0.0f && mnorm > 0.0) ? (float)((double)c.mask_weight / (mnorm * N)) : 0.0f;
    const double off = c.off_surface_weight > 0.0f ? acc[3] / N : 0.0;
    const double ins = c.inside_weight > 0.0f ? acc[4] / N : 0.0;
    const double skn = c.skinning_weight > 0.0f ? acc[5] / (double)in.n_skin : 0.0;
    const double norm = sqrt(acc[6]);
    const double prm = c.params_weight > 0.0f ? norm / (double)n_params_total : 0.0;
    const float coef = (c.params_weight > 0.0f && norm > 0.0) ? (float)((double)c.params_weight / (norm * (double)n_params_total)) : 0.0f;
    terms[1] = (float)rgb; terms[2] = 0.0f; terms[3] = (float)eik; terms[4] = (float)msk; terms[5] = (float)off; terms[6] = (float)ins;
    terms[7] = (float)prm; terms[8] = (float)skn;
    terms[0] = (float)((double)c.rgb_weight * rgb + (double)c.eikonal_weight * eik + (double)c.mask_weight * msk + (double)c.off_surface_weight * off +
                       (double)c.inside_weight * ins + (double)c.params_weight * prm + (double)c.skinning_weight * skn);
    if (!grads) return 0;
    // k_loss_grads
    const float inv_n = in.n_rays > 0 ? 1.0f / (float)in.n_rays : 0.0f;
    for (long long i = 0; i < in.n_rays; ++i) {
        if (g.rgb_values) {
            const bool on = c.rgb_weight > 0.0f && rgb_ray_counts(in.network_body_mask[i], in.body_mask[i], body_max);
            for (int ch = 0; ch < 3; ++ch) {
                dv = 0.0f;
                if (on) rgb_elem(in.rgb_values[i * 3 + ch], in.rgb_gt[i * 3 + ch], c.rgb_loss_type, &v, &dv);
                g.rgb_values[i * 3 + ch] = on ? c.rgb_weight * dv * inv_n : 0.0f;
            }
        }
        if (g.sdf_output) {
            dv = 0.0f;
            if (c.mask_weight > 0.0f && in.off_surface_mask[i]) mask_elem(in.sdf_output[i], in.body_mask[i], &v, &dv);
            g.sdf_output[i] = mask_coef * dv;
        }
    }
    if (g.grad_theta)
        for (long long i = 0; i < in.n_eikonal; ++i) {
            dv3[0] = dv3[1] = dv3[2] = 0.0f;
            if (c.eikonal_weight > 0.0f) eik_point(in.grad_theta + i * 3, &v, dv3);
            for (int k = 0; k < 3; ++k) g.grad_theta[i * 3 + k] = c.eikonal_weight * dv3[k] * inv_n;
        }
    if (g.off_surface_sdf)
        for (long long i = 0; i < in.n_off; ++i) { dv = 0.0f; if (c.off_surface_weight > 0.0f) off_point(in.off_surface_sdf[i], &v, &dv); g.off_surface_sdf[i] = c.off_surface_weight * dv * inv_n; }
    if (g.inside_sdf)
        for (long long i = 0; i < in.n_inside; ++i) { dv = 0.0f; if (c.inside_weight > 0.0f) inside_point(in.inside_sdf[i], &v, &dv); g.inside_sdf[i] = c.inside_weight * dv * inv_n; }
    if (g.pred_weights) {
        const float inv_s = in.n_skin > 0 ? 1.0f / (float)in.n_skin : 0.0f;
        for (long long i = 0; i < (long long)in.n_skin * in.n_joints; ++i) {
            dv = 0.0f;
            if (c.skinning_weight > 0.0f) skin_elem(in.pred_weights[i], in.sampled_weights[i], &v, &dv);
            g.pred_weights[i] = c.skinning_weight * dv * inv_s;
        }
    }
    for (int t = 0; t < in.n_param_tensors; ++t)
        if (g.sdf_params[t]) for (long long i = 0; i < in.sdf_params_count[t]; ++i) g.sdf_params[t][i] = coef * in.sdf_params[t][i];
    return 0;
}
