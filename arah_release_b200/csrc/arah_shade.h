// arah_shade.h — host-side entry points of arah_shade.cu (k_shade16: gradient + colour pass on tcgen05 kind::f16) for arah_api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "arah_work.cuh"

namespace arah {

// fp16 chunk images the shading kernel streams besides the forward SDF images of the root engine (arah_root.h: SdfF16Dev.hi)
struct Shade16Dev {
    void* bwd;       // SHADE16_BWD_DEV_BYTES: W_l^T of SDF layers 1..5
    void* col;       // SHADE16_COL_DEV_BYTES: colour layers lin0 | lin1 | lin2 | lin3b | lin3a | lin4
};
constexpr size_t SHADE16_BWD_DEV_BYTES = 5 * 131072;
constexpr size_t SHADE16_COL_DEV_BYTES = (size_t)(5 + 4 + 2 + 5 + 4) * 32768 + 4 * 16384;
struct Shade16Host {     // device pointers into the frame arena (fp32 parameters the epilogues read)
    const float* sdf_Wt0; const float* sdf_W0; const float* sdf_F; const float* sdf_G; const float* sdf_scale; const void* sdf_fwd_hi;
    const float* sdf_w6; const float* sdf_b6; const float* col_W5; const float* col_b[6];
};
size_t shade16_scratch_bytes_per_cta();
cudaError_t shade16_init();
// colour weights in the reference's layout ([out][in], input order [x 3 | PE 27 | n 3 | feat 256 | latent], din columns)
cudaError_t shade16_pack(const float* const sdf_W[7], const float* const col_W[6], int din, const Shade16Dev& dst, cudaStream_t st,
                         long long* launches);
// samples w.shade_list[0 .. counters[w.shade_ctr]) -> w.smp_rgb (and w.smp_sdf unless w.shade_keep_sdf)
cudaError_t shade16_launch(const FrameParams& fp, const Shade16Host& sh, const Shade16Dev& img, const Work& w, unsigned grid, cudaStream_t st,
                           long long* launches);

}  // namespace arah
