// arah_root.h — host-side entry points of arah_root.cu (the persistent root-finding kernels and their weight packing) for
// arah_api.cu.  Plain declarations: the kernels themselves live in arah_root.cu so that the two translation units compile in
// parallel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "arah_work.cuh"

namespace arah {

// device images of the skinning MLP for k_corr_persist (arah_corr_p.cuh)
struct SkinF16Dev {
    void* hi;        // 3 x 32 KB + 8 KB
    void* lo;        // same
    float* scale;    // [4][2]
};
constexpr size_t SKIN_F16_IMAGE_BYTES = 3 * 32768 + 8192;

// pack layers 1..4 of the skinning MLP (reference layout [out][in], fp32) into scaled fp16 hi / lo chunk images
cudaError_t root_pack_skin_f16(const float* const W[5], const SkinF16Dev& dst, cudaStream_t st, long long* launches);
// correspondence search of all on-samples (seeds in w.corr_seed) + the list of converged samples
cudaError_t root_corr_persist(const FrameParams& fp, const float* skin_Wt0, const float* const skin_b[5], const SkinF16Dev& img,
                              const Work& w, int n_sms, cudaStream_t st, long long* launches);
cudaError_t root_init();
// bring-up probe: D[128][N] = A[128][K] . W[N][K]^T through the fp16 split-precision path
cudaError_t root_probe_f16(const float* A, const float* W, int K, int N, float* D, int mode, cudaStream_t st);

}  // namespace arah
