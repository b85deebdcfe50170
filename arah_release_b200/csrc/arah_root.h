// arah_root.h — host-side entry points of arah_root.cu (the persistent root-finding kernels and their weight packing) for
// arah_api.cu.  Plain declarations: the kernels themselves live in arah_root.cu so that the two translation units compile in
// parallel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "arah_work.cuh"

namespace arah {

// device images of the skinning MLP for k_corr_persist (arah_corr_p.cuh)
struct SkinF16Dev {
    void* hi;        // 3 x 32 KB + 8 KB
    void* lo;        // same
    float* scale;    // [4][2]
};
constexpr size_t SKIN_F16_IMAGE_BYTES = 3 * 32768 + 8192;

// device images of the SDF hidden layers for the fp16 split-precision engine (arah_sdf16.cuh)
struct SdfF16Dev {
    void* hi;        // 5 x 128 KB
    void* lo;
    float* scale;    // [5][2]
};
constexpr size_t SDF_F16_DEV_BYTES = 5 * 131072;
struct SdfF16Host {  // what the kernels need besides the images (device pointers into the frame arena)
    const float* Wt0; const float* b[6]; const float* freq; const float* phase; const float* w6; const float* b6;
};
cudaError_t root_pack_sdf_f16(const float* const W[7], const SdfF16Dev& dst, cudaStream_t st, long long* launches);
// sphere tracing of the rays listed in w.listA (k_trace_begin), persistent; false if the vertex index does not fit in shared memory
bool root_trace_fits(int n_verts);
cudaError_t root_trace_persist(const FrameParams& fp, const SdfF16Host& sh, const SdfF16Dev& img, const KnnIndex& ix, const Work& w, int n_sms,
                               cudaStream_t st, long long* launches);
// SDF value (metres) of every sample of w.shade_list -> w.smp_sdf, single-pass fp16 tensor-core tiles (arah_sdf_fwd16.cuh)
cudaError_t root_sdf_fwd16(const FrameParams& fp, const SdfF16Host& sh, const SdfF16Dev& img, const Work& w, int n_sms, cudaStream_t st,
                           long long* launches);
// canonical SDF lattice (row f1): raw network output at the N^3 lattice points of [-1, 1]^3
cudaError_t root_sdf_grid16(const SdfF16Host& sh, const SdfF16Dev& img, int N, float voxel, long long n, float* out, int n_sms, cudaStream_t st);
// the lattice marching cubes needs (arah_sdf_grid_banded): one fp16 pass over all points, then split precision for the corners of
// every cell whose coarse values lie within eps of straddling `level`.  flag: N^3 bytes, list: N^3 ints, stats: 2 ints
// (refined points, refined points whose coarse value was off by more than eps) — all device scratch of the caller.
cudaError_t root_sdf_grid_banded(const SdfF16Host& sh, const SdfF16Dev& img, int N, float voxel, float level, float eps, float* out,
                                 uint8_t* flag, int* list, int* stats, int n_sms, cudaStream_t st, long long* launches);
// joint search of the rays listed in w.listA (k_iso_prepare) whose state k_iso_init_tc3 has written, persistent
cudaError_t root_iso_persist(const FrameParams& fp, const SdfF16Host& sh, const SdfF16Dev& img, const float* skin_Wt0, const float* const skin_b[5],
                             const SkinF16Dev& skimg, const Work& w, int n_sms, cudaStream_t st, long long* launches);
// pack layers 1..4 of the skinning MLP (reference layout [out][in], fp32) into scaled fp16 hi / lo chunk images
cudaError_t root_pack_skin_f16(const float* const W[5], const SkinF16Dev& dst, cudaStream_t st, long long* launches);
// correspondence search of all on-samples (seeds in w.corr_seed) + the list of converged samples
cudaError_t root_corr_persist(const FrameParams& fp, const float* skin_Wt0, const float* const skin_b[5], const SkinF16Dev& img,
                              const Work& w, int n_sms, cudaStream_t st, long long* launches);
cudaError_t root_init();
// bring-up probe: D[128][N] = A[128][K] . W[N][K]^T through the fp16 split-precision path
cudaError_t root_probe_f16(const float* A, const float* W, int K, int N, float* D, int mode, cudaStream_t st);

}  // namespace arah
