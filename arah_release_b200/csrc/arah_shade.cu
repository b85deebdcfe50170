// arah_shade.cu — k_shade16 (arah_shade16.cuh) in its own translation unit, with its weight packing and launcher.
#include "arah_shade.h"

#include "arah_shade16.cuh"

namespace arah {

static inline unsigned cdiv_s(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

size_t shade16_scratch_bytes_per_cta() { return (size_t)SH16_SCRATCH_FLOATS * 4; }

cudaError_t shade16_init() {
    static_assert(SHADE16_BWD_DEV_BYTES == SHADE16_BWD_BYTES && SHADE16_COL_DEV_BYTES == SHADE16_COL_BYTES, "image sizes");
    return cudaFuncSetAttribute(k_shade16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shade16_smem_bytes());
}

cudaError_t shade16_pack(const float* const sdf_W[7], const float* const col_W[6], int din, const Shade16Dev& dst, cudaStream_t st,
                         long long* launches) {
    auto up = [&](const float* src, int ld, __half* d, int N, int K, int nchunks, int split, int lo, int hi, int transpose) {
        k_pack_f16<<<cdiv_s((size_t)nchunks * N * HK, 256), 256, 0, st>>>(src, ld, d, N, K, nchunks, split, lo, hi, transpose);
        if (launches) *launches += 1;
    };
    __half* bwd = reinterpret_cast<__half*>(dst.bwd);
    for (int l = 1; l <= 5; ++l) up(sdf_W[l], 256, bwd + (size_t)(l - 1) * 65536, 256, 256, 4, 256, 0, 0, 1);
    // colour: our K order is [feat 256 | x, PE, n 33 | zero padding] (the reference's input is [x | PE | n | feat | latent])
    __half* c = reinterpret_cast<__half*>(dst.col);
    const int COLK = 256 + 33;
    up(col_W[0], din, c, 256, COLK, 5, 256, 33, 0, 0);              c += 5 * 16384;
    up(col_W[1], 256, c, 256, 256, 4, 256, 0, 0, 0);                c += 4 * 16384;
    up(col_W[2], 256, c, 128, 256, 4, 256, 0, 0, 0);                c += 4 * 8192;
    up(col_W[3], din + 128, c, 256, 128, 2, 128, din, 0, 0);        c += 2 * 16384;     // lin2-output part of the skip layer
    up(col_W[3], din + 128, c, 256, COLK, 5, 256, 33, 0, 0);        c += 5 * 16384;     // network-input part
    up(col_W[4], 256, c, 256, 256, 4, 256, 0, 0, 0);
    return cudaGetLastError();
}

cudaError_t shade16_launch(const FrameParams& fp, const Shade16Host& sh, const Shade16Dev& img, const Work& w, unsigned grid, cudaStream_t st,
                           long long* launches) {
    Shade16 tc;
    tc.sdf_Wt0 = sh.sdf_Wt0; tc.sdf_W0 = sh.sdf_W0; tc.sdf_F = sh.sdf_F; tc.sdf_G = sh.sdf_G; tc.sdf_scale = sh.sdf_scale;
    tc.sdf_fwd = reinterpret_cast<const __half*>(sh.sdf_fwd_hi);
    tc.sdf_bwd = reinterpret_cast<const __half*>(img.bwd);
    tc.sdf_w6 = sh.sdf_w6; tc.sdf_b6 = sh.sdf_b6;
    const __half* c = reinterpret_cast<const __half*>(img.col);
    tc.col0 = c;   c += 5 * 16384;
    tc.col1 = c;   c += 4 * 16384;
    tc.col2 = c;   c += 4 * 8192;
    tc.col3b = c;  c += 2 * 16384;
    tc.col3a = c;  c += 5 * 16384;
    tc.col4 = c;
    tc.col_W5 = sh.col_W5;
    for (int l = 0; l < 6; ++l) tc.col_b[l] = sh.col_b[l];
    k_shade16<<<grid, SH16_THREADS, shade16_smem_bytes(), st>>>(fp, tc, w);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace arah
