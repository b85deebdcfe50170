// arah_root.cu — the persistent root-finding kernels (fp16 split-precision tensor-core engine) and their host launchers.
#include "arah_root.h"

#include "arah_corr_p.cuh"
#include "arah_iso_p.cuh"
#include "arah_trace_p.cuh"
#include "arah_sdf_fwd16.cuh"
#include "arah_grid16.cuh"

namespace arah {

static inline unsigned cdiv_u(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

cudaError_t root_init() {
    cudaError_t e = cudaFuncSetAttribute(k_corr_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)corr_persist_smem_bytes());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sdf_fwd16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sdf_fwd16_smem_bytes());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_sdf_grid16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sdf_grid16_smem_bytes());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_iso_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)iso_persist_smem_bytes());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_umma_f16_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 2 * 256 * HK * 2 + 1024 + 64);
}

constexpr size_t MAX_DYN_SMEM = 232448;        // 227 KB per CTA on sm_100
bool root_trace_fits(int n_verts) { return trace_persist_smem_bytes(n_verts) <= MAX_DYN_SMEM; }

cudaError_t root_pack_sdf_f16(const float* const W[7], const SdfF16Dev& dst, cudaStream_t st, long long* launches) {
    __half* hi = reinterpret_cast<__half*>(dst.hi);
    __half* lo = reinterpret_cast<__half*>(dst.lo);
    ScaleJobs jobs{};
    for (int l = 1; l <= 5; ++l) { jobs.W[l - 1] = W[l]; jobs.n[l - 1] = 256 * 256; jobs.out[l - 1] = dst.scale + 2 * (l - 1); }
    k_layer_scales<<<5, 1024, 0, st>>>(jobs);
    if (launches) *launches += 1;
    for (int l = 1; l <= 5; ++l) {
        const size_t off = (size_t)(l - 1) * 131072 / 2;            // halfs
        k_pack_f16x2<<<cdiv_u((size_t)4 * 256 * HK, 256), 256, 0, st>>>(W[l], 256, dst.scale + 2 * (l - 1), hi + off, lo + off, 256, 256, 256, 4);
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}

static SdfF16 make_sdf16(const SdfF16Host& sh, const SdfF16Dev& img) {
    SdfF16 sd;
    sd.Wt0 = sh.Wt0; sd.freq = sh.freq; sd.phase = sh.phase; sd.w6 = sh.w6; sd.b6 = sh.b6;
    for (int l = 0; l < 6; ++l) sd.b[l] = sh.b[l];
    sd.hi = reinterpret_cast<const __half*>(img.hi); sd.lo = reinterpret_cast<const __half*>(img.lo); sd.scale = img.scale;
    return sd;
}

cudaError_t root_trace_persist(const FrameParams& fp, const SdfF16Host& sh, const SdfF16Dev& img, const KnnIndex& ix, const Work& w, int n_sms,
                               cudaStream_t st, long long* launches) {
    static int attr_for = -1;
    const size_t smem = trace_persist_smem_bytes(fp.n_verts);
    if (attr_for != (int)smem) {
        cudaError_t e = cudaFuncSetAttribute(k_trace_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_for = (int)smem;
    }
    const size_t tiles = ((size_t)w.P + UM - 1) / UM;
    const unsigned g = (unsigned)(tiles < (size_t)n_sms ? tiles : (size_t)n_sms);
    k_trace_persist<<<g, S16_THREADS, smem, st>>>(fp, make_sdf16(sh, img), ix, w);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t root_sdf_fwd16(const FrameParams& fp, const SdfF16Host& sh, const SdfF16Dev& img, const Work& w, int n_sms, cudaStream_t st,
                           long long* launches) {
    const size_t tiles = ((size_t)w.P * w.S + UM - 1) / UM;
    const unsigned g = (unsigned)(tiles < (size_t)n_sms ? tiles : (size_t)n_sms);
    k_sdf_fwd16<<<g, F16_THREADS, sdf_fwd16_smem_bytes(), st>>>(fp, make_sdf16(sh, img), w, Fwd16Grid{0, 0.f, 0, nullptr});
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t root_sdf_grid16(const SdfF16Host& sh, const SdfF16Dev& img, int N, float voxel, long long n, float* out, int n_sms, cudaStream_t st) {
    const size_t tiles = (size_t)((n + UM - 1) / UM);
    const unsigned g = (unsigned)(tiles < (size_t)n_sms ? tiles : (size_t)n_sms);
    k_sdf_grid16<<<g, S16_THREADS, sdf_grid16_smem_bytes(), st>>>(make_sdf16(sh, img), N, voxel, n, out, nullptr, nullptr, 0.f, nullptr);
    return cudaGetLastError();
}

cudaError_t root_sdf_grid_banded(const SdfF16Host& sh, const SdfF16Dev& img, int N, float voxel, float level, float eps, float* out,
                                 uint8_t* flag, int* list, int* stats, int n_sms, cudaStream_t st, long long* launches) {
    const int n = N * N * N;                                       // N <= 1024 (checked by the caller): fits an int
    const size_t tiles = ((size_t)n + UM - 1) / UM;
    const unsigned g = (unsigned)(tiles < (size_t)n_sms ? tiles : (size_t)n_sms);
    cudaError_t e = cudaMemsetAsync(flag, 0, (size_t)n, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(stats, 0, 2 * sizeof(int), st);
    if (e != cudaSuccess) return e;
    // 1. every lattice point in one fp16 pass
    FrameParams fp0{};
    Work w0{};
    k_sdf_fwd16<<<g, F16_THREADS, sdf_fwd16_smem_bytes(), st>>>(fp0, make_sdf16(sh, img), w0, Fwd16Grid{N, voxel, n, out});
    // 2. corners of the cells that may straddle the level, as a list
    k_grid_band_flag<<<dim3(cdiv_u((size_t)N - 1, 128), N - 1, N - 1), 128, 0, st>>>(out, N, level, eps, flag);
    k_grid_band_list<<<4 * n_sms, 256, 0, st>>>(flag, n, list, stats);
    // 3. those points again in split precision (grid: the list length is only known on the device; idle CTAs exit at once)
    k_sdf_grid16<<<g, S16_THREADS, sdf_grid16_smem_bytes(), st>>>(make_sdf16(sh, img), N, voxel, 0, out, list, stats, eps, stats + 1);
    if (launches) *launches += 4;
    return cudaGetLastError();
}

cudaError_t root_iso_persist(const FrameParams& fp, const SdfF16Host& sh, const SdfF16Dev& img, const float* skin_Wt0, const float* const skin_b[5],
                             const SkinF16Dev& skimg, const Work& w, int n_sms, cudaStream_t st, long long* launches) {
    SkinF16 sk;
    sk.Wt0 = skin_Wt0;
    for (int l = 0; l < 5; ++l) sk.b[l] = skin_b[l];
    sk.hi = reinterpret_cast<const __half*>(skimg.hi); sk.lo = reinterpret_cast<const __half*>(skimg.lo); sk.scale = skimg.scale;
    const size_t tiles = ((size_t)w.P + UM - 1) / UM;
    const unsigned g = (unsigned)(tiles < (size_t)n_sms ? tiles : (size_t)n_sms);
    k_iso_persist<<<g, S16_THREADS, iso_persist_smem_bytes(), st>>>(fp, make_sdf16(sh, img), sk, w);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t root_pack_skin_f16(const float* const W[5], const SkinF16Dev& dst, cudaStream_t st, long long* launches) {
    __half* hi = reinterpret_cast<__half*>(dst.hi);
    __half* lo = reinterpret_cast<__half*>(dst.lo);
    ScaleJobs jobs{};
    for (int l = 1; l <= 4; ++l) { jobs.W[l - 1] = W[l]; jobs.n[l - 1] = ((l < 4) ? 128 : 25) * 128; jobs.out[l - 1] = dst.scale + 2 * (l - 1); }
    k_layer_scales<<<4, 1024, 0, st>>>(jobs);
    if (launches) *launches += 1;
    for (int l = 1; l <= 4; ++l) {
        const int N = (l < 4) ? 128 : 25, Npad = (l < 4) ? 128 : 32;
        const size_t off = (size_t)(l - 1) * 32768 / 2;             // halfs
        k_pack_f16x2<<<cdiv_u((size_t)2 * Npad * HK, 256), 256, 0, st>>>(W[l], 128, dst.scale + 2 * (l - 1), hi + off, lo + off, N, Npad, 128, 2);
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}

cudaError_t root_corr_persist(const FrameParams& fp, const float* skin_Wt0, const float* const skin_b[5], const SkinF16Dev& img,
                              const Work& w, int n_sms, cudaStream_t st, long long* launches) {
    SkinF16 sk;
    sk.Wt0 = skin_Wt0;
    for (int l = 0; l < 5; ++l) sk.b[l] = skin_b[l];
    sk.hi = reinterpret_cast<const __half*>(img.hi);
    sk.lo = reinterpret_cast<const __half*>(img.lo);
    sk.scale = img.scale;
    const size_t PS = (size_t)w.P * w.S;
    const unsigned g = (unsigned)((PS + 2 * UM - 1) / (2 * UM) < (size_t)n_sms ? (PS + 2 * UM - 1) / (2 * UM) : (size_t)n_sms);
    k_corr_persist<<<g, CP_THREADS, corr_persist_smem_bytes(), st>>>(fp, sk, w);
    const unsigned gc = (unsigned)((PS + 1023) / 1024 < (size_t)(4 * n_sms) ? (PS + 1023) / 1024 : (size_t)(4 * n_sms));
    k_shade_compact<<<gc, 1024, 0, st>>>(w);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

cudaError_t root_probe_f16(const float* A, const float* W, int K, int N, float* D, int mode, cudaStream_t st) {
    const int nch = K / HK;
    __half* img = nullptr;
    float* sc = nullptr;
    cudaError_t e = cudaMalloc(&img, (size_t)2 * nch * N * HK * 2);
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&sc, 8);
    if (e != cudaSuccess) { cudaFree(img); return e; }
    __half* hi = img;
    __half* lo = img + (size_t)nch * N * HK;
    k_layer_scale<<<1, 1024, 0, st>>>(W, N * K, sc);
    k_pack_f16x2<<<cdiv_u((size_t)nch * N * HK, 256), 256, 0, st>>>(W, K, sc, hi, lo, N, N, K, nch);
    const int smem = 2 * nch * N * HK * 2 + 1024 + 64;
    k_umma_f16_probe<<<1, 128, smem, st>>>(A, hi, lo, sc, K, N, D, mode);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(img);
    cudaFree(sc);
    return e;
}

}  // namespace arah
