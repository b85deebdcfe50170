// TEST INFRASTRUCTURE — host instantiation of csrc/arah_image_core.h (the per-element arithmetic of arah_image.cu).
//
// Each function below is the serial counterpart of one kernel of arah_image.cu: same core calls, same index expressions, a for
// loop instead of a grid, a plain minimum instead of atomicMin.  tests/test_images_host.py compares it with the numpy oracle and
// the reference's golden images in the build container (no GPU there).  Never loaded by the product.
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -o libarah_image_host.so host_image.cpp
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../arah_release_b200/csrc/arah_image_core.h"

using namespace arah_img;

extern "C" {

// k_img_scatter + k_img_normals
void host_frame_images(const float* rgb, const float* pts, const int32_t* pix, int P, int H, int W, float* pred_pixels, float* pred_normals) {
    const size_t n = (size_t)H * W;
    std::vector<float> img_pts(n * 3, 0.0f);
    memset(pred_pixels, 0, n * 3 * sizeof(float));
    for (int k = 0; k < P; ++k) {
        const int p = pix[k];
        if (p < 0 || p >= (int)n) continue;
        for (int c = 0; c < 3; ++c) { pred_pixels[(size_t)p * 3 + c] = rgb[(size_t)k * 3 + c]; img_pts[(size_t)p * 3 + c] = pts[(size_t)k * 3 + c]; }
    }
    for (size_t i = 0; i < n; ++i) depth_normal(img_pts.data(), H, W, (int)(i / W), (int)(i % W), pred_normals + i * 3);
}

// k_project
void host_project(const float* verts, int n, const float* cam16, float* ndc) {
    Camera c;
    memcpy(c.R, cam16, 9 * sizeof(float)); memcpy(c.T, cam16 + 9, 3 * sizeof(float));
    c.fx = cam16[12]; c.fy = cam16[13]; c.px = cam16[14]; c.py = cam16[15];
    for (int v = 0; v < n; ++v) project(c, verts + (size_t)v * 3, ndc + (size_t)v * 3);
}

// k_raster_faces + k_raster_resolve
void host_rasterize(const float* ndc, const int32_t* faces, int n_faces, int n_verts, int H, int W, int32_t* pix_to_face, float* zbuf) {
    const size_t n = (size_t)H * W;
    std::vector<unsigned long long> keys(n, KEY_EMPTY);
    for (int f = 0; f < n_faces; ++f) {
        const int i0 = faces[(size_t)f * 3], i1 = faces[(size_t)f * 3 + 1], i2 = faces[(size_t)f * 3 + 2];
        if ((unsigned)i0 >= (unsigned)n_verts || (unsigned)i1 >= (unsigned)n_verts || (unsigned)i2 >= (unsigned)n_verts) continue;
        const FaceSetup s = face_setup(ndc + (size_t)i0 * 3, ndc + (size_t)i1 * 3, ndc + (size_t)i2 * 3);
        if (!s.drawable) continue;
        int x_lo, x_hi, y_lo, y_hi;
        pixel_range(s.xmin, s.xmax, W, H, &x_lo, &x_hi);
        pixel_range(s.ymin, s.ymax, H, W, &y_lo, &y_hi);
        if (x_lo > x_hi || y_lo > y_hi) continue;
        const int bw = x_hi - x_lo + 1;
        const long long npx = (long long)bw * (y_hi - y_lo + 1);
        if (npx > 256) {                                 // the kernel's warp-cooperative shape: a flat index over the pixel box
            for (long long i = 0; i < npx; ++i) {
                const int y = y_lo + (int)(i / bw), x = x_lo + (int)(i % bw);
                float pz;
                if (face_covers(s, pix_to_ndc(W - 1 - x, W, H), pix_to_ndc(H - 1 - y, H, W), &pz)) {
                    const unsigned long long k = raster_key(pz, f);
                    if (k < keys[(size_t)y * W + x]) keys[(size_t)y * W + x] = k;
                }
            }
            continue;
        }
        for (int y = y_lo; y <= y_hi; ++y) {
            const float py = pix_to_ndc(H - 1 - y, H, W);
            if (py < s.ymin || py > s.ymax) continue;
            for (int x = x_lo; x <= x_hi; ++x) {
                float pz;
                if (face_covers(s, pix_to_ndc(W - 1 - x, W, H), py, &pz)) {
                    const unsigned long long k = raster_key(pz, f);
                    if (k < keys[(size_t)y * W + x]) keys[(size_t)y * W + x] = k;
                }
            }
        }
    }
    for (size_t i = 0; i < n; ++i) {
        const bool bg = keys[i] == KEY_EMPTY;
        pix_to_face[i] = bg ? -1 : (int32_t)(unsigned)(keys[i] & 0xffffffffull);
        union { uint32_t u; float f; } c; c.u = (uint32_t)(keys[i] >> 32);
        zbuf[i] = bg ? -1.0f : c.f;
    }
}

// k_normal_image
void host_normal_image(const float* verts, int n_verts, const int32_t* faces, int n_faces, const int32_t* pix_to_face, int H, int W, float sign,
                       const float* rot, float background, float* image) {
    const size_t n = (size_t)H * W;
    for (size_t i = 0; i < n; ++i) {
        const int f = pix_to_face[i];
        float* o = image + i * 3;
        o[0] = o[1] = o[2] = to_unit(background);
        if (f >= 0 && f < n_faces) {
            const int i0 = faces[(size_t)f * 3], i1 = faces[(size_t)f * 3 + 1], i2 = faces[(size_t)f * 3 + 2];
            if ((unsigned)i0 < (unsigned)n_verts && (unsigned)i1 < (unsigned)n_verts && (unsigned)i2 < (unsigned)n_verts)
                face_normal_pixel(verts + (size_t)i0 * 3, verts + (size_t)i1 * 3, verts + (size_t)i2 * 3, sign, rot, o);
        }
    }
}

}  // extern "C"
