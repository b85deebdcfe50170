/* TEST INFRASTRUCTURE — CPU restatement of the canonical-mesh extraction.
 *
 * What it follows: /root/reference/im2mesh/utils/sdf_meshing.py:13-114
 *   :20-38   lattice coordinates  (index * voxel_size + origin, fp32, one rounding per operation)
 *   :93-98   skimage.measure.marching_cubes_lewiner(sdf, level=0.0, spacing=[voxel_size]*3)
 *   :100-105 mesh_points = voxel_grid_origin + verts
 *
 * PARITY UNPINNED for the triangulation: skimage 0.18.1 (environment.yml of the reference) is a third-party dependency
 * that is absent from /root/reference and from this image, and the reference holds no mesh fixtures.  This file therefore
 * restates the *published* marching-cubes scheme the GPU kernels implement (arah_release_b200/csrc/arah_mesh.cu):
 *   - a cube corner c sits at (c & 1, (c >> 1) & 1, (c >> 2) & 1); it is "inside" when value < level;
 *   - one vertex per sign-changing lattice edge at t = v0 / (v0 - v1) (linear interpolation); skimage's Lewiner code
 *     weights the two end points by 1 / (eps + |v|), which is the same point up to fp32 rounding;
 *   - on every cube face the crossings are joined by segments; an ambiguous face (inside corners on a diagonal) always
 *     separates the two inside corners, so neighbouring cubes agree and the surface is closed;
 *   - the directed segments of the six faces chain into closed polygons that are fan-triangulated from the first vertex
 *     whose diagonals do not lie in a cube face; triangles wind so that the normal points towards larger values;
 *   - output order: lattice order (x slowest, z fastest), per lattice point its +x, +y, +z edge vertices, per cube the
 *     triangles in polygon order.
 * What IS checked against this file bit for bit: the CUDA kernels' vertices and faces; what is checked independently of
 * both: closedness (every edge shared by exactly two triangles with opposite directions), Euler characteristic and
 * enclosed volume of analytic bodies (tests/test_mesh_oracle.py).
 */
#pragma GCC optimize("fp-contract=off")
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void corners_of_edge(int e, int *c0, int *c1) {
    const int axis = e / 4, u = e % 2, v = (e / 2) % 2;
    int bit_u, bit_v;
    if (axis == 0) { bit_u = 1; bit_v = 2; } else if (axis == 1) { bit_u = 0; bit_v = 2; } else { bit_u = 0; bit_v = 1; }
    *c0 = (u << bit_u) | (v << bit_v);
    *c1 = *c0 | (1 << axis);
}
static int edge_of_corners(int a, int b) {
    for (int e = 0; e < 12; ++e) {
        int c0, c1;
        corners_of_edge(e, &c0, &c1);
        if ((c0 == a && c1 == b) || (c0 == b && c1 == a)) return e;
    }
    return -1;
}
/* the two faces (2 * axis + side) an edge lies on */
static void faces_of_edge(int e, int *f0, int *f1) {
    const int axis = e / 4, u = e % 2, v = (e / 2) % 2;
    const int au = (axis == 0) ? 1 : 0, av = (axis == 2) ? 1 : 2;
    *f0 = 2 * au + u;
    *f1 = 2 * av + v;
}
static int same_face(int e1, int e2) {
    int a0, a1, b0, b1;
    faces_of_edge(e1, &a0, &a1);
    faces_of_edge(e2, &b0, &b1);
    return a0 == b0 || a0 == b1 || a1 == b0 || a1 == b1;
}

typedef struct { int n; int t[16]; } Case;
static Case g_case[256];
static int g_ready = 0;

/* directed segment A -> B on the face with outward normal n: the inside corner `cin` lies on the side of -(n x (B - A)) */
static void link(int *next, int eA, int eB, int cin, int axis, int side) {
    double pa[3], pb[3], pc[3], n[3] = {0, 0, 0};
    int c0, c1;
    corners_of_edge(eA, &c0, &c1);
    for (int k = 0; k < 3; ++k) pa[k] = 0.5 * (((c0 >> k) & 1) + ((c1 >> k) & 1));
    corners_of_edge(eB, &c0, &c1);
    for (int k = 0; k < 3; ++k) pb[k] = 0.5 * (((c0 >> k) & 1) + ((c1 >> k) & 1));
    for (int k = 0; k < 3; ++k) pc[k] = (cin >> k) & 1;
    n[axis] = side ? 1.0 : -1.0;
    const double d[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
    const double s[3] = {n[1] * d[2] - n[2] * d[1], n[2] * d[0] - n[0] * d[2], n[0] * d[1] - n[1] * d[0]};
    double dot = 0;
    for (int k = 0; k < 3; ++k) dot += s[k] * (pc[k] - 0.5 * (pa[k] + pb[k]));
    if (dot < 0) next[eA] = eB; else next[eB] = eA;
}

static void make_cases(void) {
    for (int cs = 0; cs < 256; ++cs) {
        int next[12];
        for (int e = 0; e < 12; ++e) next[e] = -1;
        for (int axis = 0; axis < 3; ++axis)
            for (int side = 0; side < 2; ++side) {
                const int bu = (axis == 0) ? 1 : 0, bv = (axis == 2) ? 1 : 2;
                const int ring_u[4] = {0, 1, 1, 0}, ring_v[4] = {0, 0, 1, 1};
                int c[4], in[4], nin = 0;
                for (int k = 0; k < 4; ++k) {
                    c[k] = (side << axis) | (ring_u[k] << bu) | (ring_v[k] << bv);
                    in[k] = (cs >> c[k]) & 1;
                    nin += in[k];
                }
                if (nin == 0 || nin == 4) continue;
                if (nin == 1 || (nin == 2 && in[0] == in[2])) {            /* every inside corner is cut off alone */
                    for (int k = 0; k < 4; ++k)
                        if (in[k]) link(next, edge_of_corners(c[k], c[(k + 3) % 4]), edge_of_corners(c[k], c[(k + 1) % 4]), c[k], axis, side);
                } else if (nin == 3) {                                     /* the single outside corner is cut off */
                    for (int k = 0; k < 4; ++k)
                        if (!in[k]) link(next, edge_of_corners(c[k], c[(k + 3) % 4]), edge_of_corners(c[k], c[(k + 1) % 4]), c[(k + 2) % 4], axis, side);
                } else {                                                   /* two neighbouring inside corners */
                    for (int k = 0; k < 4; ++k)
                        if (in[k] && in[(k + 1) % 4])
                            link(next, edge_of_corners(c[k], c[(k + 3) % 4]), edge_of_corners(c[(k + 1) % 4], c[(k + 2) % 4]), c[k], axis, side);
                }
            }
        Case *C = &g_case[cs];
        C->n = 0;
        int seen[12] = {0};
        for (int start = 0; start < 12; ++start) {
            if (next[start] < 0 || seen[start]) continue;
            int poly[12], m = 0;
            for (int e = start; e >= 0 && !seen[e]; e = next[e]) { seen[e] = 1; poly[m++] = e; }
            int apex = 0;
            for (int s = 0; s < m; ++s) {
                int ok = 1;
                for (int k = 2; k + 1 < m; ++k) if (same_face(poly[s], poly[(s + k) % m])) ok = 0;
                if (ok) { apex = s; break; }
            }
            for (int k = 1; k + 1 < m; ++k) {
                if (C->n + 3 > 15) { C->n = -1; break; }
                C->t[C->n++] = poly[apex]; C->t[C->n++] = poly[(apex + k) % m]; C->t[C->n++] = poly[(apex + k + 1) % m];
            }
            if (C->n < 0) break;
        }
    }
    g_ready = 1;
}

/* tri[256][16] (-1 terminated) / ntri[256], for the table-equality test */
int arah_oracle_mc_table(int8_t *tri, uint8_t *ntri) {
    if (!g_ready) make_cases();
    for (int cs = 0; cs < 256; ++cs) {
        if (g_case[cs].n < 0) return -1;
        ntri[cs] = (uint8_t)(g_case[cs].n / 3);
        for (int k = 0; k < 16; ++k) tri[cs * 16 + k] = (k < g_case[cs].n) ? (int8_t)g_case[cs].t[k] : (int8_t)-1;
    }
    return 0;
}

/* sdf [N][N][N] -> verts [.][3], faces [.][3]; counts[0..1] = totals (outputs are truncated at max_*). */
int arah_oracle_mc(const float *sdf, int N, float level, float voxel, const float *origin, float *verts, int max_verts,
                   int32_t *faces, int max_faces, int32_t *counts) {
    if (!g_ready) make_cases();
    const size_t n = (size_t)N * N * N, sx = (size_t)N * N, sy = (size_t)N;
    int32_t *first = (int32_t *)malloc(n * sizeof(int32_t));      /* id of the first vertex owned by a lattice point */
    uint8_t *own = (uint8_t *)malloc(n);
    if (!first || !own) { free(first); free(own); return -2; }
    int nv = 0;
    for (size_t p = 0; p < n; ++p) {
        const int iz = (int)(p % N), iy = (int)((p / N) % N), ix = (int)(p / sx);
        const int idx[3] = {ix, iy, iz};
        const size_t st[3] = {sx, sy, 1};
        const int in0 = sdf[p] < level;
        first[p] = nv;
        own[p] = 0;
        for (int a = 0; a < 3; ++a) {
            if (idx[a] + 1 >= N) continue;
            const int in1 = sdf[p + st[a]] < level;
            if (in1 == in0) continue;
            own[p] |= (uint8_t)(1 << a);
            if (nv < max_verts) {
                const float v0 = sdf[p] - level, v1 = sdf[p + st[a]] - level;
                const float t = v0 / (v0 - v1);
                for (int k = 0; k < 3; ++k) {
                    float g = (float)idx[k];
                    if (k == a) g = g + t;
                    const float m = g * voxel;
                    verts[3 * (size_t)nv + k] = m + origin[k];
                }
            }
            ++nv;
        }
    }
    int nf = 0;
    for (size_t p = 0; p < n; ++p) {
        const int iz = (int)(p % N), iy = (int)((p / N) % N), ix = (int)(p / sx);
        if (ix + 1 >= N || iy + 1 >= N || iz + 1 >= N) continue;
        int cs = 0;
        for (int c = 0; c < 8; ++c)
            if (sdf[p + (c & 1) * sx + ((c >> 1) & 1) * sy + ((c >> 2) & 1)] < level) cs |= 1 << c;
        const Case *C = &g_case[cs];
        for (int k = 0; k < C->n; k += 3) {
            if (nf < max_faces)
                for (int j = 0; j < 3; ++j) {
                    const int e = C->t[k + j], a = e / 4, u = e % 2, v = (e / 2) % 2;
                    const size_t su = (a == 0) ? sy : sx, sv = (a == 2) ? sy : 1;
                    const size_t q = p + u * su + v * sv;
                    int rank = 0;
                    for (int b = 0; b < a; ++b) rank += (own[q] >> b) & 1;
                    faces[3 * (size_t)nf + j] = first[q] + rank;
                }
            ++nf;
        }
    }
    counts[0] = nv;
    counts[1] = nf;
    free(first);
    free(own);
    return 0;
}
