/*
 * oracle/mc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, 64-bit indexing) of the reference's dense-grid
 * marching cubes, /root/reference/src/prim3d/Utility/marching_cubes.cu.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path under
 * primitive3d_b200/ never does.
 *
 * What is restated (reference file:line):
 *   layout idx = i*(Ry*Rz) + j*Rz + k .......................... marching_cubes.cu:20
 *   inside predicate  d > thresh (NaN, ==thresh are outside) ... :25,31,37,43,50-57
 *   a vertex on every sign-changing +x/+y/+z edge owned by the
 *   lower voxel, for p_axis < R_axis-1 ......................... :29-45, :100-137
 *   dt = (thresh - d_self) / (d_next - d_self);  p_axis + dt .... :105-109,118-122,131-135
 *   cube case bits, corner order ............................... :168-176
 *   edge -> owning (voxel, axis) map ........................... :178-192
 *   triangles = table row up to the first -1, in table order ... :194-208
 *   world transform v*scale + offset, two roundings, with the
 *   reference's y-scale term (upper[2]-lower[1])/Ry [sic] ...... :290-298
 *
 * What is deliberately different: the reference hands out vertex and face
 * slots with atomicAdd (:104,117,130,199), so its output ORDER is
 * non-deterministic.  This restatement fixes the order a single thread
 * walking x, then y, then z would produce: vertex ids in (x, y, z, axis)
 * order, faces in (x, y, z) cell order.  Comparisons against the real
 * reference are therefore made on canonical (order-free) forms, see
 * tests/canonical.py.
 *
 * Parity status: pinned against the compiled reference (oracle/_ref) on the
 * GPU box by tests/test_reference_cuda.py, and against the survey's
 * known-answer counts in tests/test_oracle_mc.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: the reference is compiled without fast-math, so
 * every fp32 op below is a separately rounded IEEE operation.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "mc_case_rows.h"

static int8_t g_table[256][16];
static int g_ntri[256];
static int g_table_ready = 0;

static void build_table(void) {
    if (g_table_ready) return;
    for (int c = 0; c < 256; ++c) {
        const char *row = P3D_ORACLE_CASE_ROWS[c];
        int n = (int)strlen(row);
        for (int i = 0; i < 16; ++i) g_table[c][i] = -1;
        for (int i = 0; i < n; ++i) {
            char ch = row[i];
            g_table[c][i] = (int8_t)(ch <= '9' ? ch - '0' : ch - 'a' + 10);
        }
        g_ntri[c] = n / 3;
    }
    g_table_ready = 1;
}

/* Expanded int8[256][16] table, -1 terminated rows (marching_cubes.h:21-277). */
void p3d_oracle_mc_table(int8_t *out4096) {
    build_table();
    memcpy(out4096, g_table, sizeof g_table);
}

typedef struct {
    const float *d;
    int64_t rx, ry, rz;
    float thresh;
} grid_t;

static inline float at(const grid_t *g, int64_t i, int64_t j, int64_t k) {
    return g->d[i * (g->ry * g->rz) + j * g->rz + k]; /* marching_cubes.cu:20 */
}

static inline int cube_case(const grid_t *g, int64_t x, int64_t y, int64_t z) {
    const float t = g->thresh; /* marching_cubes.cu:168-176 */
    int m = 0;
    if (at(g, x, y, z) > t) m |= 1;
    if (at(g, x + 1, y, z) > t) m |= 2;
    if (at(g, x + 1, y + 1, z) > t) m |= 4;
    if (at(g, x, y + 1, z) > t) m |= 8;
    if (at(g, x, y, z + 1) > t) m |= 16;
    if (at(g, x + 1, y, z + 1) > t) m |= 32;
    if (at(g, x + 1, y + 1, z + 1) > t) m |= 64;
    if (at(g, x, y + 1, z + 1) > t) m |= 128;
    return m;
}

/* Count the vertices owned by plane x and the triangles of the cells of plane x
 * (count_vertices_faces_kernel, marching_cubes.cu:4-68). */
static void count_plane(const grid_t *g, int64_t x, int64_t *nv, int64_t *nt) {
    int64_t v = 0, t = 0;
    for (int64_t y = 0; y < g->ry; ++y)
        for (int64_t z = 0; z < g->rz; ++z) {
            const int inside = at(g, x, y, z) > g->thresh;
            if (x < g->rx - 1 && inside != (at(g, x + 1, y, z) > g->thresh)) ++v;
            if (y < g->ry - 1 && inside != (at(g, x, y + 1, z) > g->thresh)) ++v;
            if (z < g->rz - 1 && inside != (at(g, x, y, z + 1) > g->thresh)) ++v;
            if (x < g->rx - 1 && y < g->ry - 1 && z < g->rz - 1) t += g_ntri[cube_case(g, x, y, z)];
        }
    *nv = v;
    *nt = t;
}

int p3d_oracle_mc_count(const float *grid, int64_t rx, int64_t ry, int64_t rz, float thresh,
                        int64_t *num_vertices, int64_t *num_faces, int threads) {
    if (rx < 1 || ry < 1 || rz < 1) return 1;
    build_table();
    grid_t g = {grid, rx, ry, rz, thresh};
    int64_t V = 0, F = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : V, F)
    for (int64_t x = 0; x < rx; ++x) {
        int64_t nv, nt;
        count_plane(&g, x, &nv, &nt);
        V += nv;
        F += nt;
    }
    *num_vertices = V;
    *num_faces = F;
    return 0;
}

/* Assign ids to the vertices owned by plane x in (y, z, axis) order starting at
 * `base`; ids[(y*Rz+z)*3+axis] = id or -1.  When verts != NULL also writes the
 * interpolated positions (gen_vertices_kernel, marching_cubes.cu:70-138) followed
 * by the bounding-box transform (:290-298). */
static void plane_vertices(const grid_t *g, int64_t x, int64_t base, int32_t *ids, float *verts,
                           const float scale[3], const float offset[3]) {
    int64_t next = base;
    for (int64_t y = 0; y < g->ry; ++y)
        for (int64_t z = 0; z < g->rz; ++z) {
            const float self = at(g, x, y, z);
            const int inside = self > g->thresh;
            int32_t *slot = ids + (y * g->rz + z) * 3;
            for (int axis = 0; axis < 3; ++axis) {
                slot[axis] = -1;
                const int64_t p = axis == 0 ? x : (axis == 1 ? y : z);
                const int64_t r = axis == 0 ? g->rx : (axis == 1 ? g->ry : g->rz);
                if (p >= r - 1) continue;
                const float nb = at(g, x + (axis == 0), y + (axis == 1), z + (axis == 2));
                if (inside == (nb > g->thresh)) continue;
                slot[axis] = (int32_t)next;
                if (verts) {
                    const float dt = (g->thresh - self) / (nb - self);
                    float pos[3] = {(float)x, (float)y, (float)z};
                    pos[axis] = pos[axis] + dt;
                    for (int c = 0; c < 3; ++c) {
                        const float scaled = pos[c] * scale[c];
                        verts[next * 3 + c] = scaled + offset[c];
                    }
                }
                ++next;
            }
        }
}

/* Faces of the cells of plane x (gen_faces_kernel, marching_cubes.cu:140-209).
 * ids0 / ids1 are the vertex-id planes of x and x+1. */
static void plane_faces(const grid_t *g, int64_t x, const int32_t *ids0, const int32_t *ids1,
                        int32_t *faces, int64_t tri_base, int *missing) {
    int64_t t = tri_base;
    const int64_t rz = g->rz;
    for (int64_t y = 0; y < g->ry - 1; ++y)
        for (int64_t z = 0; z < rz - 1; ++z) {
            const int m = cube_case(g, x, y, z);
            if (g_ntri[m] == 0) continue;
#define ID(plane, yy, zz, ax) (plane)[((yy)*rz + (zz)) * 3 + (ax)]
            const int32_t e[12] = {
                ID(ids0, y, z, 0),     ID(ids1, y, z, 1),         ID(ids0, y + 1, z, 0),     ID(ids0, y, z, 1),
                ID(ids0, y, z + 1, 0), ID(ids1, y, z + 1, 1),     ID(ids0, y + 1, z + 1, 0), ID(ids0, y, z + 1, 1),
                ID(ids0, y, z, 2),     ID(ids1, y, z, 2),         ID(ids1, y + 1, z, 2),     ID(ids0, y + 1, z, 2)};
#undef ID
            for (int i = 0; i < 3 * g_ntri[m]; ++i) {
                const int32_t id = e[g_table[m][i]];
                if (id < 0) *missing = 1; /* the reference printf()s here, :204-206 */
                faces[t * 3 + i] = id;
            }
            t += g_ntri[m];
        }
}

/* Full extraction.  verts: float[V*3], faces: int32[F*3]; V and F must come from
 * p3d_oracle_mc_count.  Returns 0 ok, 1 bad args, 2 V does not fit int32,
 * 3 a table-referenced edge had no vertex (cannot happen for a consistent table). */
int p3d_oracle_mc_extract(const float *grid, int64_t rx, int64_t ry, int64_t rz, float thresh,
                          const float *lower, const float *upper, float *verts, int32_t *faces,
                          int threads) {
    if (rx < 1 || ry < 1 || rz < 1) return 1;
    build_table();
    grid_t g = {grid, rx, ry, rz, thresh};
    /* marching_cubes.cu:290-298 -- note upper[2] in the y term, as in the reference. */
    const float scale[3] = {(upper[0] - lower[0]) / (float)rx, (upper[2] - lower[1]) / (float)ry,
                            (upper[2] - lower[2]) / (float)rz};
    const float offset[3] = {lower[0], lower[1], lower[2]};

#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    int64_t *vbase = (int64_t *)malloc((size_t)(rx + 1) * sizeof(int64_t));
    int64_t *tbase = (int64_t *)malloc((size_t)(rx + 1) * sizeof(int64_t));
    if (!vbase || !tbase) return 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t x = 0; x < rx; ++x) count_plane(&g, x, &vbase[x + 1], &tbase[x + 1]);
    vbase[0] = tbase[0] = 0;
    for (int64_t x = 0; x < rx; ++x) {
        vbase[x + 1] += vbase[x];
        tbase[x + 1] += tbase[x];
    }
    if (vbase[rx] > INT32_MAX) {
        free(vbase);
        free(tbase);
        return 2;
    }

    int missing = 0;
#pragma omp parallel reduction(| : missing)
    {
        /* each thread walks a contiguous block of planes with two rolling id planes, so the
         * ids of plane x+1 computed for the cells of plane x are reused as plane x+1's own */
        int nt = 1, tid = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        tid = omp_get_thread_num();
#endif
        const int64_t xa = rx * tid / nt, xb = rx * (tid + 1) / nt;
        const size_t plane = (size_t)ry * (size_t)rz * 3;
        int32_t *ids0 = (int32_t *)malloc(plane * sizeof(int32_t));
        int32_t *ids1 = (int32_t *)malloc(plane * sizeof(int32_t));
        for (int64_t x = xa; x < xb; ++x) {
            if (x == xa) plane_vertices(&g, x, vbase[x], ids0, verts, scale, offset);
            if (x + 1 < rx) {
                /* positions of plane x+1 are written here only if this thread owns that plane */
                plane_vertices(&g, x + 1, vbase[x + 1], ids1, x + 1 < xb ? verts : NULL, scale, offset);
                plane_faces(&g, x, ids0, ids1, faces, tbase[x], &missing);
            }
            int32_t *t = ids0;
            ids0 = ids1;
            ids1 = t;
        }
        free(ids0);
        free(ids1);
    }
    free(vbase);
    free(tbase);
    return missing ? 3 : 0;
}
