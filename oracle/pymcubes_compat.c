/*
 * oracle/pymcubes_compat.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A PyMCubes-compatible CPU marching cubes, written because the reference's CPU
 * path (prim3d/utility/marching_cubes.py:66-81) and its examples
 * (examples/sphere.py:2,23-30, examples/bunny_sdf.py:4,24-31) call the third-party
 * package `mcubes` (PyMCubes, un-pinned: the reference has no requirements file),
 * which is not installed in this image and cannot be installed offline.
 *
 * PARITY UNPINNED: PyMCubes' source is not under /root/reference, so this is a
 * restatement of its published algorithm (mcubes/src/marchingcubes.h) from the
 * algorithm description, not a verified port:
 *   - cells visited x outer, y, z inner; corner numbering as in the reference's
 *     marching_cubes.cu:50-57;
 *   - case bit m set iff v[m] <= isovalue (the opposite polarity to the
 *     reference's CUDA path, which sets it iff v > thresh);
 *   - one shared vertex per sign-changing grid edge, interpolated in double:
 *       x1 + (x2 - x1) * (iso - f1) / (f2 - f1), midpoint when f1 == f2;
 *   - coordinates in index space; triangles follow the Bourke table row of the
 *     case in table order.
 * The only contract the reference pins at this boundary is equality of vertex and
 * face COUNTS with its CUDA path (examples/sphere.py:27-28, bunny_sdf.py:28-29).
 * It is single-threaded, like PyMCubes.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mc_case_rows.h"

typedef struct {
    double *v;
    size_t n, cap;
} dvec;
typedef struct {
    uint64_t *v;
    size_t n, cap;
} uvec;

static int dpush3(dvec *a, double x, double y, double z) {
    if (a->n + 3 > a->cap) {
        size_t cap = a->cap ? a->cap * 2 : 3 << 12;
        double *p = (double *)realloc(a->v, cap * sizeof(double));
        if (!p) return 1;
        a->v = p;
        a->cap = cap;
    }
    a->v[a->n++] = x;
    a->v[a->n++] = y;
    a->v[a->n++] = z;
    return 0;
}
static int upush(uvec *a, uint64_t x) {
    if (a->n + 1 > a->cap) {
        size_t cap = a->cap ? a->cap * 2 : 3 << 12;
        uint64_t *p = (uint64_t *)realloc(a->v, cap * sizeof(uint64_t));
        if (!p) return 1;
        a->v = p;
        a->cap = cap;
    }
    a->v[a->n++] = x;
    return 0;
}

static double interp(double iso, double f1, double f2, double x1, double x2) {
    if (f2 == f1) return (x2 + x1) / 2;
    return (x2 - x1) * (iso - f1) / (f2 - f1) + x1;
}

/* corner m -> (dx,dy,dz), and edge e -> (corner a, corner b) with a the owner (lower) corner */
static const int CORNER[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const int EDGE[12][3] = {/* owner corner, other corner, axis */
                                {0, 1, 0}, {1, 2, 1}, {3, 2, 0}, {0, 3, 1}, {4, 5, 0}, {5, 6, 1},
                                {7, 6, 0}, {4, 7, 1}, {0, 4, 2}, {1, 5, 2}, {2, 6, 2}, {3, 7, 2}};

/* volume: double[nx*ny*nz] C-order.  On success *verts (double[3*nv]) and *tris
 * (uint64[3*nt]) are malloc'd; free with p3d_mcubes_free. */
int p3d_mcubes_marching_cubes(const double *vol, int64_t nx, int64_t ny, int64_t nz, double iso,
                              double **verts, int64_t *nv, uint64_t **tris, int64_t *nt) {
    *verts = NULL; *tris = NULL; *nv = 0; *nt = 0;
    if (nx < 2 || ny < 2 || nz < 2) return 0;
    dvec V = {0, 0, 0};
    uvec T = {0, 0, 0};
    /* rolling per-voxel edge-id planes for layers x and x+1: slot[(y*nz+z)*3+axis] */
    const size_t plane = (size_t)ny * (size_t)nz * 3;
    int64_t *ids[2];
    ids[0] = (int64_t *)malloc(plane * sizeof(int64_t));
    ids[1] = (int64_t *)malloc(plane * sizeof(int64_t));
    if (!ids[0] || !ids[1]) return 1;
    memset(ids[0], 0xFF, plane * sizeof(int64_t));
    int rc = 0;
    for (int64_t i = 0; i + 1 < nx && !rc; ++i) {
        int64_t *cur = ids[i & 1], *nxt = ids[(i + 1) & 1];
        memset(nxt, 0xFF, plane * sizeof(int64_t));
        for (int64_t j = 0; j + 1 < ny && !rc; ++j)
            for (int64_t k = 0; k + 1 < nz; ++k) {
                double v[8];
                unsigned c = 0;
                for (int m = 0; m < 8; ++m) {
                    v[m] = vol[((i + CORNER[m][0]) * ny + (j + CORNER[m][1])) * nz + (k + CORNER[m][2])];
                    if (v[m] <= iso) c |= 1u << m;
                }
                const char *row = P3D_ORACLE_CASE_ROWS[c];
                for (const char *p = row; *p; ++p) {
                    const int e = *p <= '9' ? *p - '0' : *p - 'a' + 10;
                    const int a = EDGE[e][0], b = EDGE[e][1], axis = EDGE[e][2];
                    int64_t *slot = (CORNER[a][0] ? nxt : cur) +
                                    ((size_t)(j + CORNER[a][1]) * nz + (size_t)(k + CORNER[a][2])) * 3 + axis;
                    if (*slot < 0) {
                        double pos[3] = {(double)(i + CORNER[a][0]), (double)(j + CORNER[a][1]),
                                         (double)(k + CORNER[a][2])};
                        pos[axis] = interp(iso, v[a], v[b], pos[axis], pos[axis] + 1.0);
                        *slot = (int64_t)(V.n / 3);
                        if (dpush3(&V, pos[0], pos[1], pos[2])) { rc = 1; break; }
                    }
                    if (upush(&T, (uint64_t)*slot)) { rc = 1; break; }
                }
            }
    }
    free(ids[0]);
    free(ids[1]);
    if (rc) { free(V.v); free(T.v); return rc; }
    *verts = V.v; *nv = (int64_t)(V.n / 3);
    *tris = T.v;  *nt = (int64_t)(T.n / 3);
    return 0;
}

void p3d_mcubes_free(void *p) { free(p); }
