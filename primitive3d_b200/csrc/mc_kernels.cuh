// primitive3d_b200/csrc/mc_kernels.cuh -- sm_100a kernels of the dense-grid marching cubes path.
//
// Replaces count_vertices_faces_kernel / gen_vertices_kernel / gen_faces_kernel of
// src/prim3d/Utility/marching_cubes.cu:4-209 (reference) with a design that reads the fp32
// grid once and carries 1 bit per sample between passes:
//
//   K1 classify     grid (fp32, streamed once, 128-bit loads) -> inside-bit words, 32 samples
//                   of one z-row per word.  inside = value > thresh (marching_cubes.cu:25).
//   K2 count+scan   per (x,y) row: popcounts of the x/y/z crossing masks and the triangle
//                   counts of the row's cells, all from the bit words; one decoupled
//                   look-back scan over CTA tiles turns them into absolute offsets in the
//                   same launch (no CUB/thrust).
//   K3 emit         per row: recomputes the masks, ranks every crossing edge with
//                   popc + warp scans, fetches the two fp32 endpoints only for crossing
//                   edges, interpolates in the reference's fp32 operation order
//                   (marching_cubes.cu:105-109, :298) and writes faces in voxel-major order.
//
// Vertex numbering (a free choice: the reference's is atomicAdd-arbitrary): rows in C order;
// inside row r = (x,y): all x-edge vertices by z, then all y-edge, then all z-edge vertices.
// The id of the edge (voxel p, axis a) is therefore
//     rowv[row(p)].{vx|vy|vz} + popc(mask_a(row) below z)
// which any cell can evaluate for its 12 edges from the bit words of its 2x2 rows plus the
// 16-byte row-table entries of those rows -- no dense vertex-id volume (the reference's
// 12 B/voxel vertex_grids, marching_cubes.cu:257-259).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p3d {

constexpr int kRowsPerTile = 8;          // warps per CTA in K2/K3: one (x,y) row per warp
constexpr int kPieceWords = 32;          // a warp handles a row in pieces of 32 words = 1024 samples

struct McGeom {
    int64_t rx, ry, rz;   // local dims
    int64_t owned_x;      // planes whose rows this launch owns
    int32_t wz;           // bit words per row = ceil(rz/32)
    int32_t pieces;       // ceil(wz/32)
    int64_t owned_rows;   // owned_x * ry
    int64_t num_tiles;    // ceil(owned_rows / kRowsPerTile)
};

// Workspace header (device).  Zeroed before every count.
struct McHeader {
    unsigned long long total_v;    // inclusive totals written by the last tile
    unsigned long long total_f;
    unsigned int ticket;           // dynamic tile id for the look-back scan
    unsigned int ticket_emit;
    unsigned int pad[2];
};

struct McWorkspace {
    McHeader *header;
    uint32_t *bits;                // [rx*ry][wz]
    uint4 *rowv;                   // [rx*ry] {vx, vy, vz, nf(low 32)}: first id of the row's x/y/z-edge vertices
    unsigned long long *rowf;      // [rx*ry] first face of the row's cells
    unsigned long long *status_v;  // [num_tiles] look-back status words
    unsigned long long *status_f;  // [num_tiles]
};

struct McEmitParams {
    float thresh;
    float scale[3];
    float offset[3];
    int64_t x_origin;              // global dim-0 index of local plane 0
    int32_t vertex_id_base;        // added to every face index
};

void launch_classify(const float *grid, const McGeom &g, float thresh, uint32_t *bits, cudaStream_t s);
void launch_count_scan(const McGeom &g, const McWorkspace &ws, cudaStream_t s);
void launch_emit(const float *grid, const McGeom &g, const McWorkspace &ws, const McEmitParams &p,
                 float *verts, int32_t *faces, cudaStream_t s);
void launch_import_halo(uint4 *halo_rows, const uint32_t *table_in, int64_t ry, uint32_t delta, cudaStream_t s);

}  // namespace p3d
