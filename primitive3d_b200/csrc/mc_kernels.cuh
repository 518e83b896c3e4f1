// primitive3d_b200/csrc/mc_kernels.cuh -- sm_100a kernels of the dense-grid marching cubes path.
//
// Replaces count_vertices_faces_kernel / gen_vertices_kernel / gen_faces_kernel of
// src/prim3d/Utility/marching_cubes.cu:4-209 (reference) with a design that reads the fp32
// grid ONCE and carries 1 bit per sample (+ 16 bytes per 128 samples) between two passes:
//
//   pass A  k_tile    persistent CTAs (four per SM, one staged tile each) walk 8x8x128-sample tiles
//                     (+1 halo in x, y, z) that TMA (cp.async.bulk.tensor.3d, mbarrier) stages in
//                     shared memory.  Per tile: inside bits (value > thresh, marching_cubes.cu:25), a
//                     thread per 32-sample word; crossing masks per word; triangle counts from
//                     popcounts of the edge masks (crossed edges - 2 per loop) with a table lookup
//                     only for cells with an ambiguous face; a two-level single-pass scan over
//                     tiles for the tile's first vertex id, then the vertices themselves,
//                     interpolated from the staged fp32 samples in the reference's operation
//                     order (marching_cubes.cu:105-109, :298).
//                     Side products: bit words, one 16-byte table entry per (row, 128-sample
//                     piece) = first id of its x-/y-/z-edge vertices + its triangle count.
//                     Triangle counts are also summed (RED) per chunk of 128 consecutive (row, piece)
//                     pairs in voxel-major order; k_round_sums adds them per round of 256 chunks.
//   pass B  k_faces   warps take chunks by ticket; a chunk's first face index is the sum of the
//                     rounds before its round and of the chunks before it in its round (one batch
//                     of loads of final data: no scan kernel, no CUB/thrust, no waiting).  A lane
//                     per half piece recomputes crossing masks from the bit words, ranks any cube
//                     edge as table base + popc(mask below z), and writes faces in voxel-major cell
//                     order, table order inside a cell (marching_cubes.cu:194-208).
//
// Vertex numbering (a free choice: the reference's is atomicAdd-arbitrary): tiles in a fixed
// order that depends on the grid shape only; inside a tile, (x,y) rows in C order; inside a
// row's 128-sample piece all x-edge vertices by z, then y-edge, then z-edge vertices.  The id of
// edge (voxel p, axis a) is therefore  table[row(p)][piece(p)].{vx|vy|vz} + popc(mask_a below z)
// -- no dense vertex-id volume (the reference's 12 B/voxel vertex_grids, marching_cubes.cu:257-259).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p3d {

constexpr int kTileX = 8, kTileY = 8, kTileZ = 128;  // voxels owned by a tile
constexpr int kPieceWords = kTileZ / 32;             // bit words per (row, piece)
constexpr int kBoxZ = kTileZ + 4;                    // staged samples per row: +1 halo, 16-byte multiple
constexpr int kBoxRows = (kTileX + 1) * (kTileY + 1);
constexpr int kStageBytes = ((kBoxRows * kBoxZ * 4 + 127) / 128) * 128;
constexpr int kTileThreads = 256;                    // one thread per owned bit word
constexpr int kFacePieces = 16;                      // pieces per warp iteration of the face pass (a "group")
constexpr int kFaceChunk = 8;                        // groups per ticket: 128 pieces, 4 per lane
constexpr int kRoundTiles = 256;                     // scan items (tiles / face chunks) per round of the two-level scan

struct McGeom {
    int64_t rx, ry, rz;   // local dims (rx includes the halo plane, if any)
    int64_t owned_x;      // planes whose edges / cells this launch owns
    int32_t np;           // 128-sample pieces per row = ceil(rz / 128)
    int32_t nxb, nyb;     // tile blocks along x (owned planes) and y
    int32_t band;         // y-blocks per band of the tile order
    int64_t ntiles;       // nxb * nyb * np
    int64_t npieces;      // owned_x * ry * np
    int64_t nchunks;      // ceil(npieces / 128): face-pass chunks
    int64_t nrounds;      // ceil(ntiles / 256): rounds of the tile scan
    int64_t nfrounds;     // ceil(nchunks / 256): rounds of face chunks
    uint64_t magic_np;    // floor(2^64 / np) + 1: n / np == umul64hi(n, magic_np) for n < 2^32 (np > 1)
    // block-sparse form (p3d_mc_extract_sparse): the tile pass visits only these tiles, in this order (ntiles = their
    // number), ids = (x-block * nyb + y-block) * np + piece; nullptr = every tile, in band order
    const uint32_t *tile_list;
};

// Workspace header (device).  Zeroed before every count.
struct McHeader {
    unsigned long long total_v;  // written by the last tile
    unsigned long long total_f;  // accumulated tile by tile (k_tile)
    unsigned int ticket;         // dynamic tile id of k_tile
    unsigned int ticket_faces;   // dynamic chunk id of k_faces (reset by launch_faces)
    unsigned int pad[2];
    unsigned long long vertex_base;  // sum of the lower shards' V, computed on the device from the exchange buffer
};

// State of one single-pass scan (two levels: items, rounds of 256 items).  Zeroed before use.
struct McScan {
    unsigned long long *status;        // [n] published count of each item
    unsigned long long *round_acc;     // [rounds] arrivals<<48 | sum of each round
    unsigned long long *round_prefix;  // [rounds + 1] published exclusive prefix of each round
};

struct McWorkspace {
    McHeader *header;
    McScan vscan;                  // vertex counts over tiles (k_tile)
    uint32_t *chunk_sum;           // [nchunks] triangles of each chunk of 128 consecutive pieces (k_tile, RED)
    unsigned long long *fround_sum;  // [nfrounds] triangles of each round of 256 chunks (k_tile, RED)
    uint4 *ptab;                   // [rx*ry*np] {vx, vy, vz, nf}: ids of the piece's first x-/y-/z-edge vertex (relative to
                                   // the tile until the tile's first id is known, absolute after the tile pass)
    uint32_t *nf;                  // [npieces] triangles of the four bit words of each piece, one byte per word (<= 160)
    uint32_t *bits;                // [rx*ry][4*np] inside bits, 32 samples per word
};

struct McEmitParams {
    float thresh;
    float scale[3];
    float offset[3];
    int64_t x_origin;              // global dim-0 index of local plane 0
    int32_t vertex_id_base;        // added to every face index
};

// mode 0: full pass (side products + vertices with id < vertex_capacity);
// mode 1: vertices only, after a completed mode-0 pass on the same workspace.
// dtype: p3d_dtype of the grid's elements (include/prim3d_b200.h); every sample is converted to float32 on chip.
void launch_tile_pass(const void *grid, int dtype, const McGeom &g, const McWorkspace &ws, const McEmitParams &p,
                      float *verts, int64_t vertex_capacity, int mode, cudaStream_t s);
// face_capacity: faces the buffer holds; the pass writes nothing if the workspace's F exceeds it
// vertex_base_from_header: add header->vertex_base (launch_apply_exchange) to every face index
void launch_faces(const McGeom &g, const McWorkspace &ws, const McEmitParams &p, int32_t *faces, int64_t face_capacity,
                  bool vertex_base_from_header, cudaStream_t s);
// Row-streaming form of the face pass (mc_faces_rows.cu), for rows of 17..128 bit words; launch_faces uses it
// when it applies.
bool faces_rows_applicable(const McGeom &g);
void launch_faces_rows(const McGeom &g, const McWorkspace &ws, const McEmitParams &p, int32_t *faces, int64_t face_capacity,
                       bool vertex_base_from_header, cudaStream_t s);
// Multi-GPU exchange without the host: `out` = this shard's first-plane table + {V, F} (two int64) appended;
// `gathered` = the all-gather of every shard's `out` ([world][words + 4] int32).
void launch_export_exchange(uint32_t *out, const McGeom &g, const McWorkspace &ws, cudaStream_t s);
void launch_apply_exchange(const McGeom &g, const McWorkspace &ws, const uint32_t *gathered, int rank, int world,
                           cudaStream_t s);
void launch_export_plane(uint32_t *table_out, const McGeom &g, const McWorkspace &ws, cudaStream_t s);
void launch_import_halo(const McGeom &g, const McWorkspace &ws, const uint32_t *table_in, uint32_t delta, cudaStream_t s);
// ---- small grids and batches of them: one kernel launch (mc_small.cu) ----
constexpr int kSmallMaxCtas = 1024;
struct SmallGrid {            // one grid of a batch
    const float *grid;
    float *vertices;          // nullptr: count only
    int32_t *faces;
    int64_t vertex_capacity, face_capacity;
    int64_t word0;            // index of the grid's first 32-sample word in the batch
    int32_t rx, ry, rz, wpr;  // wpr = words per row = ceil(rz / 32)
    float thresh;
    float scale[3], offset[3];
};
struct SmallBatch {
    int64_t nwords;           // 32-sample words of all grids
    int32_t ngrids;
    SmallGrid g0;             // the grid itself when ngrids == 1 (no descriptor array in device memory)
};
struct SmallHeader {          // written by the launch (no zeroing needed)
    unsigned long long total_v, total_f;
    unsigned int pad[4];
};
struct SmallWorkspace {
    SmallHeader *header;
    unsigned long long *cta_sums;   // [ctas][2] vertices / faces of a CTA's words
    unsigned long long *grid_base;  // [ngrids][2] first vertex id / face index of a grid in the batch-wide numbering
    uint4 *corner;                  // [nwords] inside bits {a, b, c, d} of the four rows of a word's cells
    uint32_t *cnt;                  // [nwords] nx | ny << 6 | nz << 12 | nf << 18 | next bits << 28
    uint2 *first;                   // [nwords] batch-wide first vertex id / first face index of a word
    unsigned int *sync;             // {barrier arrivals, exits}: library-owned, zero between launches (set by launch_small)
};
size_t small_workspace_bytes(int64_t nwords, int ngrids);
SmallWorkspace bind_small(void *base, int64_t nwords, int ngrids, SmallGrid **grids_dev);
// totals_host: optional pinned, device-visible landing place of the batch totals {V, F}.  The caller must wait for the
// stream before the same host thread launches again (the barrier words are per host thread and device); a cooperative
// launch, so that launches of several host threads cannot starve each other of SMs at the barrier.  Returns false if
// the barrier words could not be allocated or the launch was refused.
bool launch_small(const SmallBatch &b, const SmallGrid *grids_dev, const SmallWorkspace &ws, cudaStream_t s,
                  unsigned long long *totals_host);

// ---- shard-boundary exchange over peer memory (mc_peer.cu) ----
constexpr int kMaxPeers = 32;                 // ranks of one exchange (one NVLink domain)
constexpr int kPeerDoneWord = 2 * kMaxPeers;  // control words of a mailbox: flags[2][kMaxPeers], done, timeout
constexpr int kPeerTimeoutWord = kPeerDoneWord + 1;
constexpr int kPeerDataOffset = 512;          // bytes: recv[2][world][n + 1] uint4 follow the control words
struct PeerParams {
    char *peer[kMaxPeers];  // every rank's mailbox, mapped into this process (peer[rank] = my own)
    int rank, world;
    long long n;            // table entries of a plane (ry * np)
};
void launch_export_p2p(const PeerParams &pp, const McWorkspace &ws, uint32_t epoch, cudaStream_t s);
void launch_wait_p2p(char *own_base, int world, uint32_t epoch, cudaStream_t s);

const char *tile_pass_error();  // non-null if the last launch_tile_pass could not build its TMA descriptor

}  // namespace p3d
