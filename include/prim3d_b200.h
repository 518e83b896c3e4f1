/*
 * prim3d_b200.h -- C ABI of the B200-native marching cubes / marching tetrahedra hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point
 * names the reference interface it replaces (paths relative to lzhnb/Primitive3D @ 56af21e).
 * The reference-facing module prim3d.libPrim3D (primitive3d_b200/csrc/bindings.cpp) is a thin
 * pybind11 layer over these calls; INTEGRATION.md shows the binding a reference maintainer
 * would add.
 *
 * Conventions
 *   - all `grid`, `workspace`, output and table pointers are DEVICE pointers on the current
 *     CUDA device unless a parameter is documented as host memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns a p3d_status; p3d_last_error() gives the message of the last
 *     failure on the calling thread;
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     P3D_ERR_CUDA.
 */
#ifndef PRIM3D_B200_H_
#define PRIM3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P3D_ABI_VERSION 5

typedef enum p3d_status {
    P3D_OK = 0,
    P3D_ERR_INVALID = 1,   /* bad argument (null pointer, non-positive size, ...)          */
    P3D_ERR_CUDA = 2,      /* a CUDA runtime call or kernel launch failed                   */
    P3D_ERR_OVERFLOW = 3,  /* the vertex count does not fit the int32 face-index contract   */
    P3D_ERR_WORKSPACE = 4  /* the workspace passed is smaller than *_workspace_bytes()      */
} p3d_status;

int p3d_abi_version(void);
const char *p3d_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Dense-grid marching cubes.
 * Replaces prim3d::marching_cubes, src/prim3d/Utility/marching_cubes.h:14-15 and
 * marching_cubes.cu:212-305 (count_vertices_faces_kernel :4-68, gen_vertices_kernel :70-138,
 * gen_faces_kernel :140-209, bounding-box epilogue :289-298).
 *
 * The grid is float32, C-contiguous, idx = i*(ry*rz) + j*rz + k (marching_cubes.cu:20).
 * A sample is inside iff value > thresh (NaN and == thresh are outside).
 *
 * The reference needs the output sizes before it can allocate (it synchronises twice,
 * marching_cubes.cu:251-252).  Here the grid is read ONCE: the pass that classifies and counts
 * also writes the vertices, into a caller-provided buffer of speculative capacity, because the
 * fp32 samples a vertex interpolates are on chip at that moment and nowhere else later:
 *
 *   p3d_mc_count()     one pass over the grid: inside bits, per-piece tables, {V, F} to the host
 *                      (one stream synchronise) and every vertex whose id < vertex_capacity;
 *   p3d_mc_vertices()  only needed when V > vertex_capacity: re-reads the grid and writes all
 *                      vertices into an exact-size buffer (same ids, same values);
 *   p3d_mc_faces()     faces int32 [F,3] from the side products alone (asynchronous).
 *
 * Output order is deterministic.  Vertices are numbered tile by tile (8 x 8 rows x 128 samples,
 * in an order that depends on the grid shape only), inside a tile row by row, inside a row's
 * 128-sample piece all x-edge vertices by z, then y-edge, then z-edge vertices.  Faces are in
 * voxel-major cell order with the triangle-table order inside a cell.  The reference's own
 * order is atomicAdd-arbitrary (marching_cubes.cu:104,117,130,199).
 *
 * Slab decomposition (multi-GPU, dim 0): `rx` planes are present in memory, the first
 * `owned_x` of them are owned by this call, a trailing plane (rx == owned_x + 1) is a halo
 * owned by the next shard.  `x_origin` is the global index of local plane 0 and `global_rx`
 * the full-grid Rx used by the bounding-box scale.  Single GPU: owned_x = global_rx = rx,
 * x_origin = 0.
 * ---------------------------------------------------------------------------------------- */
typedef struct p3d_mc_desc {
    int64_t rx, ry, rz;  /* planes/rows/samples present in `grid`                            */
    int64_t owned_x;     /* leading planes owned by this call: rx, or rx-1 with a halo plane */
    int64_t x_origin;    /* global dim-0 index of local plane 0                              */
    int64_t global_rx;   /* full-grid Rx (bounding-box scale, marching_cubes.cu:294)         */
    float thresh;
    float lower[3];      /* bounding box, marching_cubes.cu:290-297 (host values)            */
    float upper[3];
} p3d_mc_desc;

/* Bytes of device workspace for this descriptor (1 bit per sample + 20 bytes per 128 samples
 * of a row + scan state; the reference's vertex_grids is 12 bytes per sample).  0 on an invalid
 * descriptor. */
size_t p3d_mc_workspace_bytes(const p3d_mc_desc *desc);

/* A vertex capacity that covers smooth fields without a second pass: owned samples / 16
 * (+ slack), never more than 3 per sample.  Callers that extract similar grids repeatedly
 * should pass the previous V plus a margin instead. */
int64_t p3d_mc_vertex_capacity_hint(const p3d_mc_desc *desc);

/* counts_host[0] = V (vertices owned by this shard), counts_host[1] = F (faces of its cells);
 * host memory.  vertices: float[3*vertex_capacity] (device) or NULL with capacity 0; vertex i
 * is written iff i < vertex_capacity, already scaled and offset (marching_cubes.cu:298), so
 * the buffer is final when V <= vertex_capacity.  Synchronises `stream`.
 * P3D_ERR_OVERFLOW if V > INT32_MAX. */
p3d_status p3d_mc_count(const p3d_mc_desc *desc, const float *grid, void *workspace,
                        size_t workspace_bytes, float *vertices, int64_t vertex_capacity,
                        int64_t *counts_host, void *stream);

/* vertices: float[3*vertex_capacity]; writes every vertex with id < vertex_capacity.  Must
 * follow p3d_mc_count on the same workspace and grid.  Asynchronous. */
p3d_status p3d_mc_vertices(const p3d_mc_desc *desc, const float *grid, void *workspace,
                           float *vertices, int64_t vertex_capacity, void *stream);

/* Fused dtype ingest.  The reference's Python wrapper casts any input to float32 in a separate
 * pass before the kernels see it (prim3d/utility/marching_cubes.py:86-87; examples/sphere.py
 * feeds int64).  The *_typed entry points read the grid in its own element type and convert
 * each sample to float32 (round to nearest even, exactly what the cast does) as the tile is
 * staged on chip: classification, interpolation and every output are bit-identical to
 * cast-then-run, without the extra read + write + read of the cast.  float32 grids take the TMA
 * path; the others are staged with coalesced row loads. */
typedef enum p3d_dtype {
    P3D_F32 = 0,
    P3D_F16 = 1,
    P3D_BF16 = 2,
    P3D_F64 = 3,
    P3D_I64 = 4,
    P3D_I32 = 5,
    P3D_I16 = 6,
    P3D_U8 = 7
} p3d_dtype;

p3d_status p3d_mc_count_typed(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace,
                              size_t workspace_bytes, float *vertices, int64_t vertex_capacity,
                              int64_t *counts_host, void *stream);
p3d_status p3d_mc_vertices_typed(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace,
                                 float *vertices, int64_t vertex_capacity, void *stream);

/* faces: int32[3*F] (device).  Face indices are written as vertex_id_base + local id
 * (vertex_id_base = exclusive prefix of V over lower shards; 0 on a single GPU).  Must follow
 * p3d_mc_count (and, for a slab with a halo plane, p3d_mc_import_halo_plane) on the same
 * workspace.  Asynchronous. */
p3d_status p3d_mc_faces(const p3d_mc_desc *desc, const void *workspace, int32_t *faces,
                        int64_t vertex_id_base, void *stream);

/* 1 if p3d_mc_extract handles this grid by its single-launch path (float32 grids of up to P3D_MC_SMALL_SINGLE_MAX
 * samples, default 2^20: one kernel for the whole extraction instead of the tiled passes' launch chain -- 36 us
 * against 56 us per call at the reference's bunny 66^3 example; 0 in the environment variable sends every single grid
 * through the tiled passes), else 0.  Vertices are then numbered voxel-major by (row, 32-sample word) instead of tile
 * by tile (the same mesh, other ids), and after such an extraction the workspace does NOT hold the
 * state p3d_mc_vertices / p3d_mc_faces continue from: an output that did not fit its capacity is redone by calling
 * p3d_mc_extract again with an exact buffer for it (capacity 0 for the output that did fit).  The path keeps two
 * barrier words per host thread and device in memory the library owns: a host thread's calls must not overlap, which
 * they cannot, since p3d_mc_extract and p3d_mc_extract_batch return after their one stream wait. */
int p3d_mc_single_launch(const p3d_mc_desc *desc, int dtype);

/* Whole extraction with ONE host synchronisation (single GPU: owned_x == rx).  Both passes are
 * queued back to back into buffers of speculative capacity; the host waits once, for {V, F}.
 * vertices: float[3*vertex_capacity], faces: int32[3*face_capacity] (device).  On return
 * counts_host = {V, F}; the vertex buffer is final iff V <= vertex_capacity (else call
 * p3d_mc_vertices_typed with an exact buffer) and the face buffer is final iff F <=
 * face_capacity (else nothing was written to it: call p3d_mc_faces with an exact buffer).
 * The reference synchronises twice and allocates in between (marching_cubes.cu:251-263). */
p3d_status p3d_mc_extract(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace,
                          size_t workspace_bytes, float *vertices, int64_t vertex_capacity,
                          int32_t *faces, int64_t face_capacity, int64_t *counts_host, void *stream);

/* Batched small grids (the reference's own TODO, marching_cubes.cu:255-256): num_grids independent
 * extractions queued back to back on `stream`, ONE host synchronisation for all of them.  A 66^3
 * grid costs ~25 us of GPU time but ~3x that per call when every call waits for its own counts;
 * here the wait is paid once per batch.  descs[i] / grids[i] as in p3d_mc_extract (any shapes;
 * one `dtype` for the batch); `workspace` is shared, sized for the largest grid
 * (p3d_mc_workspace_bytes); vertices[i] / faces[i] are per-grid buffers of speculative capacity
 * with the semantics of p3d_mc_extract; counts_host = int64[2*num_grids] = {V_0, F_0, V_1, ...}.
 * The pointer arrays are host arrays of device pointers.  A grid whose output did not fit is
 * redone by the caller with p3d_mc_extract into exact buffers. */
/* Workspace for p3d_mc_extract_batch: the largest single-grid workspace, or -- when every grid is small enough for
 * the single-launch path (float32, up to 4 Mi samples each, 64 Mi in total) -- what one launch over the whole batch
 * needs (32 bytes per 32 samples).  With less, p3d_mc_extract_batch runs the grids one after the other. */
size_t p3d_mc_batch_workspace_bytes(int64_t num_grids, const p3d_mc_desc *descs);
p3d_status p3d_mc_extract_batch(int64_t num_grids, const p3d_mc_desc *descs, const void *const *grids, int dtype,
                                void *workspace, size_t workspace_bytes, float *const *vertices,
                                const int64_t *vertex_capacities, int32_t *const *faces,
                                const int64_t *face_capacities, int64_t *counts_host, void *stream);

/* Block-sparse form: p3d_mc_extract over a dense grid of which only the listed TILES are examined (a level set
 * touches a few per cent of a large grid; the caller usually knows where, e.g. from the previous frame or from a
 * coarse pass).  A tile is 8 x 8 rows x 128 samples: tile id = (x / 8 * ceil(ry / 8) + y / 8) * ceil(rz / 128) + z / 128.
 *   grid       float32 (dtype = P3D_F32);
 *   tiles      device array of num_tiles DISTINCT tile ids; it must hold every tile that contains a sample whose +x,
 *              +y or +z edge is crossed, or a cell with mixed corners (then the mesh is the dense call's mesh);
 *              sorted ids keep neighbouring tiles close in time (L2 reuse of the shared halo planes);
 *   workspace  p3d_mc_workspace_bytes(desc), zeroed by the call;
 *   outputs as in p3d_mc_extract (complete iff V <= vertex_capacity and F <= face_capacity; otherwise call again
 *   with buffers of the returned sizes).  Vertices are numbered tile by tile in LIST order, faces are in voxel-major
 *   order.  The tile pass reads 32 KB of samples per listed tile instead of the whole grid; the face pass still walks
 *   every row (one count word per 128 samples).  Whole grids only (no halo plane).  The reference has no such entry. */
p3d_status p3d_mc_extract_sparse(const p3d_mc_desc *desc, const void *grid, int dtype, const uint32_t *tiles,
                                 int64_t num_tiles, void *workspace, size_t workspace_bytes, float *vertices,
                                 int64_t vertex_capacity, int32_t *faces, int64_t face_capacity,
                                 int64_t *counts_host, void *stream);

/* Profiling hook (bench.py times each kernel with CUDA events through it): runs ONE stage of
 * p3d_mc_count asynchronously on `stream` -- 0: reset scan state, 1: tile pass (classify,
 * count, look-back, vertices).  The face stage (scan over chunks + faces) is p3d_mc_faces itself. */
p3d_status p3d_mc_debug_stage(const p3d_mc_desc *desc, const float *grid, void *workspace, int stage,
                              float *vertices, int64_t vertex_capacity, void *stream);

/* Multi-GPU halo exchange of vertex numbering (16 bytes per 128-sample piece of one plane):
 * export copies the piece table of local plane 0 into table_out (uint32[
 * p3d_mc_plane_table_words(desc)], device); import installs the next shard's exported table as
 * the numbering of this shard's halo plane, shifted by `delta` = this shard's V.  Both
 * asynchronous on `stream`. */
int64_t p3d_mc_plane_table_words(const p3d_mc_desc *desc);
p3d_status p3d_mc_export_first_plane(const p3d_mc_desc *desc, const void *workspace,
                                     uint32_t *table_out, void *stream);
p3d_status p3d_mc_import_halo_plane(const p3d_mc_desc *desc, void *workspace,
                                    const uint32_t *table_in, int64_t delta, void *stream);

/* The same exchange without a host round trip between the two passes (the multi-GPU driver's fast
 * path: tile pass, exchange, face pass are all queued; the host waits once, for the gathered counts):
 *   p3d_mc_tile_async()       p3d_mc_count without the readback / synchronise;
 *   p3d_mc_export_exchange()  out = uint32[p3d_mc_exchange_words(desc)]: the first-plane table followed by
 *                             this shard's {V, F} as two int64 -- the all-gather payload of one shard;
 *   p3d_mc_faces_exchanged()  gathered = the all-gather of every shard's payload, in rank order: computes
 *                             this shard's vertex_id_base (sum of the lower shards' V), installs the next
 *                             shard's table as the halo-plane numbering (shifted by this shard's V) and runs
 *                             the face pass into `faces` (capacity as in p3d_mc_extract: nothing is written
 *                             if F exceeds it).  All on the device, asynchronous. */
p3d_status p3d_mc_tile_async(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace,
                             size_t workspace_bytes, float *vertices, int64_t vertex_capacity, void *stream);
int64_t p3d_mc_exchange_words(const p3d_mc_desc *desc);
p3d_status p3d_mc_export_exchange(const p3d_mc_desc *desc, const void *workspace, uint32_t *out, void *stream);
p3d_status p3d_mc_faces_exchanged(const p3d_mc_desc *desc, void *workspace, const uint32_t *gathered, int rank,
                                  int world, int32_t *faces, int64_t face_capacity, void *stream);

/* The whole multi-GPU extraction of ONE shard as one call (what primitive3d_b200/sharded.py does through the three
 * calls above, for C / C++ callers that hold an NCCL communicator): tile pass -> payload -> ncclAllGather over
 * `nccl_comm` (an ncclComm_t of `world` ranks, this process being `rank`; NCCL is resolved at run time from the
 * libnccl.so.2 already loaded in the process, else from the default library path) -> vertex id base and halo
 * numbering on the device -> face pass, all queued on `stream`; the host waits once, for the counts.
 *   exchange_send  device uint32[p3d_mc_exchange_words(desc)], exchange_recv device uint32[world * that]: scratch
 *   counts_host    int64[2 * world] = {V_0, F_0, V_1, F_1, ...}: this shard's vertices are global ids
 *                  [sum of V_r below rank, + V_rank), its faces (GLOBAL vertex ids) start at the sum of F_r below rank
 * Capacities as in p3d_mc_extract (vertices beyond vertex_capacity are not written: redo with p3d_mc_vertices;
 * nothing is written to `faces` if F_rank exceeds face_capacity: redo with p3d_mc_faces and the vertex id base).
 * The reference has no multi-GPU path; this is the dim-0 slab sharding of SURVEY.md section 8(e). */
p3d_status p3d_mc_sharded_extract(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace,
                                  size_t workspace_bytes, void *nccl_comm, int rank, int world, uint32_t *exchange_send,
                                  uint32_t *exchange_recv, float *vertices, int64_t vertex_capacity, int32_t *faces,
                                  int64_t face_capacity, int64_t *counts_host, void *stream);

/* The same one-call shard extraction with the boundary exchange over PEER MEMORY instead of a collective: every rank
 * owns a mailbox in device memory which the other ranks of the node map through CUDA IPC; after its tile pass a rank
 * stores its first-plane table into the mailbox of the rank below it and its {V, F} into every mailbox (NVLink /
 * NVSwitch stores from a kernel), raises a flag, and waits for the others' flags; the face pass follows on the same
 * stream.  No NCCL call, nothing of the payload travels to a rank that does not read it.
 *   p3d_mc_peer_create    allocates this rank's mailbox for shards whose planes are desc's (ry x rz) and returns its
 *                         IPC handle (p3d_mc_peer_handle_bytes() bytes) in handle_out;
 *   p3d_mc_peer_connect   handles = the handles of ALL ranks in rank order (the caller exchanges them with whatever
 *                         it has: MPI, torch.distributed, a file), maps the other ranks' mailboxes;
 *   p3d_mc_sharded_extract_p2p   arguments and results as p3d_mc_sharded_extract; COLLECTIVE: every rank of the
 *                         mailbox makes the same sequence of calls (a rank that never arrives ends the others' wait
 *                         with P3D_ERR_CUDA after a few seconds instead of hanging the device);
 *   p3d_mc_peer_disconnect  unmaps the other ranks' mailboxes (after a barrier of the caller's: no rank may still be
 *                         inside a call);
 *   p3d_mc_peer_destroy   disconnects if that has not been done, and frees this rank's mailbox.  A mailbox should not
 *                         be freed while another process still maps it: disconnect on every rank, barrier, destroy.
 * One process per GPU (IPC handles cannot be opened by the process that made them); at most 32 ranks. */
typedef struct p3d_mc_peer p3d_mc_peer;
size_t p3d_mc_peer_handle_bytes(void);
p3d_status p3d_mc_peer_create(const p3d_mc_desc *desc, int rank, int world, p3d_mc_peer **out, void *handle_out);
p3d_status p3d_mc_peer_connect(p3d_mc_peer *peer, const void *handles);
void p3d_mc_peer_disconnect(p3d_mc_peer *peer);
void p3d_mc_peer_destroy(p3d_mc_peer *peer);
p3d_status p3d_mc_sharded_extract_p2p(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace,
                                      size_t workspace_bytes, p3d_mc_peer *peer, float *vertices, int64_t vertex_capacity,
                                      int32_t *faces, int64_t face_capacity, int64_t *counts_host, void *stream);

/* Marching cubes of a grid in HOST memory, pipelined slab by slab on one device: while slab k is
 * extracted, slab k+1 uploads and the mesh of slab k-1 downloads, so the call costs about
 * max(upload, download) instead of upload + compute + download (the reference wrapper's
 * `.cuda()` ... caller's `.cpu()`, prim3d/utility/marching_cubes.py:86-95), and the device holds
 * two slabs at a time -- the grid may be larger than device memory.
 *   desc          the whole grid (owned_x = rx = global_rx, x_origin = 0);
 *   host_grid     rx*ry*rz elements of `dtype` in host memory (pinned memory for full PCIe speed);
 *   slab_planes   dim-0 planes per slab (rounded up to a multiple of 8), 0 = choose (~16 slabs);
 *   host_vertices float[3*vertex_capacity], host_faces int32[3*face_capacity] (host memory);
 *   counts_host   {V, F}.  The outputs are complete iff V <= vertex_capacity and F <= face_capacity
 *                 (otherwise call again with buffers of the returned sizes);
 *   device_arena  optional device scratch of p3d_mc_extract_host_arena_bytes() bytes, 256-byte
 *                 aligned (two slabs + their workspaces and output buffers), e.g. from the caller's
 *                 caching allocator; NULL or too small: cudaMalloc / cudaFree inside the call.
 * Vertices are numbered slab by slab (the multi-GPU numbering with world = number of slabs), faces
 * are in voxel-major order with global ids.  Uses its own (non-blocking) streams and is synchronous
 * for the caller, but does NOT order itself behind work the caller has queued: device work that still
 * writes `host_grid`, or that still uses the memory handed in as `device_arena` (a block a caching
 * allocator recycled from a tensor with kernels pending), must have completed before the call --
 * synchronise that stream first (prim3d / capi.marching_cubes_host do). */
size_t p3d_mc_extract_host_arena_bytes(const p3d_mc_desc *desc, int dtype, int64_t slab_planes);
p3d_status p3d_mc_extract_host(const p3d_mc_desc *desc, const void *host_grid, int dtype, int64_t slab_planes,
                               float *host_vertices, int64_t vertex_capacity, int32_t *host_faces,
                               int64_t face_capacity, int64_t *counts_host, void *device_arena,
                               size_t arena_bytes);

/* One-shot convenience for C/C++ callers: workspace and a vertex buffer of
 * p3d_mc_vertex_capacity_hint() through the callback, count, faces (and an exact vertex buffer
 * + p3d_mc_vertices if the hint was too small).  alloc(ctx, bytes) must return device memory on
 * the current device (or NULL). */
typedef void *(*p3d_alloc_fn)(void *ctx, size_t bytes);
p3d_status p3d_mc_run(const p3d_mc_desc *desc, const float *grid, p3d_alloc_fn alloc, void *alloc_ctx,
                      float **vertices, int32_t **faces, int64_t *num_vertices, int64_t *num_faces,
                      void *stream);

/* ------------------------------------------------------------------------------------------
 * Marching tetrahedra.
 * Replaces prim3d/utility/marching_tetrahedras.py:89-235 (a chain of ~25 torch ops in the
 * reference: gather, torch.det, torch.unique(dim=0), masked scatter, gathers).
 *
 * points float32 [P,3], tets int64 [T,4] (MUTATED IN PLACE like the reference, :148: columns
 * 0 and 1 of negatively oriented tets are swapped), sdf float32 [P]; occupancy is sdf > 0.
 * Outputs follow the reference exactly: vertex i is the i-th crossing edge in lexicographic
 * (min id, max id) order (the order torch.unique(dim=0) defines, :157-173), interpolated as
 * p0*w0 + p1*w1 with w = (-s1, s0)/(s0 - s1) (:177-189); faces int64 [F,3] list all
 * one-triangle tets first, then all two-triangle tets (:205-223); tet_idx int64 [F] is the
 * source tet of each face (:225-234).
 *
 * Output sizes are data dependent twice over (F after classification, V after the edges are
 * de-duplicated), so the call is split in three; the first two synchronise `stream`:
 *
 *   p3d_mt_classify()  orientation fix (in place), occupancy code per tet -> codes[T];
 *                      counts_host = {n1, n2, ne}: tets with one / two triangles and the
 *                      number of crossing-edge instances.  F = n1 + 2*n2.
 *   p3d_mt_index()     compacts the valid tets in order, sorts and de-duplicates the crossing
 *                      edges (hand-written LSD radix sort + look-back scans);
 *                      counts_host = {V}.
 *   p3d_mt_emit()      verts float32 [V,3], edges int64 [V,2] (the (min,max) point ids each
 *                      vertex interpolates, needed by the backward), faces int64 [F,3],
 *                      tet_idx int64 [F].  Asynchronous.
 * ---------------------------------------------------------------------------------------- */
/* Bytes of the device buffer `codes` (one byte per tet plus the classify counters). */
size_t p3d_mt_codes_bytes(int64_t num_tets);

/* oriented: 0 = fix the orientation of `tets` in place; 1 = they went through this library before
 * (see p3d_mt_extract). */
p3d_status p3d_mt_classify(const float *points, int64_t num_points, int64_t *tets, int64_t num_tets,
                           const float *sdf, int oriented, uint8_t *codes, int64_t *counts_host,
                           void *stream);

/* Bytes of device workspace p3d_mt_index / p3d_mt_emit need, from p3d_mt_classify's counts. */
size_t p3d_mt_workspace_bytes(int64_t num_tets, int64_t n1, int64_t n2, int64_t ne);

p3d_status p3d_mt_index(const int64_t *tets, int64_t num_tets, int64_t num_points, const float *sdf,
                        const uint8_t *codes, int64_t n1, int64_t n2, int64_t ne, void *workspace,
                        size_t workspace_bytes, int64_t *counts_host, void *stream);

/* edges and tet_idx may be NULL. */
p3d_status p3d_mt_emit(const float *points, const int64_t *tets, int64_t num_tets, const float *sdf,
                       const uint8_t *codes, int64_t n1, int64_t n2, int64_t ne, int64_t num_vertices,
                       const void *workspace, float *verts, int64_t *edges, int64_t *faces,
                       int64_t *tet_idx, void *stream);

/* The whole of marching tetrahedra in ONE call with one host synchronisation (three launches), for the
 * case the path exists for: an isosurface that cuts a small part of the tets (valid tets and crossing-edge
 * instances in the thousands to low millions).  Same outputs as classify + index + emit.
 *
 * Every capacity is the caller's guess (the previous call's counts plus a margin; anything for the first
 * call): the kernels count everything, write only what fits, and counts_host says what happened:
 *   counts_host[0..3] = {n1, n2, ne, V}   (F = n1 + 2*n2);
 *   counts_host[4]    = 0  outputs complete;
 *                       1  an output capacity was too small (slot_capacity < max(n1, n2),
 *                          vertex_capacity < V or face_capacity < F): call again with capacities from
 *                          the counts and oriented = 1;
 *                       2  a bucket overflowed: {n1, n2, ne} are valid, V is not (V <= ne).  If n1 + n2 is
 *                          above the slot_capacity passed or ne above the key_capacity passed, the buckets
 *                          were sized for too few entries: call again with slot_capacity = n1 + n2,
 *                          key_capacity = ne and oriented = 1.  Otherwise the ids are crowded into few
 *                          buckets: use the staged calls above (oriented = 1), which sort with the general
 *                          radix sort;
 *                       3  nothing was run: more than ~1 M valid tets or ~4 M crossing edges expected
 *                          (slot_capacity / key_capacity): use the staged calls.
 *   oriented         0: fix the orientation of `tets` in place (the reference's behaviour);
 *                    1: `tets` went through a call of this library before (state 1 or 2 above): leave
 *                       them alone.  Fixing twice is NOT the same as fixing once for a tet whose volume
 *                       is within rounding of zero, and the reference fixes once;
 *   slot_capacity    valid tets expected (sizes the slot buckets), and the capacity of each of the two
 *                    scratch lists (one-triangle / two-triangle tets);
 *   key_capacity     crossing-edge instances expected: sizes the key buckets (it bounds nothing else);
 *   verts float[3*vertex_capacity], edges int64[2*vertex_capacity] or NULL,
 *   faces int64[3*face_capacity], tet_idx int64[face_capacity] or NULL;
 *   workspace        p3d_mt_extract_workspace_bytes(...) bytes of device memory, 256-byte aligned.
 * A tet naming a point outside [0, num_points) fails the call with P3D_ERR_INVALID (the reference's
 * indexing raises a device-side assert there). */
size_t p3d_mt_extract_workspace_bytes(int64_t num_tets, int64_t num_points, int64_t slot_capacity,
                                      int64_t key_capacity);
p3d_status p3d_mt_extract(const float *points, int64_t num_points, int64_t *tets, int64_t num_tets,
                          const float *sdf, int oriented, void *workspace, size_t workspace_bytes,
                          int64_t slot_capacity, int64_t key_capacity, float *verts, int64_t *edges,
                          int64_t vertex_capacity, int64_t *faces, int64_t *tet_idx,
                          int64_t face_capacity, int64_t *counts_host, void *stream);

/* Backward of the vertex interpolation (the reference's verts are differentiable w.r.t.
 * points and sdf, marching_tetrahedras.py:175-189): accumulates into grad_points [P,3] and
 * grad_sdf [P] (both pre-zeroed by the caller) from grad_verts [V,3] and edges [V,2]. */
p3d_status p3d_mt_backward(const float *points, const float *sdf, const int64_t *edges,
                           int64_t num_vertices, const float *grad_verts, float *grad_points,
                           float *grad_sdf, void *stream);

/* ------------------------------------------------------------------------------------------
 * Binary PLY body, assembled on the device.
 * Replaces the per-vertex / per-face ofstream.write loops of prim3d::save_mesh_as_ply,
 * src/prim3d/Utility/marching_cubes.cu:333-349 (same bytes): vertex_records receives
 * num_vertices records of 15 bytes {float x, y, z; uchar r, g, b} (buffer of 15*num_vertices
 * bytes rounded up to a multiple of 4, 4-byte aligned), face_records num_faces records of
 * 16 bytes {int32 3, a, b, c} (16-byte aligned).  All pointers are device pointers;
 * asynchronous on `stream`.
 * ---------------------------------------------------------------------------------------- */
p3d_status p3d_ply_pack(const float *vertices, const uint8_t *colors, int64_t num_vertices,
                        const int32_t *faces, int64_t num_faces, void *vertex_records,
                        void *face_records, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PRIM3D_B200_H_ */
