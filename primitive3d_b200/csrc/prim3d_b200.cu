// primitive3d_b200/csrc/prim3d_b200.cu -- host side of the C ABI declared in include/prim3d_b200.h.
// Torch-free: plain CUDA runtime.  Replaces the orchestration of prim3d::marching_cubes,
// /root/reference/src/prim3d/Utility/marching_cubes.cu:212-305.
#include "../../include/prim3d_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <climits>
#include <cstdlib>
#include <vector>
#include <cstdio>
#include <cstring>
#include <string>

#include "mc_kernels.cuh"
#include "p3d_error.h"

#ifndef P3D_MC_SMALL_MAX_DEFAULT
#define P3D_MC_SMALL_MAX_DEFAULT (4 << 20)  // samples: grids up to this size take the single-launch path
#endif

namespace {

thread_local std::string g_last_error;

p3d_status fail(p3d_status st, const std::string &msg) { return p3d::set_error(st, msg); }

#define P3D_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(P3D_ERR_CUDA, std::string(#expr " failed: ") + cudaGetErrorString(e_));          \
    } while (0)

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

bool make_geom(const p3d_mc_desc *d, p3d::McGeom *g) {
    if (!d || d->rx < 1 || d->ry < 1 || d->rz < 1) return false;
    if (d->owned_x < 0 || d->owned_x > d->rx || d->rx - d->owned_x > 1) return false;
    if (d->rx > (int64_t)INT_MAX - 64 || d->ry > (int64_t)INT_MAX - 64 || d->rz > (int64_t)INT_MAX - 256) return false;
    g->rx = d->rx;
    g->ry = d->ry;
    g->rz = d->rz;
    g->owned_x = d->owned_x;
    g->np = (int32_t)((d->rz + p3d::kTileZ - 1) / p3d::kTileZ);
    g->nxb = (int32_t)((d->owned_x + p3d::kTileX - 1) / p3d::kTileX);
    g->nyb = (int32_t)((d->ry + p3d::kTileY - 1) / p3d::kTileY);
    // tile order: bands of y-blocks sized so that one x-block of a band is ~16 MB of samples; the halo plane
    // a block shares with the next x-block is then still in L2 when that block reads it
    const int64_t column_bytes = (int64_t)p3d::kTileX * p3d::kTileY * 4 * d->rz;
    int64_t band = ((int64_t)16 << 20) / column_bytes;
    if (band < 1) band = 1;
    if (band > g->nyb) band = g->nyb;
    g->band = (int32_t)band;
    g->ntiles = (int64_t)g->nxb * g->nyb * g->np;
    g->npieces = d->owned_x * d->ry * g->np;
    g->nchunks = (g->npieces + p3d::kFacePieces * p3d::kFaceChunk - 1) / (p3d::kFacePieces * p3d::kFaceChunk);
    g->nrounds = (g->ntiles + p3d::kRoundTiles - 1) / p3d::kRoundTiles;
    g->nfrounds = (g->nchunks + p3d::kRoundTiles - 1) / p3d::kRoundTiles;
    g->magic_np = g->np > 1 ? ~0ull / (uint64_t)g->np + 1 : 0;
    g->tile_list = nullptr;
    if (g->ntiles > ((int64_t)1 << 31) || d->rx * d->ry * (int64_t)g->np > ((int64_t)1 << 40)) return false;
    return true;
}

struct Layout {
    size_t header, status, round_acc, round_prefix, chunk_sum, fround_sum, zero_end, ptab, nf, bits, total;
};

Layout make_layout(const p3d::McGeom &g) {
    Layout l;
    const size_t all_pieces = (size_t)(g.rx * g.ry) * (size_t)g.np;
    size_t off = 0;
    l.header = off;   off += align_up(sizeof(p3d::McHeader));
    l.status = off;   off += align_up((size_t)g.ntiles * 8);
    l.round_acc = off;    off += align_up((size_t)g.nrounds * 8);
    l.round_prefix = off; off += align_up((size_t)(g.nrounds + 1) * 8);
    l.chunk_sum = off;    off += align_up((size_t)g.nchunks * 4);
    l.fround_sum = off;   off += align_up((size_t)g.nfrounds * 8);
    l.zero_end = off;  // everything above is zeroed before a count
    l.ptab = off;     off += align_up(all_pieces * sizeof(uint4));
    l.nf = off;       off += align_up((size_t)g.npieces * 4 + 32);
    l.bits = off;     off += align_up(all_pieces * 16);
    l.total = off;
    return l;
}

p3d::McWorkspace bind(void *base, const Layout &l) {
    char *b = static_cast<char *>(base);
    p3d::McWorkspace ws;
    ws.header = reinterpret_cast<p3d::McHeader *>(b + l.header);
    ws.vscan.status = reinterpret_cast<unsigned long long *>(b + l.status);
    ws.vscan.round_acc = reinterpret_cast<unsigned long long *>(b + l.round_acc);
    ws.vscan.round_prefix = reinterpret_cast<unsigned long long *>(b + l.round_prefix);
    ws.chunk_sum = reinterpret_cast<uint32_t *>(b + l.chunk_sum);
    ws.fround_sum = reinterpret_cast<unsigned long long *>(b + l.fround_sum);
    ws.ptab = reinterpret_cast<uint4 *>(b + l.ptab);
    ws.nf = reinterpret_cast<uint32_t *>(b + l.nf);
    ws.bits = reinterpret_cast<uint32_t *>(b + l.bits);
    return ws;
}

p3d::McEmitParams make_params(const p3d_mc_desc *desc, int64_t vertex_id_base) {
    p3d::McEmitParams prm;
    prm.thresh = desc->thresh;
    // marching_cubes.cu:290-297 -- offset = lower; scale = (upper - lower) / resolution, where the
    // reference's y term reads upper[2] (not upper[1]); replicated because it is observable.
    prm.scale[0] = (desc->upper[0] - desc->lower[0]) / static_cast<float>(desc->global_rx);
    prm.scale[1] = (desc->upper[2] - desc->lower[1]) / static_cast<float>(desc->ry);
    prm.scale[2] = (desc->upper[2] - desc->lower[2]) / static_cast<float>(desc->rz);
    prm.offset[0] = desc->lower[0];
    prm.offset[1] = desc->lower[1];
    prm.offset[2] = desc->lower[2];
    prm.x_origin = desc->x_origin;
    prm.vertex_id_base = static_cast<int32_t>(vertex_id_base);
    return prm;
}

// ---- small grids: one kernel launch for the whole extraction (mc_small.cu) ----
int64_t small_max_samples() {
    static const int64_t v = [] {
        const char *e = getenv("P3D_MC_SMALL_MAX");  // samples; 0 sends every grid through the tiled path
        return e ? (int64_t)atoll(e) : (int64_t)P3D_MC_SMALL_MAX_DEFAULT;
    }();
    return v;
}

// A single grid of up to P3D_MC_SMALL_SINGLE_MAX samples (default 2^20, about 100^3) goes through the single-launch kernel:
// measured on B200 through the pybind module, bunny 66^3 costs 36 us per call there against 56 us for the tiled passes'
// launch chain (kernel 22 us; no memset node before it, no copy node behind it); at 128^3 the two are level (sphere: 66
// against 59 us, gyroid: 97 against 98 us), above that the tiled passes win.  0 sends every single grid through the
// tiled passes (one vertex numbering for all sizes).
#ifndef P3D_MC_SMALL_SINGLE_MAX_DEFAULT
#define P3D_MC_SMALL_SINGLE_MAX_DEFAULT (1 << 20)
#endif
int64_t small_single_max_samples() {
    static const int64_t v = [] {
        const char *e = getenv("P3D_MC_SMALL_SINGLE_MAX");
        return e ? (int64_t)atoll(e) : (int64_t)P3D_MC_SMALL_SINGLE_MAX_DEFAULT;
    }();
    return v;
}

inline int64_t small_words(const p3d_mc_desc *d) { return d->rx * d->ry * ((d->rz + 31) / 32); }

bool small_applicable(const p3d_mc_desc *d, int dtype, int64_t max_samples) {
    return dtype == P3D_F32 && d->rx == d->owned_x && d->x_origin == 0 && d->global_rx == d->rx && d->rx >= 1 && d->ry >= 1 &&
           d->rz >= 1 && d->rx * d->ry * d->rz <= max_samples;
}
bool small_applicable(const p3d_mc_desc *d, int dtype) { return small_applicable(d, dtype, small_max_samples()); }
bool small_single(const p3d_mc_desc *d, int dtype) { return small_applicable(d, dtype, small_single_max_samples()); }

p3d::SmallGrid small_grid(const p3d_mc_desc *d, const void *grid, float *vertices, int64_t vcap, int32_t *faces, int64_t fcap,
                          int64_t word0) {
    const p3d::McEmitParams prm = make_params(d, 0);
    p3d::SmallGrid g;
    g.grid = static_cast<const float *>(grid);
    g.vertices = vcap > 0 ? vertices : nullptr;
    g.faces = fcap > 0 ? faces : nullptr;
    g.vertex_capacity = vcap, g.face_capacity = fcap;
    g.word0 = word0;
    g.rx = (int32_t)d->rx, g.ry = (int32_t)d->ry, g.rz = (int32_t)d->rz, g.wpr = (int32_t)((d->rz + 31) / 32);
    g.thresh = d->thresh;
    for (int i = 0; i < 3; ++i) g.scale[i] = prm.scale[i], g.offset[i] = prm.offset[i];
    return g;
}

// One pinned landing pad per host thread for the {V,F} readbacks (grown on demand for batches).
int64_t *pinned_counts(size_t pairs = 1) {
    thread_local int64_t *buf = nullptr;
    thread_local size_t cap = 0;
    if (pairs > cap) {
        if (buf) cudaFreeHost(buf);
        cap = pairs < 64 ? 64 : pairs * 2;
        if (cudaHostAlloc(reinterpret_cast<void **>(&buf), cap * 2 * sizeof(int64_t), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            buf = nullptr;
            cap = 0;
        }
    }
    return buf;
}

}  // namespace

namespace p3d {
p3d_status set_error(p3d_status st, const std::string &msg) {
    g_last_error = msg;
    return st;
}
}  // namespace p3d

extern "C" {

int p3d_abi_version(void) { return P3D_ABI_VERSION; }

const char *p3d_last_error(void) { return g_last_error.c_str(); }

size_t p3d_mc_workspace_bytes(const p3d_mc_desc *desc) {
    p3d::McGeom g;
    if (!make_geom(desc, &g)) return 0;
    size_t n = make_layout(g).total;
    if (small_single(desc, P3D_F32)) {
        const size_t m = p3d::small_workspace_bytes(small_words(desc), 1);
        if (m > n) n = m;
    }
    return n;
}

int p3d_mc_single_launch(const p3d_mc_desc *desc, int dtype) {
    p3d::McGeom g;
    return make_geom(desc, &g) && small_single(desc, dtype) ? 1 : 0;
}

size_t p3d_mc_batch_workspace_bytes(int64_t num_grids, const p3d_mc_desc *descs) {
    if (num_grids <= 0 || !descs) return 0;
    size_t n = 0;
    int64_t words = 0;
    bool small = true;
    for (int64_t i = 0; i < num_grids; ++i) {
        const size_t m = p3d_mc_workspace_bytes(&descs[i]);
        if (!m) return 0;
        if (m > n) n = m;
        small = small && small_applicable(&descs[i], P3D_F32);
        words += small_words(&descs[i]);
    }
    if (small && words * 32 <= 16 * small_max_samples()) {
        const size_t m = p3d::small_workspace_bytes(words, (int)num_grids);
        if (m > n) n = m;
    }
    return n;
}

int64_t p3d_mc_plane_table_words(const p3d_mc_desc *desc) {
    p3d::McGeom g;
    if (!make_geom(desc, &g)) return 0;
    return g.ry * (int64_t)g.np * 4;
}

p3d_status p3d_mc_count(const p3d_mc_desc *desc, const float *grid, void *workspace, size_t workspace_bytes,
                        float *vertices, int64_t vertex_capacity, int64_t *counts_host, void *stream) {
    return p3d_mc_count_typed(desc, grid, P3D_F32, workspace, workspace_bytes, vertices, vertex_capacity, counts_host, stream);
}

p3d_status p3d_mc_count_typed(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace, size_t workspace_bytes,
                              float *vertices, int64_t vertex_capacity, int64_t *counts_host, void *stream) {
    p3d::McGeom g;
    if (dtype < P3D_F32 || dtype > P3D_U8) return fail(P3D_ERR_INVALID, "p3d_mc_count: unknown dtype");
    if (!make_geom(desc, &g)) return fail(P3D_ERR_INVALID, "p3d_mc_count: invalid descriptor");
    if (!grid || !workspace || !counts_host) return fail(P3D_ERR_INVALID, "p3d_mc_count: null pointer");
    if (desc->global_rx < 1) return fail(P3D_ERR_INVALID, "p3d_mc_count: global_rx must be >= 1");
    if (vertex_capacity < 0 || (vertex_capacity > 0 && !vertices))
        return fail(P3D_ERR_INVALID, "p3d_mc_count: vertex_capacity without a vertex buffer");
    const Layout l = make_layout(g);
    if (workspace_bytes < l.total) return fail(P3D_ERR_WORKSPACE, "p3d_mc_count: workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace) % kAlign) return fail(P3D_ERR_INVALID, "p3d_mc_count: workspace must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const p3d::McWorkspace ws = bind(workspace, l);

    // header, the tile scan state and the chunk sums are contiguous: one memset
    P3D_CUDA(cudaMemsetAsync(workspace, 0, l.zero_end, s));
    p3d::launch_tile_pass(grid, dtype, g, ws, make_params(desc, 0), vertices, vertex_capacity, 0, s);
    if (p3d::tile_pass_error()) return fail(P3D_ERR_CUDA, std::string("p3d_mc_count: ") + p3d::tile_pass_error());
    P3D_CUDA(cudaGetLastError());

    int64_t *pin = pinned_counts();
    int64_t *dst = pin ? pin : counts_host;
    P3D_CUDA(cudaMemcpyAsync(dst, &ws.header->total_v, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    P3D_CUDA(cudaStreamSynchronize(s));
    counts_host[0] = dst[0];
    counts_host[1] = dst[1];
    if (counts_host[0] > INT32_MAX)
        return fail(P3D_ERR_OVERFLOW, "p3d_mc_count: vertex count exceeds the int32 face-index contract");
    return P3D_OK;
}

p3d_status p3d_mc_vertices(const p3d_mc_desc *desc, const float *grid, void *workspace, float *vertices,
                           int64_t vertex_capacity, void *stream) {
    return p3d_mc_vertices_typed(desc, grid, P3D_F32, workspace, vertices, vertex_capacity, stream);
}

p3d_status p3d_mc_vertices_typed(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace, float *vertices,
                                 int64_t vertex_capacity, void *stream) {
    p3d::McGeom g;
    if (dtype < P3D_F32 || dtype > P3D_U8) return fail(P3D_ERR_INVALID, "p3d_mc_vertices: unknown dtype");
    if (!make_geom(desc, &g)) return fail(P3D_ERR_INVALID, "p3d_mc_vertices: invalid descriptor");
    if (!grid || !workspace || (!vertices && vertex_capacity > 0)) return fail(P3D_ERR_INVALID, "p3d_mc_vertices: null pointer");
    if (desc->global_rx < 1) return fail(P3D_ERR_INVALID, "p3d_mc_vertices: global_rx must be >= 1");
    const Layout l = make_layout(g);
    const p3d::McWorkspace ws = bind(workspace, l);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    P3D_CUDA(cudaMemsetAsync(&ws.header->ticket, 0, sizeof(unsigned int), s));
    p3d::launch_tile_pass(grid, dtype, g, ws, make_params(desc, 0), vertices, vertex_capacity, 1, s);
    if (p3d::tile_pass_error()) return fail(P3D_ERR_CUDA, std::string("p3d_mc_vertices: ") + p3d::tile_pass_error());
    P3D_CUDA(cudaGetLastError());
    return P3D_OK;
}

p3d_status p3d_mc_faces(const p3d_mc_desc *desc, const void *workspace, int32_t *faces, int64_t vertex_id_base,
                        void *stream) {
    p3d::McGeom g;
    if (!make_geom(desc, &g)) return fail(P3D_ERR_INVALID, "p3d_mc_faces: invalid descriptor");
    if (!workspace) return fail(P3D_ERR_INVALID, "p3d_mc_faces: null pointer");
    if (vertex_id_base < 0 || vertex_id_base > INT32_MAX) return fail(P3D_ERR_OVERFLOW, "p3d_mc_faces: vertex_id_base out of int32 range");
    const Layout l = make_layout(g);
    const p3d::McWorkspace ws = bind(const_cast<void *>(workspace), l);
    p3d::McEmitParams prm = make_params(desc, vertex_id_base);
    p3d::launch_faces(g, ws, prm, faces, INT64_MAX, false, static_cast<cudaStream_t>(stream));
    P3D_CUDA(cudaGetLastError());
    return P3D_OK;
}

p3d_status p3d_mc_extract(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace, size_t workspace_bytes,
                          float *vertices, int64_t vertex_capacity, int32_t *faces, int64_t face_capacity,
                          int64_t *counts_host, void *stream) {
    p3d::McGeom g;
    if (!make_geom(desc, &g)) return fail(P3D_ERR_INVALID, "p3d_mc_extract: invalid descriptor");
    if (dtype < P3D_F32 || dtype > P3D_U8) return fail(P3D_ERR_INVALID, "p3d_mc_extract: unknown dtype");
    if (!grid || !workspace || !counts_host) return fail(P3D_ERR_INVALID, "p3d_mc_extract: null pointer");
    if (desc->global_rx < 1) return fail(P3D_ERR_INVALID, "p3d_mc_extract: global_rx must be >= 1");
    if (desc->rx != desc->owned_x) return fail(P3D_ERR_INVALID, "p3d_mc_extract: a slab with a halo plane needs the staged calls");
    if (vertex_capacity < 0 || (vertex_capacity > 0 && !vertices) || face_capacity < 0 || (face_capacity > 0 && !faces))
        return fail(P3D_ERR_INVALID, "p3d_mc_extract: capacity without a buffer");
    const Layout l = make_layout(g);
    if (workspace_bytes < l.total) return fail(P3D_ERR_WORKSPACE, "p3d_mc_extract: workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace) % kAlign) return fail(P3D_ERR_INVALID, "p3d_mc_extract: workspace must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (small_single(desc, dtype) && workspace_bytes >= p3d::small_workspace_bytes(small_words(desc), 1)) {
        // small grid: the whole extraction is ONE kernel launch (mc_small.cu)
        p3d::SmallGrid *unused = nullptr;
        const p3d::SmallWorkspace sw = p3d::bind_small(workspace, small_words(desc), 1, &unused);
        p3d::SmallBatch b;
        b.nwords = small_words(desc), b.ngrids = 1;
        b.g0 = small_grid(desc, grid, vertices, vertex_capacity, faces, face_capacity, 0);
        // the kernel writes {V, F} straight into the pinned landing pad (mapped into the device's address space): no
        // memset node before it, no copy node behind it; the host reads the counts as soon as the stream is idle
        int64_t *pin = pinned_counts();
        if (!p3d::launch_small(b, nullptr, sw, s, reinterpret_cast<unsigned long long *>(pin)))
            return fail(P3D_ERR_CUDA, "p3d_mc_extract: the single-launch kernel could not be launched (barrier words or cooperative launch)");
        P3D_CUDA(cudaGetLastError());
        if (!pin) P3D_CUDA(cudaMemcpyAsync(counts_host, &sw.header->total_v, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        P3D_CUDA(cudaStreamSynchronize(s));
        if (pin) counts_host[0] = pin[0], counts_host[1] = pin[1];
        return P3D_OK;
    }
    const p3d::McWorkspace ws = bind(workspace, l);

    // everything is queued before the host waits: the face pass starts the moment the tile pass ends
    P3D_CUDA(cudaMemsetAsync(workspace, 0, l.zero_end, s));
    p3d::launch_tile_pass(grid, dtype, g, ws, make_params(desc, 0), vertices, vertex_capacity, 0, s);
    if (p3d::tile_pass_error()) return fail(P3D_ERR_CUDA, std::string("p3d_mc_extract: ") + p3d::tile_pass_error());
    if (faces) p3d::launch_faces(g, ws, make_params(desc, 0), faces, face_capacity, false, s);
    P3D_CUDA(cudaGetLastError());
    int64_t *pin = pinned_counts();
    int64_t *dst = pin ? pin : counts_host;
    P3D_CUDA(cudaMemcpyAsync(dst, &ws.header->total_v, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    P3D_CUDA(cudaStreamSynchronize(s));
    counts_host[0] = dst[0];
    counts_host[1] = dst[1];
    if (counts_host[0] > INT32_MAX)
        return fail(P3D_ERR_OVERFLOW, "p3d_mc_extract: vertex count exceeds the int32 face-index contract");
    return P3D_OK;
}

p3d_status p3d_mc_extract_sparse(const p3d_mc_desc *desc, const void *grid, int dtype, const uint32_t *tiles, int64_t num_tiles,
                                 void *workspace, size_t workspace_bytes, float *vertices, int64_t vertex_capacity, int32_t *faces,
                                 int64_t face_capacity, int64_t *counts_host, void *stream) {
    p3d::McGeom g;
    if (!make_geom(desc, &g)) return fail(P3D_ERR_INVALID, "p3d_mc_extract_sparse: invalid descriptor");
    if (dtype != P3D_F32) return fail(P3D_ERR_INVALID, "p3d_mc_extract_sparse: float32 grids only");
    if (!grid || !workspace || !counts_host || num_tiles < 0 || (num_tiles > 0 && !tiles))
        return fail(P3D_ERR_INVALID, "p3d_mc_extract_sparse: null pointer or negative tile count");
    if (desc->global_rx < 1 || desc->rx != desc->owned_x) return fail(P3D_ERR_INVALID, "p3d_mc_extract_sparse: whole grids only");
    if (num_tiles > g.ntiles) return fail(P3D_ERR_INVALID, "p3d_mc_extract_sparse: more tiles listed than the grid has");
    if (vertex_capacity < 0 || (vertex_capacity > 0 && !vertices) || face_capacity < 0 || (face_capacity > 0 && !faces))
        return fail(P3D_ERR_INVALID, "p3d_mc_extract_sparse: capacity without a buffer");
    const Layout l = make_layout(g);  // sized for the dense grid: the per-piece tables are indexed by position
    if (workspace_bytes < l.total) return fail(P3D_ERR_WORKSPACE, "p3d_mc_extract_sparse: workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace) % kAlign) return fail(P3D_ERR_INVALID, "p3d_mc_extract_sparse: workspace must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const p3d::McWorkspace ws = bind(workspace, l);
    // the tables and bit words of the tiles that are NOT visited must read as "nothing here": everything is zeroed
    P3D_CUDA(cudaMemsetAsync(workspace, 0, l.total, s));
    counts_host[0] = counts_host[1] = 0;
    if (num_tiles == 0) return P3D_OK;
    p3d::McGeom gs = g;       // the tile pass walks the list (its scan runs over list positions) ...
    gs.tile_list = tiles;
    gs.ntiles = num_tiles;
    gs.nrounds = (num_tiles + p3d::kRoundTiles - 1) / p3d::kRoundTiles;
    p3d::launch_tile_pass(grid, dtype, gs, ws, make_params(desc, 0), vertices, vertex_capacity, 0, s);
    if (p3d::tile_pass_error()) return fail(P3D_ERR_CUDA, std::string("p3d_mc_extract_sparse: ") + p3d::tile_pass_error());
    // ... the face pass walks every row of the grid: pieces without triangles cost it a count word each
    if (faces) p3d::launch_faces(g, ws, make_params(desc, 0), faces, face_capacity, false, s);
    P3D_CUDA(cudaGetLastError());
    int64_t *pin = pinned_counts();
    int64_t *dst = pin ? pin : counts_host;
    P3D_CUDA(cudaMemcpyAsync(dst, &ws.header->total_v, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    P3D_CUDA(cudaStreamSynchronize(s));
    counts_host[0] = dst[0];
    counts_host[1] = dst[1];
    if (counts_host[0] > INT32_MAX)
        return fail(P3D_ERR_OVERFLOW, "p3d_mc_extract_sparse: vertex count exceeds the int32 face-index contract");
    return P3D_OK;
}

p3d_status p3d_mc_extract_batch(int64_t num_grids, const p3d_mc_desc *descs, const void *const *grids, int dtype,
                                void *workspace, size_t workspace_bytes, float *const *vertices,
                                const int64_t *vertex_capacities, int32_t *const *faces, const int64_t *face_capacities,
                                int64_t *counts_host, void *stream) {
    if (num_grids < 0) return fail(P3D_ERR_INVALID, "p3d_mc_extract_batch: negative batch size");
    if (num_grids == 0) return P3D_OK;
    if (!descs || !grids || !workspace || !vertices || !vertex_capacities || !faces || !face_capacities || !counts_host)
        return fail(P3D_ERR_INVALID, "p3d_mc_extract_batch: null pointer");
    if (dtype < P3D_F32 || dtype > P3D_U8) return fail(P3D_ERR_INVALID, "p3d_mc_extract_batch: unknown dtype");
    if (reinterpret_cast<uintptr_t>(workspace) % kAlign) return fail(P3D_ERR_INVALID, "p3d_mc_extract_batch: workspace must be 256-byte aligned");
    int64_t *pin = pinned_counts((size_t)num_grids + 1);
    if (!pin) return fail(P3D_ERR_CUDA, "p3d_mc_extract_batch: pinned allocation failed");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int64_t i = 0; i < num_grids; ++i) {  // validate everything before anything is queued
        p3d::McGeom g;
        if (!make_geom(&descs[i], &g) || descs[i].rx != descs[i].owned_x || descs[i].global_rx < 1 || !grids[i])
            return fail(P3D_ERR_INVALID, "p3d_mc_extract_batch: invalid descriptor or grid");
        if (make_layout(g).total > workspace_bytes) return fail(P3D_ERR_WORKSPACE, "p3d_mc_extract_batch: workspace too small");
        if (vertex_capacities[i] < 0 || face_capacities[i] < 0 || (vertex_capacities[i] && !vertices[i]) ||
            (face_capacities[i] && !faces[i]))
            return fail(P3D_ERR_INVALID, "p3d_mc_extract_batch: capacity without a buffer");
    }
    // Small grids: ONE kernel launch for the whole batch (mc_small.cu), numbering restarts at every grid.
    {
        bool small = dtype == P3D_F32;
        int64_t words = 0;
        for (int64_t i = 0; small && i < num_grids; ++i) {
            small = small_applicable(&descs[i], dtype);
            words += small_words(&descs[i]);
        }
        if (small && words * 32 <= 16 * small_max_samples() && workspace_bytes >= p3d::small_workspace_bytes(words, (int)num_grids)) {
            p3d::SmallGrid *grids_dev = nullptr;
            const p3d::SmallWorkspace sw = p3d::bind_small(workspace, words, (int)num_grids, &grids_dev);
            // descriptors: staged in pinned memory (one buffer per host thread), copied ahead of the launch
            thread_local p3d::SmallGrid *stage = nullptr;
            thread_local size_t stage_cap = 0;
            if ((size_t)num_grids > stage_cap) {
                if (stage) cudaFreeHost(stage);
                stage_cap = (size_t)num_grids * 2;
                if (cudaHostAlloc(reinterpret_cast<void **>(&stage), stage_cap * sizeof(p3d::SmallGrid), cudaHostAllocPortable) != cudaSuccess) {
                    stage = nullptr, stage_cap = 0;
                    return fail(P3D_ERR_CUDA, "p3d_mc_extract_batch: pinned allocation failed");
                }
            }
            int64_t w0 = 0;
            for (int64_t i = 0; i < num_grids; ++i) {
                stage[i] = small_grid(&descs[i], grids[i], vertices[i], vertex_capacities[i], faces[i], face_capacities[i], w0);
                w0 += small_words(&descs[i]);
            }
            p3d::SmallBatch b;
            b.nwords = words, b.ngrids = (int32_t)num_grids;
            b.g0 = stage[0];
            P3D_CUDA(cudaMemcpyAsync(grids_dev, stage, (size_t)num_grids * sizeof(p3d::SmallGrid), cudaMemcpyHostToDevice, s));
            if (!p3d::launch_small(b, grids_dev, sw, s, nullptr))
                return fail(P3D_ERR_CUDA, "p3d_mc_extract_batch: the single-launch kernel could not be launched (barrier words or cooperative launch)");
            P3D_CUDA(cudaGetLastError());
            // per-grid counts = differences of the grids' bases (the batch totals close the last one)
            std::vector<int64_t> base(2 * (size_t)num_grids);
            P3D_CUDA(cudaMemcpyAsync(pin, sw.grid_base, 2 * (size_t)num_grids * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
            P3D_CUDA(cudaMemcpyAsync(pin + 2 * num_grids, &sw.header->total_v, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
            P3D_CUDA(cudaStreamSynchronize(s));
            for (int64_t i = 0; i < num_grids; ++i)
                for (int k = 0; k < 2; ++k) base[2 * i + k] = pin[2 * (i + 1) + k] - pin[2 * i + k];
            for (int64_t i = 0; i < 2 * num_grids; ++i) counts_host[i] = base[i];
            return P3D_OK;
        }
    }
    // The grids run one after the other on the stream and share the workspace (a grid's face pass has finished
    // with it before the next grid's memset starts); the host waits once, for all the counts.
    for (int64_t i = 0; i < num_grids; ++i) {
        p3d::McGeom g;
        make_geom(&descs[i], &g);
        const Layout l = make_layout(g);
        const p3d::McWorkspace ws = bind(workspace, l);
        const p3d::McEmitParams prm = make_params(&descs[i], 0);
        P3D_CUDA(cudaMemsetAsync(workspace, 0, l.zero_end, s));
        p3d::launch_tile_pass(grids[i], dtype, g, ws, prm, vertices[i], vertex_capacities[i], 0, s);
        if (p3d::tile_pass_error()) return fail(P3D_ERR_CUDA, std::string("p3d_mc_extract_batch: ") + p3d::tile_pass_error());
        if (faces[i]) p3d::launch_faces(g, ws, prm, faces[i], face_capacities[i], false, s);
        P3D_CUDA(cudaMemcpyAsync(pin + 2 * i, &ws.header->total_v, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    }
    P3D_CUDA(cudaGetLastError());
    P3D_CUDA(cudaStreamSynchronize(s));
    for (int64_t i = 0; i < 2 * num_grids; ++i) counts_host[i] = pin[i];
    for (int64_t i = 0; i < num_grids; ++i)
        if (counts_host[2 * i] > INT32_MAX)
            return fail(P3D_ERR_OVERFLOW, "p3d_mc_extract_batch: vertex count exceeds the int32 face-index contract");
    return P3D_OK;
}

p3d_status p3d_mc_tile_async(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace, size_t workspace_bytes,
                             float *vertices, int64_t vertex_capacity, void *stream) {
    p3d::McGeom g;
    if (!make_geom(desc, &g)) return fail(P3D_ERR_INVALID, "p3d_mc_tile_async: invalid descriptor");
    if (dtype < P3D_F32 || dtype > P3D_U8) return fail(P3D_ERR_INVALID, "p3d_mc_tile_async: unknown dtype");
    if (!grid || !workspace) return fail(P3D_ERR_INVALID, "p3d_mc_tile_async: null pointer");
    if (desc->global_rx < 1) return fail(P3D_ERR_INVALID, "p3d_mc_tile_async: global_rx must be >= 1");
    if (vertex_capacity < 0 || (vertex_capacity > 0 && !vertices))
        return fail(P3D_ERR_INVALID, "p3d_mc_tile_async: vertex_capacity without a vertex buffer");
    const Layout l = make_layout(g);
    if (workspace_bytes < l.total) return fail(P3D_ERR_WORKSPACE, "p3d_mc_tile_async: workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace) % kAlign) return fail(P3D_ERR_INVALID, "p3d_mc_tile_async: workspace must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    P3D_CUDA(cudaMemsetAsync(workspace, 0, l.zero_end, s));
    p3d::launch_tile_pass(grid, dtype, g, bind(workspace, l), make_params(desc, 0), vertices, vertex_capacity, 0, s);
    if (p3d::tile_pass_error()) return fail(P3D_ERR_CUDA, std::string("p3d_mc_tile_async: ") + p3d::tile_pass_error());
    P3D_CUDA(cudaGetLastError());
    return P3D_OK;
}

int64_t p3d_mc_exchange_words(const p3d_mc_desc *desc) {
    const int64_t w = p3d_mc_plane_table_words(desc);
    return w ? w + 4 : 0;
}

p3d_status p3d_mc_export_exchange(const p3d_mc_desc *desc, const void *workspace, uint32_t *out, void *stream) {
    p3d::McGeom g;
    if (!make_geom(desc, &g) || !workspace || !out) return fail(P3D_ERR_INVALID, "p3d_mc_export_exchange: invalid argument");
    p3d::launch_export_exchange(out, g, bind(const_cast<void *>(workspace), make_layout(g)), static_cast<cudaStream_t>(stream));
    P3D_CUDA(cudaGetLastError());
    return P3D_OK;
}

p3d_status p3d_mc_faces_exchanged(const p3d_mc_desc *desc, void *workspace, const uint32_t *gathered, int rank, int world,
                                  int32_t *faces, int64_t face_capacity, void *stream) {
    p3d::McGeom g;
    if (!make_geom(desc, &g) || !workspace || !gathered) return fail(P3D_ERR_INVALID, "p3d_mc_faces_exchanged: invalid argument");
    if (world < 1 || rank < 0 || rank >= world) return fail(P3D_ERR_INVALID, "p3d_mc_faces_exchanged: bad rank / world");
    if ((rank + 1 < world) != (g.rx == g.owned_x + 1))
        return fail(P3D_ERR_INVALID, "p3d_mc_faces_exchanged: every shard but the last must carry a halo plane");
    if (face_capacity < 0 || (face_capacity > 0 && !faces)) return fail(P3D_ERR_INVALID, "p3d_mc_faces_exchanged: capacity without a buffer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const p3d::McWorkspace ws = bind(workspace, make_layout(g));
    p3d::launch_apply_exchange(g, ws, gathered, rank, world, s);
    if (faces) p3d::launch_faces(g, ws, make_params(desc, 0), faces, face_capacity, true, s);
    P3D_CUDA(cudaGetLastError());
    return P3D_OK;
}

namespace {
// ncclAllGather, resolved at run time: the library links against no NCCL (torch brings its own copy)
typedef int (*NcclAllGatherFn)(const void *, void *, size_t, int /* ncclDataType_t */, void * /* ncclComm_t */, cudaStream_t);
NcclAllGatherFn nccl_all_gather() {
    static NcclAllGatherFn fn = [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy the process already uses, if any
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        return h ? reinterpret_cast<NcclAllGatherFn>(dlsym(h, "ncclAllGather")) : nullptr;
    }();
    return fn;
}
}  // namespace

p3d_status p3d_mc_sharded_extract(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace, size_t workspace_bytes,
                                  void *nccl_comm, int rank, int world, uint32_t *exchange_send, uint32_t *exchange_recv,
                                  float *vertices, int64_t vertex_capacity, int32_t *faces, int64_t face_capacity,
                                  int64_t *counts_host, void *stream) {
    if (!nccl_comm || !exchange_send || !exchange_recv || !counts_host) return fail(P3D_ERR_INVALID, "p3d_mc_sharded_extract: null pointer");
    if (world < 1 || rank < 0 || rank >= world) return fail(P3D_ERR_INVALID, "p3d_mc_sharded_extract: bad rank / world");
    NcclAllGatherFn gather = nccl_all_gather();
    if (!gather) return fail(P3D_ERR_CUDA, "p3d_mc_sharded_extract: libnccl.so.2 (ncclAllGather) not found");
    const int64_t words = p3d_mc_exchange_words(desc);
    if (words <= 0) return fail(P3D_ERR_INVALID, "p3d_mc_sharded_extract: invalid descriptor");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    p3d_status st = p3d_mc_tile_async(desc, grid, dtype, workspace, workspace_bytes, vertices, vertex_capacity, stream);
    if (st != P3D_OK) return st;
    st = p3d_mc_export_exchange(desc, workspace, exchange_send, stream);
    if (st != P3D_OK) return st;
    const int rc = gather(exchange_send, exchange_recv, (size_t)words, 2 /* ncclInt32 */, nccl_comm, s);
    if (rc != 0) return fail(P3D_ERR_CUDA, "p3d_mc_sharded_extract: ncclAllGather failed with code " + std::to_string(rc));
    st = p3d_mc_faces_exchanged(desc, workspace, exchange_recv, rank, world, faces, face_capacity, stream);
    if (st != P3D_OK) return st;
    // the last four words of every shard's payload are its {V, F}: one strided copy, the only synchronisation
    int64_t *pin = pinned_counts((size_t)world);
    if (!pin) return fail(P3D_ERR_CUDA, "p3d_mc_sharded_extract: pinned allocation failed");
    P3D_CUDA(cudaMemcpy2DAsync(pin, 16, exchange_recv + (words - 4), (size_t)words * 4, 16, (size_t)world,
                               cudaMemcpyDeviceToHost, s));
    P3D_CUDA(cudaStreamSynchronize(s));
    int64_t total_v = 0;
    for (int i = 0; i < 2 * world; ++i) counts_host[i] = pin[i];
    for (int r = 0; r < world; ++r) total_v += counts_host[2 * r];
    if (total_v > INT32_MAX)
        return fail(P3D_ERR_OVERFLOW, "p3d_mc_sharded_extract: global vertex count exceeds the int32 face-index contract");
    return P3D_OK;
}

// ---- the same exchange over peer memory (mc_peer.cu): no collective between the two passes ----
struct p3d_mc_peer {
    int rank = 0, world = 1, device = 0;
    int64_t n = 0;                        // table entries of a plane
    size_t bytes = 0;
    char *base = nullptr;                 // my mailbox (cudaMalloc)
    char *peer[p3d::kMaxPeers] = {};      // every rank's mailbox as mapped here; peer[rank] = base
    bool connected = false;
    uint32_t epoch = 0;                   // calls made so far (the same on every rank: the call is collective)
};

size_t p3d_mc_peer_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

p3d_status p3d_mc_peer_create(const p3d_mc_desc *desc, int rank, int world, p3d_mc_peer **out, void *handle_out) {
    p3d::McGeom g;
    if (!make_geom(desc, &g) || !out || !handle_out) return fail(P3D_ERR_INVALID, "p3d_mc_peer_create: invalid argument");
    if (world < 1 || world > p3d::kMaxPeers || rank < 0 || rank >= world)
        return fail(P3D_ERR_INVALID, "p3d_mc_peer_create: bad rank / world (at most " + std::to_string(p3d::kMaxPeers) + " ranks)");
    p3d_mc_peer *p = new p3d_mc_peer();
    p->rank = rank, p->world = world;
    p->n = g.ry * (int64_t)g.np;
    p->bytes = (size_t)p3d::kPeerDataOffset + (size_t)2 * world * (size_t)(p->n + 1) * 16;
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&p->base), p->bytes);
    if (e == cudaSuccess) e = cudaMemset(p->base, 0, p->bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->base);
    if (e != cudaSuccess) {
        if (p->base) cudaFree(p->base);
        delete p;
        return fail(P3D_ERR_CUDA, std::string("p3d_mc_peer_create: ") + cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, sizeof h);
    p->peer[rank] = p->base;
    *out = p;
    return P3D_OK;
}

p3d_status p3d_mc_peer_connect(p3d_mc_peer *p, const void *handles) {
    if (!p || !handles) return fail(P3D_ERR_INVALID, "p3d_mc_peer_connect: null pointer");
    if (p->connected) return P3D_OK;
    const cudaIpcMemHandle_t *h = static_cast<const cudaIpcMemHandle_t *>(handles);
    for (int t = 0; t < p->world; ++t) {
        if (t == p->rank) continue;
        void *q = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&q, h[t], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return fail(P3D_ERR_CUDA, "p3d_mc_peer_connect: cudaIpcOpenMemHandle of rank " + std::to_string(t) + " failed: " + cudaGetErrorString(e));
        p->peer[t] = static_cast<char *>(q);
    }
    p->connected = true;
    return P3D_OK;
}

void p3d_mc_peer_disconnect(p3d_mc_peer *p) {
    if (!p) return;
    for (int t = 0; t < p->world; ++t)
        if (t != p->rank && p->peer[t]) {
            cudaIpcCloseMemHandle(p->peer[t]);
            p->peer[t] = nullptr;
        }
    p->connected = false;
}

void p3d_mc_peer_destroy(p3d_mc_peer *p) {
    if (!p) return;
    p3d_mc_peer_disconnect(p);
    if (p->base) cudaFree(p->base);
    delete p;
}

p3d_status p3d_mc_sharded_extract_p2p(const p3d_mc_desc *desc, const void *grid, int dtype, void *workspace, size_t workspace_bytes,
                                      p3d_mc_peer *peer, float *vertices, int64_t vertex_capacity, int32_t *faces,
                                      int64_t face_capacity, int64_t *counts_host, void *stream) {
    if (!peer || !counts_host) return fail(P3D_ERR_INVALID, "p3d_mc_sharded_extract_p2p: null pointer");
    if (!peer->connected && peer->world > 1) return fail(P3D_ERR_INVALID, "p3d_mc_sharded_extract_p2p: p3d_mc_peer_connect has not been called");
    p3d::McGeom g;
    if (!make_geom(desc, &g) || g.ry * (int64_t)g.np != peer->n)
        return fail(P3D_ERR_INVALID, "p3d_mc_sharded_extract_p2p: the descriptor's plane does not match the mailbox");
    const int rank = peer->rank, world = peer->world;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    p3d_status st = p3d_mc_tile_async(desc, grid, dtype, workspace, workspace_bytes, vertices, vertex_capacity, stream);
    if (st != P3D_OK) return st;
    const uint32_t epoch = ++peer->epoch;
    p3d::PeerParams pp;
    for (int t = 0; t < p3d::kMaxPeers; ++t) pp.peer[t] = peer->peer[t];
    pp.rank = rank, pp.world = world, pp.n = peer->n;
    p3d::launch_export_p2p(pp, bind(workspace, make_layout(g)), epoch, s);
    p3d::launch_wait_p2p(peer->base, world, epoch, s);
    P3D_CUDA(cudaGetLastError());
    const int64_t words = 4 * (peer->n + 1);
    const uint32_t *recv = reinterpret_cast<const uint32_t *>(peer->base + p3d::kPeerDataOffset) + (size_t)(epoch & 1u) * world * words;
    st = p3d_mc_faces_exchanged(desc, workspace, recv, rank, world, faces, face_capacity, stream);
    if (st != P3D_OK) return st;
    // every shard's {V, F} (the last four words of its slot) and the timeout word: one wait
    int64_t *pin = pinned_counts((size_t)world + 1);
    if (!pin) return fail(P3D_ERR_CUDA, "p3d_mc_sharded_extract_p2p: pinned allocation failed");
    P3D_CUDA(cudaMemcpy2DAsync(pin, 16, recv + (words - 4), (size_t)words * 4, 16, (size_t)world, cudaMemcpyDeviceToHost, s));
    P3D_CUDA(cudaMemcpyAsync(pin + 2 * world, reinterpret_cast<const unsigned int *>(peer->base) + p3d::kPeerTimeoutWord, 4,
                             cudaMemcpyDeviceToHost, s));
    P3D_CUDA(cudaStreamSynchronize(s));
    if (*reinterpret_cast<const uint32_t *>(pin + 2 * world) != 0u)
        return fail(P3D_ERR_CUDA, "p3d_mc_sharded_extract_p2p: a peer did not deliver its payload (timeout in the exchange)");
    int64_t total_v = 0;
    for (int i = 0; i < 2 * world; ++i) counts_host[i] = pin[i];
    for (int r = 0; r < world; ++r) total_v += counts_host[2 * r];
    if (total_v > INT32_MAX)
        return fail(P3D_ERR_OVERFLOW, "p3d_mc_sharded_extract_p2p: global vertex count exceeds the int32 face-index contract");
    return P3D_OK;
}

p3d_status p3d_mc_debug_stage(const p3d_mc_desc *desc, const float *grid, void *workspace, int stage, float *vertices,
                              int64_t vertex_capacity, void *stream) {
    p3d::McGeom g;
    if (!make_geom(desc, &g) || !grid || !workspace) return fail(P3D_ERR_INVALID, "p3d_mc_debug_stage: invalid argument");
    const Layout l = make_layout(g);
    const p3d::McWorkspace ws = bind(workspace, l);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (stage) {
        case 0: P3D_CUDA(cudaMemsetAsync(workspace, 0, l.zero_end, s)); break;
        case 1:
            p3d::launch_tile_pass(grid, P3D_F32, g, ws, make_params(desc, 0), vertices, vertex_capacity, 0, s);
            if (p3d::tile_pass_error()) return fail(P3D_ERR_CUDA, p3d::tile_pass_error());
            break;
        default: return fail(P3D_ERR_INVALID, "p3d_mc_debug_stage: stage must be 0 or 1");
    }
    P3D_CUDA(cudaGetLastError());
    return P3D_OK;
}

p3d_status p3d_mc_export_first_plane(const p3d_mc_desc *desc, const void *workspace, uint32_t *table_out, void *stream) {
    p3d::McGeom g;
    if (!make_geom(desc, &g) || !workspace || !table_out) return fail(P3D_ERR_INVALID, "p3d_mc_export_first_plane: invalid argument");
    const Layout l = make_layout(g);
    const p3d::McWorkspace ws = bind(const_cast<void *>(workspace), l);
    p3d::launch_export_plane(table_out, g, ws, static_cast<cudaStream_t>(stream));
    P3D_CUDA(cudaGetLastError());
    return P3D_OK;
}

p3d_status p3d_mc_import_halo_plane(const p3d_mc_desc *desc, void *workspace, const uint32_t *table_in, int64_t delta,
                                    void *stream) {
    p3d::McGeom g;
    if (!make_geom(desc, &g) || !workspace || !table_in) return fail(P3D_ERR_INVALID, "p3d_mc_import_halo_plane: invalid argument");
    if (g.rx != g.owned_x + 1) return fail(P3D_ERR_INVALID, "p3d_mc_import_halo_plane: descriptor has no halo plane");
    if (delta < 0 || delta > INT32_MAX) return fail(P3D_ERR_OVERFLOW, "p3d_mc_import_halo_plane: delta out of range");
    const Layout l = make_layout(g);
    const p3d::McWorkspace ws = bind(workspace, l);
    p3d::launch_import_halo(g, ws, table_in, static_cast<uint32_t>(delta), static_cast<cudaStream_t>(stream));
    P3D_CUDA(cudaGetLastError());
    return P3D_OK;
}

int64_t p3d_mc_vertex_capacity_hint(const p3d_mc_desc *desc) {
    p3d::McGeom g;
    if (!make_geom(desc, &g)) return 0;
    // a closed surface crosses O(N^2) of the N^3 edges; 1/16 vertex per owned sample covers smooth fields
    // (gyroid 1024^3: 0.038), capped by the 3-per-sample maximum
    const int64_t samples = g.owned_x * g.ry * g.rz;
    int64_t cap = samples / 16 + 4096;
    if (cap > samples * 3) cap = samples * 3;
    if (cap > INT32_MAX) cap = INT32_MAX;
    return cap;
}

p3d_status p3d_mc_run(const p3d_mc_desc *desc, const float *grid, p3d_alloc_fn alloc, void *alloc_ctx, float **vertices,
                      int32_t **faces, int64_t *num_vertices, int64_t *num_faces, void *stream) {
    if (!alloc || !vertices || !faces || !num_vertices || !num_faces) return fail(P3D_ERR_INVALID, "p3d_mc_run: null pointer");
    const size_t bytes = p3d_mc_workspace_bytes(desc);
    if (!bytes) return fail(P3D_ERR_INVALID, "p3d_mc_run: invalid descriptor");
    void *ws = alloc(alloc_ctx, bytes);
    if (!ws) return fail(P3D_ERR_CUDA, "p3d_mc_run: workspace allocation failed");
    const int64_t cap = p3d_mc_vertex_capacity_hint(desc);
    float *spec = static_cast<float *>(alloc(alloc_ctx, (size_t)(cap > 0 ? cap : 1) * 12));
    if (!spec) return fail(P3D_ERR_CUDA, "p3d_mc_run: vertex allocation failed");
    int64_t counts[2] = {0, 0};
    if (p3d_mc_single_launch(desc, P3D_F32)) {  // small grid: count, then both outputs into exact buffers, one launch each
        p3d_status s1 = p3d_mc_extract(desc, grid, P3D_F32, ws, bytes, nullptr, 0, nullptr, 0, counts, stream);
        if (s1 != P3D_OK) return s1;
        *num_vertices = counts[0], *num_faces = counts[1];
        *vertices = static_cast<float *>(alloc(alloc_ctx, (size_t)(counts[0] > 0 ? counts[0] : 1) * 12));
        *faces = static_cast<int32_t *>(alloc(alloc_ctx, (size_t)(counts[1] > 0 ? counts[1] : 1) * 12));
        if (!*vertices || !*faces) return fail(P3D_ERR_CUDA, "p3d_mc_run: output allocation failed");
        if (counts[0] == 0) return P3D_OK;
        return p3d_mc_extract(desc, grid, P3D_F32, ws, bytes, *vertices, counts[0], *faces, counts[1], counts, stream);
    }
    p3d_status st = p3d_mc_count(desc, grid, ws, bytes, spec, cap, counts, stream);
    if (st != P3D_OK) return st;
    *num_vertices = counts[0];
    *num_faces = counts[1];
    *vertices = spec;
    if (counts[0] > cap) {  // the speculative buffer was too small: exact buffer, vertices-only second pass
        *vertices = static_cast<float *>(alloc(alloc_ctx, (size_t)counts[0] * 12));
        if (!*vertices) return fail(P3D_ERR_CUDA, "p3d_mc_run: output allocation failed");
        st = p3d_mc_vertices(desc, grid, ws, *vertices, counts[0], stream);
        if (st != P3D_OK) return st;
    }
    *faces = static_cast<int32_t *>(alloc(alloc_ctx, (size_t)(counts[1] > 0 ? counts[1] : 1) * 12));
    if (!*faces) return fail(P3D_ERR_CUDA, "p3d_mc_run: output allocation failed");
    return p3d_mc_faces(desc, ws, *faces, 0, stream);
}

}  // extern "C"
