// primitive3d_b200/csrc/mc_host_stream.cu -- p3d_mc_extract_host: marching cubes of a grid that lives in HOST
// memory, pipelined slab by slab.
//
// The reference's wrapper copies the whole grid to the device, extracts, and leaves the copy back to the caller
// (prim3d/utility/marching_cubes.py:86-95): upload, compute and download run one after the other, and the grid has
// to fit in device memory next to the outputs.  Here the grid is cut into dim-0 slabs (the multi-GPU decomposition
// of sharded.py on ONE device): while slab k is extracted, slab k+1 is uploading and the mesh of slab k-1 is
// downloading, so the call costs about max(upload, download) instead of their sum, and the device holds two slabs
// at a time instead of the grid.  Slab k's last-plane cells need slab k+1's vertex numbering, hence the order
//   tile pass k+1  ->  face pass k          (p3d_mc_tile_async / p3d_mc_export_exchange / p3d_mc_faces_exchanged)
// and the vertex-id base of a slab is summed on the device from the exchange payloads of the slabs before it.
// Faces come out in voxel-major order with global vertex ids, vertices slab by slab -- the same arrays the
// multi-GPU driver produces with world = number of slabs.
#include "../../include/prim3d_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <string>
#include <vector>

#include "p3d_error.h"

namespace {

size_t elem_bytes(int dtype) {
    switch (dtype) {
        case P3D_F32: case P3D_I32: return 4;
        case P3D_F16: case P3D_BF16: case P3D_I16: return 2;
        case P3D_F64: case P3D_I64: return 8;
        case P3D_U8: return 1;
        default: return 0;
    }
}

constexpr int kMaxSlabs = 4096;

// Streams and the pinned landing pad for the per-slab counts are kept per host thread and device: creating them
// costs more than a slab.
struct ThreadState {
    int device = -1;
    cudaStream_t compute = nullptr, up = nullptr, down = nullptr;
    int64_t *pinned = nullptr;
};
cudaError_t thread_state(ThreadState **out) {
    thread_local ThreadState st;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (st.device != dev) {
        if (st.device >= 0) {  // the thread moved to another device: start over (rare)
            cudaStreamDestroy(st.compute), cudaStreamDestroy(st.up), cudaStreamDestroy(st.down);
            st.compute = st.up = st.down = nullptr;
        }
        if ((e = cudaStreamCreateWithFlags(&st.compute, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&st.up, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&st.down, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if (!st.pinned && (e = cudaHostAlloc(reinterpret_cast<void **>(&st.pinned), kMaxSlabs * 2 * sizeof(int64_t),
                                             cudaHostAllocPortable)) != cudaSuccess)
            return e;
        st.device = dev;
    }
    *out = &st;
    return cudaSuccess;
}

struct Pipeline {  // what one call owns: events, device memory it had to allocate itself; released on every exit path
    ThreadState *t = nullptr;
    cudaStream_t compute = nullptr, up = nullptr, down = nullptr;
    int64_t *pinned = nullptr;
    std::vector<void *> device;
    std::vector<cudaEvent_t> events;
    ~Pipeline() {
        for (cudaStream_t s : {compute, up, down})
            if (s) cudaStreamSynchronize(s);
        for (void *p : device) cudaFree(p);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
    cudaError_t alloc(void **p, size_t bytes) {
        const cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 256));
        if (e == cudaSuccess) device.push_back(*p);
        return e;
    }
    cudaError_t event(cudaEvent_t *e) {
        const cudaError_t r = cudaEventCreateWithFlags(e, cudaEventDisableTiming);
        if (r == cudaSuccess) events.push_back(*e);
        return r;
    }
};

// slabs of the pipeline and the device arena they need: two slab buffers, two workspaces, two vertex and two face
// buffers of per-slab speculative capacity, the exchange payloads of every slab
struct Plan {
    int64_t slab_planes = 0, words = 0, slab_vcap = 0, slab_fcap = 0;
    int nslabs = 0;
    size_t ws_bytes = 0, grid_bytes = 0, off_grid[2], off_ws[2], off_verts[2], off_faces[2], off_payload = 0, total = 0;
};

p3d_mc_desc slab_desc(const p3d_mc_desc &full, int64_t slab_planes, int k) {
    p3d_mc_desc d = full;
    const int64_t x0 = (int64_t)k * slab_planes, x1 = std::min(full.rx, x0 + slab_planes);
    d.owned_x = x1 - x0;
    d.rx = d.owned_x + (x1 < full.rx ? 1 : 0);
    d.x_origin = x0;
    return d;
}

bool make_plan(const p3d_mc_desc *desc, int dtype, int64_t slab_planes, Plan *pl) {
    const size_t eb = elem_bytes(dtype);
    if (!desc || !eb || desc->rx < 1 || desc->ry < 1 || desc->rz < 1) return false;
    const int64_t rx = desc->rx, plane = desc->ry * desc->rz;
    if (slab_planes <= 0) {  // ~16 slabs, whole 8-plane tile blocks, at least 64 MB of samples each
        slab_planes = (rx + 15) / 16;
        const int64_t min_planes = ((int64_t)64 << 20) / std::max<int64_t>(plane * (int64_t)eb, 1) + 1;
        slab_planes = std::max(slab_planes, min_planes);
    }
    slab_planes = std::min<int64_t>(std::max<int64_t>((slab_planes + 7) / 8 * 8, 8), rx);
    while ((rx + slab_planes - 1) / slab_planes > kMaxSlabs) slab_planes *= 2;
    pl->slab_planes = slab_planes;
    pl->nslabs = (int)((rx + slab_planes - 1) / slab_planes);
    const p3d_mc_desc d0 = slab_desc(*desc, slab_planes, 0);  // the first slab is the largest
    pl->ws_bytes = p3d_mc_workspace_bytes(&d0);
    pl->words = p3d_mc_exchange_words(&d0);
    if (!pl->ws_bytes || !pl->words) return false;
    pl->slab_vcap = p3d_mc_vertex_capacity_hint(&d0);
    pl->slab_fcap = 2 * pl->slab_vcap;
    pl->grid_bytes = (size_t)d0.rx * plane * eb;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) / 256 * 256;
        return at;
    };
    for (int i = 0; i < 2; ++i) {
        pl->off_grid[i] = take(pl->grid_bytes);
        pl->off_ws[i] = take(pl->ws_bytes);
        pl->off_verts[i] = take((size_t)pl->slab_vcap * 12);
        pl->off_faces[i] = take((size_t)pl->slab_fcap * 12);
    }
    pl->off_payload = take((size_t)pl->nslabs * pl->words * 4);
    pl->total = off;
    return true;
}

#define HS_CUDA(expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t e_ = (expr);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return p3d::set_error(P3D_ERR_CUDA, std::string("p3d_mc_extract_host: " #expr " failed: ") + cudaGetErrorString(e_)); \
    } while (0)
#define HS_P3D(expr)                   \
    do {                               \
        p3d_status s_ = (expr);        \
        if (s_ != P3D_OK) return s_;   \
    } while (0)

}  // namespace

extern "C" size_t p3d_mc_extract_host_arena_bytes(const p3d_mc_desc *desc, int dtype, int64_t slab_planes) {
    Plan pl;
    return make_plan(desc, dtype, slab_planes, &pl) ? pl.total : 0;
}

extern "C" p3d_status p3d_mc_extract_host(const p3d_mc_desc *desc, const void *host_grid, int dtype, int64_t slab_planes,
                                          float *host_vertices, int64_t vertex_capacity, int32_t *host_faces,
                                          int64_t face_capacity, int64_t *counts_host, void *device_arena,
                                          size_t arena_bytes) {
    if (!desc || !host_grid || !counts_host) return p3d::set_error(P3D_ERR_INVALID, "p3d_mc_extract_host: null pointer");
    const size_t eb = elem_bytes(dtype);
    if (!eb) return p3d::set_error(P3D_ERR_INVALID, "p3d_mc_extract_host: unknown dtype");
    if (desc->rx < 1 || desc->ry < 1 || desc->rz < 1 || desc->owned_x != desc->rx || desc->global_rx != desc->rx || desc->x_origin != 0)
        return p3d::set_error(P3D_ERR_INVALID, "p3d_mc_extract_host: the descriptor must describe the whole grid");
    if (vertex_capacity < 0 || face_capacity < 0 || (vertex_capacity && !host_vertices) || (face_capacity && !host_faces))
        return p3d::set_error(P3D_ERR_INVALID, "p3d_mc_extract_host: capacity without a buffer");

    Plan pl;
    if (!make_plan(desc, dtype, slab_planes, &pl)) return p3d::set_error(P3D_ERR_INVALID, "p3d_mc_extract_host: invalid grid shape");
    const int64_t plane = desc->ry * desc->rz;
    slab_planes = pl.slab_planes;
    const int nslabs = pl.nslabs;
    const size_t ws_bytes = pl.ws_bytes;
    const int64_t words = pl.words, slab_vcap = pl.slab_vcap, slab_fcap = pl.slab_fcap;
    auto slab_desc = [&](int k) { return ::slab_desc(*desc, slab_planes, k); };

    Pipeline P;
    HS_CUDA(thread_state(&P.t));
    P.compute = P.t->compute, P.up = P.t->up, P.down = P.t->down, P.pinned = P.t->pinned;
    // the caller's arena (e.g. from a caching allocator) if it is large enough, else our own allocation
    char *arena = static_cast<char *>(device_arena);
    if (!arena || arena_bytes < pl.total || reinterpret_cast<uintptr_t>(arena) % 256) {
        void *own = nullptr;
        HS_CUDA(P.alloc(&own, pl.total));
        arena = static_cast<char *>(own);
    }
    void *grid_d[2], *ws_d[2], *verts_d[2], *faces_d[2], *payload_d = arena + pl.off_payload;
    for (int i = 0; i < 2; ++i) {
        grid_d[i] = arena + pl.off_grid[i];
        ws_d[i] = arena + pl.off_ws[i];
        verts_d[i] = arena + pl.off_verts[i];
        faces_d[i] = arena + pl.off_faces[i];
    }
    std::vector<cudaEvent_t> up_done(nslabs), tile_done(nslabs), faces_done(nslabs), verts_down(nslabs), faces_down(nslabs);
    for (int k = 0; k < nslabs; ++k) {
        HS_CUDA(P.event(&up_done[k]));
        HS_CUDA(P.event(&tile_done[k]));
        HS_CUDA(P.event(&faces_done[k]));
        HS_CUDA(P.event(&verts_down[k]));
        HS_CUDA(P.event(&faces_down[k]));
    }

    const char *src = static_cast<const char *>(host_grid);
    auto upload = [&](int k) -> cudaError_t {  // slab k -> grid_d[k & 1], once tile pass k - 2 has read that buffer
        const p3d_mc_desc d = slab_desc(k);
        if (k >= 2) {
            cudaError_t e = cudaStreamWaitEvent(P.up, tile_done[k - 2], 0);
            if (e != cudaSuccess) return e;
        }
        cudaError_t e = cudaMemcpyAsync(grid_d[k & 1], src + (size_t)d.x_origin * plane * eb, (size_t)d.rx * plane * eb,
                                        cudaMemcpyHostToDevice, P.up);
        return e != cudaSuccess ? e : cudaEventRecord(up_done[k], P.up);
    };

    int64_t v_total = 0, f_total = 0;
    std::vector<int64_t> v_of(nslabs, 0), f_of(nslabs, 0), v_off(nslabs, 0), f_off(nslabs, 0);
    bool fits = true;  // false once an output buffer is too small: counting goes on, copying stops

    // face pass of slab j (the exchange payloads of slabs 0 .. j + 1 are on the device) and its download
    auto faces_of = [&](int j) -> p3d_status {
        const p3d_mc_desc d = slab_desc(j);
        if (j >= 2) HS_CUDA(cudaStreamWaitEvent(P.compute, faces_down[j - 2], 0));  // faces_d[j & 1] has been downloaded
        int32_t *out = static_cast<int32_t *>(faces_d[j & 1]);
        void *exact = nullptr;
        if (f_of[j] > slab_fcap) {  // guess too small for this slab: exact buffer for this one pass
            HS_CUDA(P.alloc(&exact, (size_t)f_of[j] * 12));
            out = static_cast<int32_t *>(exact);
        }
        HS_P3D(p3d_mc_faces_exchanged(&d, ws_d[j & 1], static_cast<const uint32_t *>(payload_d), j, nslabs, out,
                                      std::max(f_of[j], slab_fcap), P.compute));
        HS_CUDA(cudaEventRecord(faces_done[j], P.compute));
        if (fits && f_of[j] > 0) {
            HS_CUDA(cudaStreamWaitEvent(P.down, faces_done[j], 0));
            HS_CUDA(cudaMemcpyAsync(host_faces + 3 * f_off[j], out, (size_t)f_of[j] * 12, cudaMemcpyDeviceToHost, P.down));
        }
        HS_CUDA(cudaEventRecord(faces_down[j], P.down));
        return P3D_OK;
    };

    HS_CUDA(upload(0));
    for (int k = 0; k < nslabs; ++k) {
        const p3d_mc_desc d = slab_desc(k);
        if (k + 1 < nslabs) HS_CUDA(upload(k + 1));
        HS_CUDA(cudaStreamWaitEvent(P.compute, up_done[k], 0));
        if (k >= 2) HS_CUDA(cudaStreamWaitEvent(P.compute, verts_down[k - 2], 0));  // verts_d[k & 1] has been downloaded
        HS_P3D(p3d_mc_tile_async(&d, grid_d[k & 1], dtype, ws_d[k & 1], ws_bytes, static_cast<float *>(verts_d[k & 1]),
                                 slab_vcap, P.compute));
        uint32_t *payload = static_cast<uint32_t *>(payload_d) + (size_t)k * words;
        HS_P3D(p3d_mc_export_exchange(&d, ws_d[k & 1], payload, P.compute));
        HS_CUDA(cudaMemcpyAsync(P.pinned + 2 * k, payload + (words - 4), 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, P.compute));
        HS_CUDA(cudaEventRecord(tile_done[k], P.compute));
        if (k >= 1) HS_P3D(faces_of(k - 1));  // queued behind tile pass k: runs while the host waits below

        HS_CUDA(cudaEventSynchronize(tile_done[k]));  // {V_k, F_k}: sizes of this slab's downloads
        v_of[k] = P.pinned[2 * k], f_of[k] = P.pinned[2 * k + 1];
        v_off[k] = v_total, f_off[k] = f_total;
        v_total += v_of[k], f_total += f_of[k];
        if (v_total > INT32_MAX) return p3d::set_error(P3D_ERR_OVERFLOW, "p3d_mc_extract_host: vertex count exceeds the int32 face-index contract");
        if (v_total > vertex_capacity || f_total > face_capacity) fits = false;
        if (fits && v_of[k] > 0) {
            const float *vsrc = static_cast<const float *>(verts_d[k & 1]);
            if (v_of[k] > slab_vcap) {  // guess too small for this slab: exact buffer, vertices-only pass
                void *exact = nullptr;
                HS_CUDA(P.alloc(&exact, (size_t)v_of[k] * 12));
                HS_P3D(p3d_mc_vertices_typed(&d, grid_d[k & 1], dtype, ws_d[k & 1], static_cast<float *>(exact), v_of[k], P.compute));
                HS_CUDA(cudaEventRecord(tile_done[k], P.compute));  // the slab buffer is read again: uploads wait for this
                vsrc = static_cast<const float *>(exact);
            }
            HS_CUDA(cudaStreamWaitEvent(P.down, tile_done[k], 0));
            HS_CUDA(cudaMemcpyAsync(host_vertices + 3 * v_off[k], vsrc, (size_t)v_of[k] * 12, cudaMemcpyDeviceToHost, P.down));
        }
        HS_CUDA(cudaEventRecord(verts_down[k], P.down));
    }
    HS_P3D(faces_of(nslabs - 1));
    HS_CUDA(cudaStreamSynchronize(P.compute));
    HS_CUDA(cudaStreamSynchronize(P.down));
    counts_host[0] = v_total;
    counts_host[1] = f_total;
    return P3D_OK;
}
