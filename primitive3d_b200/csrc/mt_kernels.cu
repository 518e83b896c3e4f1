// primitive3d_b200/csrc/mt_kernels.cu -- sm_100a kernels + C ABI of the marching-tetrahedra path.
//
// Replaces /root/reference/prim3d/utility/marching_tetrahedras.py:89-235, which is a chain of ~25
// ATen launches (gathers, torch.det = batched LU, torch.unique(dim=0) = sort of 6*T_valid int64
// pairs, masked scatters).  Here:
//
//   k_mt_classify   one streaming pass over the tets (32 B each): orientation sign from the
//                   triple product (p1-p0).((p2-p0)x(p3-p0)) in fp64 (:50-65; mt_common.cuh says why not the
//                   reference's float32 torch.det), in-place swap of
//                   columns 0/1 of negatively oriented tets (:148), occupancy code sdf>0 (:151-154)
//                   -> 1 byte per tet, plus the totals the host needs to size the outputs.
//   k_mt_compact    reads the code bytes only; a decoupled look-back scan places every valid tet
//                   in the reference's face order (all one-triangle tets, then all two-triangle
//                   tets, each group in tet order, :205-223) and appends the crossing edges of
//                   valid tets as 64-bit keys (min<<32 | max).
//   radix sort      hand-written stable LSD radix sort (8-bit digits, only the significant bits of
//                   the point ids), then k_mt_unique (look-back scan) de-duplicates: the i-th
//                   unique key is vertex i -- the order torch.unique(dim=0) gives (:157-173).
//   k_mt_verts      p0*w0 + p1*w1, w = (-s1, s0)/(s0 + (-s1)), separately rounded fp32 (:177-189).
//   k_mt_faces      per face slot: table row of the tet's code, vertex id of each edge by binary
//                   search in the sorted unique keys (:193-223); tet_idx (:225-234).
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "mt_common.cuh"
#include "p3d_error.h"
#include "scan_utils.cuh"

namespace p3d {
namespace {

constexpr int kThreads = 256;
constexpr int kTetsPerThread = 16;
constexpr int kTileTets = kThreads * kTetsPerThread;  // compaction tile
constexpr int kSortItems = 16;
constexpr int kSortTile = kThreads * kSortItems;      // radix-sort tile (4096 keys)
constexpr int kUniqTile = kThreads * 8;

struct ClassifyCounters {  // lives in the 64 bytes after codes[T]
    unsigned long long n1, n2, ne;
    unsigned long long bad;  // tets naming a point outside [0, P): the reference's indexing asserts on those
};

struct MtHeader {
    unsigned int ticket_compact;
    unsigned int ticket_unique;
    unsigned long long key_cursor;
    unsigned long long num_unique;
};

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_mt_classify(const float *__restrict__ pts, int64_t P, int64_t *tets, int64_t T,
                                                          const float *__restrict__ sdf, uint8_t *__restrict__ codes,
                                                          ClassifyCounters *counters, int oriented) {
    unsigned long long n1 = 0, n2 = 0;
    longlong2 *t2 = reinterpret_cast<longlong2 *>(tets);
    auto one = [&](int64_t t, longlong2 lo, longlong2 hi) {
        int64_t i0 = lo.x, i1 = lo.y;
        const int64_t i2 = hi.x, i3 = hi.y;
        // an index outside [0, P) would read out of bounds (the reference's torch indexing raises a device-side
        // assert): the tet is dropped and the call fails with P3D_ERR_INVALID
        if ((uint64_t)i0 >= (uint64_t)P || (uint64_t)i1 >= (uint64_t)P || (uint64_t)i2 >= (uint64_t)P || (uint64_t)i3 >= (uint64_t)P) {
            codes[t] = 0;
            atomicAdd(&counters->bad, 1ull);
            return;
        }
        const double det = oriented ? 0.0 : orient_det_f64(pts, i0, i1, i2, i3);  // marching_tetrahedras.py:61-64
        if (det < 0.0) {  // :148 tets[flip, :2] = tets[flip][:, [1, 0]]
            const int64_t s = i0;
            i0 = i1;
            i1 = s;
            t2[2 * t] = make_longlong2(i0, i1);
        }
        const uint32_t code = (__ldg(sdf + i0) > 0.f ? 1u : 0u) | (__ldg(sdf + i1) > 0.f ? 2u : 0u) |
                              (__ldg(sdf + i2) > 0.f ? 4u : 0u) | (__ldg(sdf + i3) > 0.f ? 8u : 0u);
        codes[t] = (uint8_t)code;
        const uint32_t nt = num_tri(code);
        n1 += nt == 1u;
        n2 += nt == 2u;
    };
    // two tets per iteration, their index loads issued together: the dependent gathers of one overlap the other's
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; t + stride < T; t += 2 * stride) {
        const longlong2 lo0 = t2[2 * t], hi0 = t2[2 * t + 1];
        const longlong2 lo1 = t2[2 * (t + stride)], hi1 = t2[2 * (t + stride) + 1];
        one(t, lo0, hi0);
        one(t + stride, lo1, hi1);
    }
    if (t < T) one(t, t2[2 * t], t2[2 * t + 1]);
    n1 = warp_sum64(n1);
    n2 = warp_sum64(n2);
    __shared__ unsigned long long s1[kThreads / 32], s2[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s1[warp] = n1;
        s2[warp] = n2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a = 0, b = 0;
        for (int w = 0; w < kThreads / 32; ++w) {
            a += s1[w];
            b += s2[w];
        }
        if (a) atomicAdd(&counters->n1, a);
        if (b) atomicAdd(&counters->n2, b);
        if (a | b) atomicAdd(&counters->ne, 3 * a + 4 * b);  // 3 / 4 crossing edges per 1- / 2-triangle tet
    }
}

// Block-wide exclusive scan of one 64-bit value per thread (kThreads threads).
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long *s_warp,
                                                              unsigned long long *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long incl = warp_incl_scan64(v, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long base = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const unsigned long long t = s_warp[w];
        if (w < warp) base += t;
        sum += t;
    }
    __syncthreads();
    *total = sum;
    return base + incl - v;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_mt_compact(const int64_t *__restrict__ tets, int64_t T,
                                                         const uint8_t *__restrict__ codes, int64_t num_tiles,
                                                         unsigned long long n1_total, MtHeader *hdr,
                                                         unsigned long long *status, uint32_t *__restrict__ slot_tet,
                                                         uint64_t *__restrict__ keys) {
    __shared__ unsigned long long s_warp[kThreads / 32];
    __shared__ unsigned long long s_excl, s_keybase;
    __shared__ unsigned int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&hdr->ticket_compact, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= num_tiles) break;
        const int64_t t0 = tile * kTileTets + (int64_t)threadIdx.x * kTetsPerThread;
        uint32_t code[kTetsPerThread];
        unsigned long long mine = 0;  // n1 in bits 0..30, n2 in bits 31..61
        uint32_t nkeys = 0;
        static_assert(kTetsPerThread == 16, "one 16-byte load of codes per thread");
        if (t0 + kTetsPerThread <= T) {
            const uint4 c4 = *reinterpret_cast<const uint4 *>(codes + t0);  // t0 is a multiple of 16
            const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int j = 0; j < kTetsPerThread; ++j) code[j] = (cw[j >> 2] >> (8 * (j & 3))) & 255u;
        } else {
#pragma unroll
            for (int j = 0; j < kTetsPerThread; ++j) code[j] = (t0 + j < T) ? codes[t0 + j] : 0u;
        }
#pragma unroll
        for (int j = 0; j < kTetsPerThread; ++j) {
            const uint32_t nt = num_tri(code[j]);
            mine += (nt == 1u ? 1ull : 0ull) + (nt == 2u ? (1ull << 31) : 0ull);
            nkeys += nt == 1u ? 3u : (nt == 2u ? 4u : 0u);
        }
        unsigned long long agg, kagg;
        const unsigned long long excl = block_excl_scan(mine, s_warp, &agg);
        const unsigned long long kexcl = block_excl_scan(nkeys, s_warp, &kagg);
        if (warp == 0) {
            const unsigned long long e = lookback(status, tile, agg, lane);
            if (lane == 0) s_excl = e;
        } else if (threadIdx.x == 32) {
            s_keybase = kagg ? atomicAdd(&hdr->key_cursor, kagg) : 0ull;
        }
        __syncthreads();
        const unsigned long long pos = s_excl + excl;
        uint32_t p1 = (uint32_t)(pos & 0x7fffffffull);                  // rank among one-triangle tets
        uint32_t p2 = (uint32_t)(pos >> 31);                            // rank among two-triangle tets
        uint64_t *kout = keys + s_keybase + kexcl;
#pragma unroll
        for (int j = 0; j < kTetsPerThread; ++j) {
            const uint32_t nt = num_tri(code[j]);
            if (nt == 0u) continue;
            const int64_t t = t0 + j;
            if (nt == 1u) slot_tet[p1++] = (uint32_t)t;
            else slot_tet[n1_total + p2++] = (uint32_t)t;
            const longlong2 lo = reinterpret_cast<const longlong2 *>(tets)[2 * t];
            const longlong2 hi = reinterpret_cast<const longlong2 *>(tets)[2 * t + 1];
            const int64_t id[4] = {lo.x, lo.y, hi.x, hi.y};
#pragma unroll
            for (int e = 0; e < 6; ++e) {
                const int a = (kEdgeA >> (2 * e)) & 3, b = (kEdgeB >> (2 * e)) & 3;
                if (((code[j] >> a) ^ (code[j] >> b)) & 1u) *kout++ = make_key(id[a], id[b]);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort pass over 64-bit keys, 8-bit digit at `shift`.
// ghist is digit-major: ghist[d * nblocks + block].
__device__ __forceinline__ void sort_tile_counts(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                 uint32_t (*s_cnt)[256]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (kThreads / 32) * 256; i += kThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (kSortItems * 32);
    for (int it = 0; it < kSortItems; ++it) {
        const int64_t idx = base + it * 32 + lane;
        if (idx < n) atomicAdd(&s_cnt[warp][(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads) k_sort_hist(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                        uint32_t *__restrict__ ghist, int nblocks) {
    __shared__ uint32_t s_cnt[kThreads / 32][256];
    sort_tile_counts(keys, n, shift, s_cnt);
    uint32_t total = 0;
    for (int w = 0; w < kThreads / 32; ++w) total += s_cnt[w][threadIdx.x];
    ghist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = total;
}

// Single-CTA exclusive scan (the histogram is small: 256 * nblocks entries).
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t *data, int64_t n) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < n ? data[i] : 0u;
        const uint32_t incl = warp_incl_scan(v, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_warp[lane];
            const uint32_t wi = warp_incl_scan(w, lane);
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (i < n) data[i] = carry + s_warp[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + incl;
        __syncthreads();
    }
}

template <bool FUSED_SCAN>
__global__ void __launch_bounds__(kThreads) k_sort_scatter(const uint64_t *__restrict__ in, uint64_t *__restrict__ out,
                                                           int64_t n, int shift, const uint32_t *__restrict__ ghist,
                                                           int nblocks) {
    __shared__ uint32_t s_cnt[kThreads / 32][256];
    __shared__ uint32_t s_digit[kThreads / 32];
    sort_tile_counts(in, n, shift, s_cnt);
    {  // per-warp output bases: global digit base of this tile + the lower warps' counts
        const int d = threadIdx.x;
        uint32_t run;
        if (FUSED_SCAN) {
            // few tiles: every CTA derives its offsets from the raw per-tile counts itself (keys of smaller digits +
            // keys of my digit in earlier tiles) instead of a single-CTA scan kernel between the two passes
            const uint32_t *row = ghist + (int64_t)d * nblocks;
            uint32_t before = 0, total = 0;
#pragma unroll 4
            for (int b = 0; b < nblocks; ++b) {
                const uint32_t c = row[b];
                total += c;
                if (b < (int)blockIdx.x) before += c;
            }
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const uint32_t incl = warp_incl_scan(total, lane);
            if (lane == 31) s_digit[warp] = incl;
            __syncthreads();
            uint32_t lower = 0;
            for (int w = 0; w < warp; ++w) lower += s_digit[w];
            run = lower + incl - total + before;
        } else {
            run = ghist[(int64_t)d * nblocks + blockIdx.x];
        }
        for (int w = 0; w < kThreads / 32; ++w) {
            const uint32_t c = s_cnt[w][d];
            s_cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t base = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * (kSortItems * 32);
    for (int it = 0; it < kSortItems; ++it) {
        const int64_t idx = base + it * 32 + lane;
        const bool on = idx < n;
        const uint64_t key = on ? in[idx] : 0ull;
        const uint32_t d = on ? (uint32_t)((key >> shift) & 255u) : (256u + lane);  // idle lanes match nobody
        const uint32_t peers = __match_any_sync(kFull, d);
        uint32_t dst = 0;
        if (on) dst = s_cnt[warp][d] + __popc(peers & lt);
        __syncwarp();
        if (on && (peers & lt) == 0u) s_cnt[warp][d] += __popc(peers);  // lowest peer advances the cursor
        __syncwarp();
        if (on) out[dst] = key;
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_mt_unique(const uint64_t *__restrict__ keys, int64_t n, int64_t num_tiles,
                                                        MtHeader *hdr, unsigned long long *status,
                                                        uint64_t *__restrict__ ukeys) {
    __shared__ unsigned long long s_warp[kThreads / 32];
    __shared__ unsigned long long s_excl;
    __shared__ unsigned int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kPer = kUniqTile / kThreads;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&hdr->ticket_unique, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= num_tiles) break;
        const int64_t i0 = tile * kUniqTile + (int64_t)threadIdx.x * kPer;
        uint64_t k[kPer];
        uint32_t flags = 0;
        uint64_t prev = (i0 > 0 && i0 <= n) ? keys[i0 - 1] : 0ull;
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int64_t i = i0 + j;
            k[j] = i < n ? keys[i] : 0ull;
            if (i < n && (i == 0 || k[j] != prev)) flags |= 1u << j;
            prev = k[j];
        }
        unsigned long long agg;
        const unsigned long long excl = block_excl_scan(__popc(flags), s_warp, &agg);
        if (warp == 0) {
            const unsigned long long e = lookback(status, tile, agg, lane);
            if (lane == 0) {
                s_excl = e;
                if (tile == num_tiles - 1) hdr->num_unique = e + agg;
            }
        }
        __syncthreads();
        unsigned long long pos = s_excl + excl;
#pragma unroll
        for (int j = 0; j < kPer; ++j)
            if (flags & (1u << j)) ukeys[pos++] = k[j];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_mt_verts(const float *__restrict__ pts, const float *__restrict__ sdf,
                                                       const uint64_t *__restrict__ ukeys, int64_t V,
                                                       float *__restrict__ verts, int64_t *__restrict__ edges) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    emit_vertex(pts, sdf, ukeys[i], i, verts, edges);
}

__device__ __forceinline__ int64_t find_key(const uint64_t *__restrict__ ukeys, int64_t V, uint64_t key) {
    int64_t lo = 0, hi = V;  // first index with ukeys[idx] >= key
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(ukeys + mid) < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kThreads) k_mt_faces(const int64_t *__restrict__ tets, const uint8_t *__restrict__ codes,
                                                       const uint32_t *__restrict__ slot_tet, int64_t n1, int64_t F,
                                                       const uint64_t *__restrict__ ukeys, int64_t V,
                                                       int64_t *__restrict__ faces, int64_t *__restrict__ tet_idx) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= F) return;
    int64_t slot;
    int tri;
    if (s < n1) {
        slot = s;
        tri = 0;
    } else {
        slot = n1 + ((s - n1) >> 1);
        tri = (int)((s - n1) & 1);
    }
    const int64_t t = slot_tet[slot];
    const longlong2 lo = reinterpret_cast<const longlong2 *>(tets)[2 * t];
    const longlong2 hi = reinterpret_cast<const longlong2 *>(tets)[2 * t + 1];
    const int64_t id[4] = {lo.x, lo.y, hi.x, hi.y};
    const uint32_t row = c_tri_rows[codes[t]] >> (12 * tri);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t e = (row >> (4 * k)) & 15u;
        const int a = (kEdgeA >> (2 * e)) & 3, b = (kEdgeB >> (2 * e)) & 3;
        int64_t ia = id[0], ib = id[0];
#pragma unroll
        for (int c = 1; c < 4; ++c) {
            if (a == c) ia = id[c];
            if (b == c) ib = id[c];
        }
        faces[3 * s + k] = find_key(ukeys, V, make_key(ia, ib));
    }
    if (tet_idx) tet_idx[s] = t;
}

// d verts / d points, d verts / d sdf  (v = pa*w0 + pb*w1, w0 = -sb/D, w1 = sa/D, D = sa - sb)
__global__ void __launch_bounds__(kThreads) k_mt_backward(const float *__restrict__ pts, const float *__restrict__ sdf,
                                                          const int64_t *__restrict__ edges, int64_t V,
                                                          const float *__restrict__ gv, float *gp, float *gs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int64_t a = edges[2 * i], b = edges[2 * i + 1];
    const float sa = sdf[a], sb = sdf[b];
    const float D = sa - sb, inv = 1.0f / D;
    const float w0 = -sb * inv, w1 = sa * inv;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float g = gv[3 * i + c];
        atomicAdd(gp + 3 * a + c, g * w0);
        atomicAdd(gp + 3 * b + c, g * w1);
        dot += g * (pts[3 * a + c] - pts[3 * b + c]);
    }
    atomicAdd(gs + a, dot * sb * inv * inv);
    atomicAdd(gs + b, -dot * sa * inv * inv);
}

// ------------------------------------------------------------------------------------------------
struct MtLayout {
    size_t header, status_compact, status_unique, slot_tet, keys_a, keys_b, ghist, total;
    int64_t tiles_compact, tiles_unique;
    int sort_blocks;
};

MtLayout mt_layout(int64_t T, int64_t n1, int64_t n2, int64_t ne) {
    MtLayout l;
    l.tiles_compact = (T + kTileTets - 1) / kTileTets;
    l.tiles_unique = (ne + kUniqTile - 1) / kUniqTile;
    l.sort_blocks = (int)((ne + kSortTile - 1) / kSortTile);
    size_t off = 0;
    l.header = off;         off += up256(sizeof(MtHeader));
    l.status_compact = off; off += up256((size_t)l.tiles_compact * 8);
    l.status_unique = off;  off += up256((size_t)l.tiles_unique * 8);
    l.slot_tet = off;       off += up256((size_t)(n1 + n2) * 4);
    l.keys_a = off;         off += up256((size_t)ne * 8);
    l.keys_b = off;         off += up256((size_t)ne * 8);
    l.ghist = off;          off += up256((size_t)l.sort_blocks * 256 * 4);
    l.total = off < 256 ? 256 : off;
    return l;
}

int64_t *mt_pinned() {
    thread_local int64_t *buf = nullptr;
    if (!buf && cudaHostAlloc(reinterpret_cast<void **>(&buf), 4 * sizeof(int64_t), cudaHostAllocPortable) != cudaSuccess)
        buf = nullptr;
    return buf;
}

}  // namespace

}  // namespace p3d

using namespace p3d;

#define MT_FAIL(st, msg) return p3d::set_error(st, msg)
#define MT_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) MT_FAIL(P3D_ERR_CUDA, std::string(#expr " failed: ") + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

size_t p3d_mt_codes_bytes(int64_t num_tets) {
    if (num_tets < 0) return 0;
    return up256((size_t)num_tets) + 256;  // code bytes + the classify counters
}

p3d_status p3d_mt_classify(const float *points, int64_t num_points, int64_t *tets, int64_t num_tets, const float *sdf,
                           int oriented, uint8_t *codes, int64_t *counts_host, void *stream) {
    if (num_tets < 0 || num_points < 0 || !counts_host || !codes) MT_FAIL(P3D_ERR_INVALID, "p3d_mt_classify: invalid argument");
    if (num_tets > 0 && (!points || !tets || !sdf)) MT_FAIL(P3D_ERR_INVALID, "p3d_mt_classify: null pointer");
    if (num_tets > (int64_t)INT32_MAX || num_points > ((int64_t)1 << 32)) MT_FAIL(P3D_ERR_OVERFLOW, "p3d_mt_classify: more than 2^31 tets or 2^32 points");
    if (reinterpret_cast<uintptr_t>(tets) & 15) MT_FAIL(P3D_ERR_INVALID, "p3d_mt_classify: tets must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    ClassifyCounters *ctr = reinterpret_cast<ClassifyCounters *>(codes + up256((size_t)num_tets));
    MT_CUDA(cudaMemsetAsync(ctr, 0, sizeof(ClassifyCounters), s));
    if (num_tets > 0) {
        const int64_t want = (num_tets + kThreads - 1) / kThreads;
        const int64_t cap = (int64_t)sm_count() * 16;
        k_mt_classify<<<(unsigned)(want < cap ? want : cap), kThreads, 0, s>>>(points, num_points, tets, num_tets, sdf, codes, ctr, oriented);
        MT_CUDA(cudaGetLastError());
    }
    int64_t *pin = mt_pinned();
    int64_t *dst = pin ? pin : counts_host;
    MT_CUDA(cudaMemcpyAsync(dst, ctr, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    MT_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < 3; ++i) counts_host[i] = dst[i];
    if (dst[3] != 0) MT_FAIL(P3D_ERR_INVALID, "p3d_mt_classify: " + std::to_string(dst[3]) + " tets name a point outside [0, num_points)");
    return P3D_OK;
}

size_t p3d_mt_workspace_bytes(int64_t num_tets, int64_t n1, int64_t n2, int64_t ne) {
    if (num_tets < 0 || n1 < 0 || n2 < 0 || ne < 0) return 0;
    return mt_layout(num_tets, n1, n2, ne).total;
}

p3d_status p3d_mt_index(const int64_t *tets, int64_t num_tets, int64_t num_points, const float *sdf, const uint8_t *codes,
                        int64_t n1, int64_t n2, int64_t ne, void *workspace, size_t workspace_bytes, int64_t *counts_host,
                        void *stream) {
    (void)sdf;
    if (!workspace || !counts_host || num_tets < 0 || n1 < 0 || n2 < 0 || ne < 0) MT_FAIL(P3D_ERR_INVALID, "p3d_mt_index: invalid argument");
    const MtLayout l = mt_layout(num_tets, n1, n2, ne);
    if (workspace_bytes < l.total) MT_FAIL(P3D_ERR_WORKSPACE, "p3d_mt_index: workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace) & 255) MT_FAIL(P3D_ERR_INVALID, "p3d_mt_index: workspace must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char *base = static_cast<char *>(workspace);
    MtHeader *hdr = reinterpret_cast<MtHeader *>(base + l.header);
    unsigned long long *st_c = reinterpret_cast<unsigned long long *>(base + l.status_compact);
    unsigned long long *st_u = reinterpret_cast<unsigned long long *>(base + l.status_unique);
    uint32_t *slot_tet = reinterpret_cast<uint32_t *>(base + l.slot_tet);
    uint64_t *keys_a = reinterpret_cast<uint64_t *>(base + l.keys_a);
    uint64_t *keys_b = reinterpret_cast<uint64_t *>(base + l.keys_b);
    uint32_t *ghist = reinterpret_cast<uint32_t *>(base + l.ghist);

    MT_CUDA(cudaMemsetAsync(base, 0, l.slot_tet, s));  // header + both status arrays
    counts_host[0] = 0;
    if (num_tets == 0 || n1 + n2 == 0) return P3D_OK;
    const int sms = sm_count();
    {
        const int64_t cap = (int64_t)sms * 8;
        k_mt_compact<<<(unsigned)(l.tiles_compact < cap ? l.tiles_compact : cap), kThreads, 0, s>>>(
            tets, num_tets, codes, l.tiles_compact, (unsigned long long)n1, hdr, st_c, slot_tet, keys_a);
    }
    // LSD radix sort on the significant bits of (min id << 32 | max id)
    const int bits = id_bits(num_points);
    uint64_t *src = keys_a, *dst = keys_b;
    for (int half = 0; half < 2; ++half)
        for (int sh = 0; sh < bits; sh += 8) {
            const int shift = half * 32 + sh;
            k_sort_hist<<<l.sort_blocks, kThreads, 0, s>>>(src, ne, shift, ghist, l.sort_blocks);
            if (l.sort_blocks <= 512) {
                k_sort_scatter<true><<<l.sort_blocks, kThreads, 0, s>>>(src, dst, ne, shift, ghist, l.sort_blocks);
            } else {
                k_sort_scan<<<1, 1024, 0, s>>>(ghist, (int64_t)l.sort_blocks * 256);
                k_sort_scatter<false><<<l.sort_blocks, kThreads, 0, s>>>(src, dst, ne, shift, ghist, l.sort_blocks);
            }
            uint64_t *t = src;
            src = dst;
            dst = t;
        }
    // the pass count is even, so the sorted keys are back in keys_a and keys_b takes the unique keys
    {
        const int64_t cap = (int64_t)sms * 8;
        k_mt_unique<<<(unsigned)(l.tiles_unique < cap ? l.tiles_unique : cap), kThreads, 0, s>>>(src, ne, l.tiles_unique, hdr,
                                                                                           st_u, dst);
    }
    MT_CUDA(cudaGetLastError());
    int64_t *pin = mt_pinned();
    int64_t *out = pin ? pin : counts_host;
    MT_CUDA(cudaMemcpyAsync(out, &hdr->num_unique, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    MT_CUDA(cudaStreamSynchronize(s));
    counts_host[0] = out[0];
    return P3D_OK;
}

p3d_status p3d_mt_emit(const float *points, const int64_t *tets, int64_t num_tets, const float *sdf, const uint8_t *codes,
                       int64_t n1, int64_t n2, int64_t ne, int64_t num_vertices, const void *workspace, float *verts,
                       int64_t *edges, int64_t *faces, int64_t *tet_idx, void *stream) {
    if (num_tets < 0 || n1 < 0 || n2 < 0 || ne < 0 || num_vertices < 0) MT_FAIL(P3D_ERR_INVALID, "p3d_mt_emit: invalid argument");
    const int64_t F = n1 + 2 * n2;
    if (F == 0 && num_vertices == 0) return P3D_OK;
    if (!workspace || !points || !tets || !sdf || !codes || (num_vertices && !verts) || (F && !faces))
        MT_FAIL(P3D_ERR_INVALID, "p3d_mt_emit: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const MtLayout l = mt_layout(num_tets, n1, n2, ne);
    const char *base = static_cast<const char *>(workspace);
    const uint32_t *slot_tet = reinterpret_cast<const uint32_t *>(base + l.slot_tet);
    const uint64_t *ukeys = reinterpret_cast<const uint64_t *>(base + l.keys_b);  // see p3d_mt_index
    if (num_vertices > 0)
        k_mt_verts<<<(unsigned)((num_vertices + kThreads - 1) / kThreads), kThreads, 0, s>>>(points, sdf, ukeys, num_vertices,
                                                                                          verts, edges);
    if (F > 0)
        k_mt_faces<<<(unsigned)((F + kThreads - 1) / kThreads), kThreads, 0, s>>>(tets, codes, slot_tet, n1, F, ukeys,
                                                                               num_vertices, faces, tet_idx);
    MT_CUDA(cudaGetLastError());
    return P3D_OK;
}

p3d_status p3d_mt_backward(const float *points, const float *sdf, const int64_t *edges, int64_t num_vertices,
                           const float *grad_verts, float *grad_points, float *grad_sdf, void *stream) {
    if (num_vertices < 0) MT_FAIL(P3D_ERR_INVALID, "p3d_mt_backward: invalid argument");
    if (num_vertices == 0) return P3D_OK;
    if (!points || !sdf || !edges || !grad_verts || !grad_points || !grad_sdf) MT_FAIL(P3D_ERR_INVALID, "p3d_mt_backward: null pointer");
    k_mt_backward<<<(unsigned)((num_vertices + kThreads - 1) / kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        points, sdf, edges, num_vertices, grad_verts, grad_points, grad_sdf);
    MT_CUDA(cudaGetLastError());
    return P3D_OK;
}

}  // extern "C"
