// primitive3d_b200/csrc/mt_extract.cu -- marching tetrahedra in ONE call, four launches, one host
// synchronisation (p3d_mt_extract, include/prim3d_b200.h).
//
// Same outputs as the staged calls of mt_kernels.cu (and as the reference,
// prim3d/utility/marching_tetrahedras.py:89-235); what changes is how they are reached when the crossing
// edges are few against the tets, the case the path exists for (an isosurface through a tet grid cuts
// under 1 % of the tets: Kuhn 128^3 has 12.3 M tets, 85 k valid ones, 341 k crossing-edge instances):
//
//   k_mtx_classify  the streaming pass over the tets (32 B each, the only HBM-sized traffic of the path),
//                   with no barrier and no scan in it: orientation fix in place (:148), occupancy code
//                   (:151-154), totals.  No per-tet code array is written or read back.  The orientation
//                   sign comes from a float32 triple product with a forward error bound and falls back to
//                   the float64 one (orient_negative_f64, shared with the staged kernel) when the bound
//                   does not decide, so both paths flip exactly the same tets.  The rare valid tet is
//                   appended, through atomic cursors, to BUCKETS: the tet itself (id, code) to the slot
//                   bucket of its id, each of its crossing edges, a 64-bit key (min id << 32 | max id), to
//                   the key bucket of its min id.  A bucket is a fixed-capacity region; the bucket index is
//                   the id scaled onto the bucket count, monotone in the id.
//   k_mtx_sort      one CTA per bucket: the bucket is sorted in shared memory by counting ranks inside
//                   sub-buckets (which also drops duplicates), written back compacted, counted; the counts
//                   are also summed per group of 64 buckets.
//   k_mtx_emit      one CTA per bucket: the counts of all buckets before it (whole groups, then its own
//                   group: one load per thread) number its entries globally, which gives
//                     key buckets   the global number of each unique key -- bucket order then key order IS
//                                   the lexicographic order torch.unique(dim=0) defines (:157-173) -- whose
//                                   vertex is interpolated and written (:177-189);
//                     slot buckets  each valid tet's place in the reference's face order: all one-triangle
//                                   tets, then all two-triangle tets, each group in tet order (:205-223).
//   k_mtx_faces     per face slot: the tet's table row (:193-223), each edge's vertex id = unique base of
//                   the key's bucket + a binary search among that bucket's unique keys only; tet_idx.
//
// All capacities are speculative (the caller's guess, the previous call's counts): the kernels count
// everything, write only what fits, and the one host read at the end says whether the outputs are complete.
// Inputs this layout does not suit (a bucket over its capacity: crossing edges in the millions, or ids
// crowded into a few buckets) are reported as such and go through the staged calls, which sort with the
// general LSD radix sort.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "mt_common.cuh"
#include "p3d_error.h"
#include "scan_utils.cuh"

namespace p3d {
namespace {

constexpr int kXThreads = 256;
constexpr int kBucketCap = 2048;                // entries per bucket (16 KB each in global memory)
constexpr int kBucketMean = 128;                // buckets are sized for this many entries on average
constexpr int kMaxKeyBucketBits = 13;           // at most 8192 key buckets (64 MB of bucket regions)
constexpr int kMaxSlotBucketBits = 11;          // at most 2048 slot buckets
constexpr int kSortThreads = 128;
#ifndef P3D_MTX_SUBBITS
#define P3D_MTX_SUBBITS 4
#endif
constexpr int kSubBits = P3D_MTX_SUBBITS, kSub = 1 << kSubBits;  // sub-buckets of the counting sort inside a bucket
constexpr int kGroup = 64;                      // buckets per group of the two-level sum of the bucket counts
constexpr uint64_t kHole = ~0ull;               // no key and no slot entry has this value

// Bucket of an id in [0, N): the id scaled onto [0, num_buckets), monotone in the id, so bucket order then
// entry order is entry order.  scale = floor(num_buckets * 2^32 / N).
__device__ __forceinline__ uint32_t bucket_of(uint64_t id, uint64_t scale) { return (uint32_t)((id * scale) >> 32); }

struct XHeader {
    unsigned int pad;
    unsigned int overflow;            // a bucket received more than kBucketCap entries
    unsigned long long n1, n2, ne;    // one-triangle tets, two-triangle tets, crossing-edge instances
    unsigned long long bad;           // tets naming a point outside [0, P)
    unsigned long long num_unique;    // V
};

struct XBuckets {   // one class of buckets (slots or keys)
    uint32_t *cursors;
    uint64_t *regions;
    unsigned long long *counts;      // per bucket, once sorted: unique keys, or one- | two-triangle tets << 31
    unsigned long long *group_sum;   // per group of kGroup buckets: the sum of their counts
    uint64_t scale;
    int count;
};

// The rare valid tet (under 1 % of them on an isosurface) and its crossing edges go to their buckets; out of line so
// that its registers are not the streaming loop's.  The cursors are taken together, then the stores.
static __device__ __noinline__ void append_valid(int64_t t, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3, uint32_t code,
                                                 uint32_t *slot_cursors, uint64_t *slot_regions, uint64_t slot_scale,
                                                 uint32_t *key_cursors, uint64_t *key_regions, uint64_t key_scale) {
    const uint32_t id[4] = {i0, i1, i2, i3};
    uint64_t key[4];
    int nk = 0;
#pragma unroll
    for (int e = 0; e < 6; ++e) {
        const int a = (kEdgeA >> (2 * e)) & 3, b = (kEdgeB >> (2 * e)) & 3;
        if (((code >> a) ^ (code >> b)) & 1u) {
            const uint64_t k = make_key(id[a], id[b]);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q == nk) key[q] = k;
            ++nk;
        }
    }
    const uint32_t sb = bucket_of((uint64_t)t, slot_scale);
    const uint32_t sat = atomicAdd(slot_cursors + sb, 1u);
    uint32_t kb[4], kat[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (q < nk) {
            kb[q] = bucket_of(key[q] >> 32, key_scale);
            kat[q] = atomicAdd(key_cursors + kb[q], 1u);
        }
    if (sat < (uint32_t)kBucketCap) slot_regions[(size_t)sb * kBucketCap + sat] = ((uint64_t)t << 4) | code;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (q < nk && kat[q] < (uint32_t)kBucketCap) key_regions[(size_t)kb[q] * kBucketCap + kat[q]] = key[q];
}

// ------------------------------------------------------------------------------------------------
// (x, y, z, sdf) per point, 16 bytes: the classify pass gathers four corners per tet, and one 16-byte load per
// corner costs the L1 a quarter of what three position loads and an sdf load do.  Same values, no arithmetic.
__global__ void __launch_bounds__(kXThreads) k_mtx_pack(const float *__restrict__ pts, const float *__restrict__ sdf, int64_t P,
                                                        float4 *__restrict__ packed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) packed[i] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], sdf[i]);
}

#ifndef P3D_MTX_CTAS
#define P3D_MTX_CTAS 4
#endif
#ifndef P3D_MTX_WIDE
#define P3D_MTX_WIDE 1
#endif
#ifndef P3D_MTX_PACK
#define P3D_MTX_PACK 1
#endif
#ifndef P3D_MTX_GRID
#define P3D_MTX_GRID 4
#endif
// A tet's four int64 indices: one 32-byte load (LDG.256, new with sm_100) when the array is 32-byte aligned, so that
// every sector fetched is used by the instruction that fetched it; two 16-byte loads otherwise.
template <bool WIDE>
__device__ __forceinline__ void tet_load(const longlong2 *p, longlong2 &lo, longlong2 &hi) {
    if (WIDE) {
        asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(lo.x), "=l"(lo.y), "=l"(hi.x), "=l"(hi.y) : "l"(p));
    } else {
        lo = p[0];
        hi = p[1];
    }
}
template <bool WIDE>
__global__ void __launch_bounds__(kXThreads, P3D_MTX_CTAS)
k_mtx_classify(const float *__restrict__ pts, uint32_t P, int64_t *tets, int64_t T, const float *__restrict__ sdf,
               const float4 *__restrict__ packed, XHeader *hdr, XBuckets slots, XBuckets keys, int oriented) {
    const int lane = threadIdx.x & 31;
    longlong2 *t2 = reinterpret_cast<longlong2 *>(tets);
    unsigned long long n1 = 0, n2 = 0, bad = 0;
    auto one = [&](int64_t t, longlong2 lo, longlong2 hi) {
        uint32_t id[4] = {(uint32_t)lo.x, (uint32_t)lo.y, (uint32_t)hi.x, (uint32_t)hi.y};
        if ((uint64_t)lo.x >= P || (uint64_t)lo.y >= P || (uint64_t)hi.x >= P || (uint64_t)hi.y >= P) {  // P3D_ERR_INVALID, see k_mt_classify
            ++bad;
            return;
        }
#if P3D_MTX_PACK
        // one 16-byte gather per corner from the packed (x, y, z, sdf) copy of the points
        const float4 c0 = __ldg(packed + id[0]), c1 = __ldg(packed + id[1]), c2 = __ldg(packed + id[2]), c3 = __ldg(packed + id[3]);
        uint32_t code = (c0.w > 0.f ? 1u : 0u) | (c1.w > 0.f ? 2u : 0u) | (c2.w > 0.f ? 4u : 0u) | (c3.w > 0.f ? 8u : 0u);
        if (!oriented) {
            const float ax = c1.x - c0.x, ay = c1.y - c0.y, az = c1.z - c0.z;
            const float bx = c2.x - c0.x, by = c2.y - c0.y, bz = c2.z - c0.z;
            const float cx = c3.x - c0.x, cy = c3.y - c0.y, cz = c3.z - c0.z;
#else
        const float s0 = __ldg(sdf + id[0]), s1 = __ldg(sdf + id[1]), s2 = __ldg(sdf + id[2]), s3 = __ldg(sdf + id[3]);
        uint32_t code = (s0 > 0.f ? 1u : 0u) | (s1 > 0.f ? 2u : 0u) | (s2 > 0.f ? 4u : 0u) | (s3 > 0.f ? 8u : 0u);
        if (!oriented) {
            const float *q0 = pts + 3ull * id[0], *q1 = pts + 3ull * id[1], *q2 = pts + 3ull * id[2], *q3 = pts + 3ull * id[3];
            const float p0x = __ldg(q0), p0y = __ldg(q0 + 1), p0z = __ldg(q0 + 2);
            const float ax = __ldg(q1) - p0x, ay = __ldg(q1 + 1) - p0y, az = __ldg(q1 + 2) - p0z;
            const float bx = __ldg(q2) - p0x, by = __ldg(q2 + 1) - p0y, bz = __ldg(q2 + 2) - p0z;
            const float cx = __ldg(q3) - p0x, cy = __ldg(q3 + 1) - p0y, cz = __ldg(q3 + 2) - p0z;
#endif
            // a . (b x c) in float32 with the sum of the magnitudes of its terms: about nine roundings, each 2^-24
            // relative, so a result above 2^-20 of that sum (plus a floor against underflow) has the sign of the exact
            // value -- and of the float64 evaluation, whose own error is 2^-29 times smaller
            const float m1 = by * cz, m2 = bz * cy, m3 = bz * cx, m4 = bx * cz, m5 = bx * cy, m6 = by * cx;
            const float det = ax * (m1 - m2) + ay * (m3 - m4) + az * (m5 - m6);
            const float mag = fabsf(ax) * (fabsf(m1) + fabsf(m2)) + fabsf(ay) * (fabsf(m3) + fabsf(m4)) + fabsf(az) * (fabsf(m5) + fabsf(m6));
            bool neg;
            if (fabsf(det) > fmaf(mag, 0x1p-20f, 1e-30f)) neg = det < 0.f;
            else neg = orient_negative_f64(pts, id[0], id[1], id[2], id[3]);
            if (neg) {  // :148 tets[flip, :2] = tets[flip][:, [1, 0]]
                t2[2 * t] = make_longlong2(lo.y, lo.x);
                const uint32_t s = id[0];
                id[0] = id[1];
                id[1] = s;
                code = (code & 12u) | ((code & 1u) << 1) | ((code >> 1) & 1u);
            }
        }
        const uint32_t nt = num_tri(code);
        if (nt == 0u) return;
        n1 += nt == 1u;
        n2 += nt == 2u;
        append_valid(t, id[0], id[1], id[2], id[3], code, slots.cursors, slots.regions, slots.scale, keys.cursors, keys.regions, keys.scale);
    };
    // two tets per iteration, their index loads issued together: the dependent gathers of one overlap the other's
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    longlong2 lo0, hi0, lo1, hi1;
    for (; t + stride < T; t += 2 * stride) {
        tet_load<WIDE>(t2 + 2 * t, lo0, hi0);
        tet_load<WIDE>(t2 + 2 * (t + stride), lo1, hi1);
        one(t, lo0, hi0);
        one(t + stride, lo1, hi1);
    }
    if (t < T) {
        tet_load<WIDE>(t2 + 2 * t, lo0, hi0);
        one(t, lo0, hi0);
    }
    n1 = warp_sum64(n1 | (n2 << 32));  // both below 2^31 per warp
    bad = warp_sum64(bad);
    if (lane == 0) {
        if (n1 & 0xffffffffull) atomicAdd(&hdr->n1, n1 & 0xffffffffull);
        if (n1 >> 32) atomicAdd(&hdr->n2, n1 >> 32);
        if (bad) atomicAdd(&hdr->bad, bad);
    }
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan of one 64-bit value per thread over the CTA; *total = the sum.
__device__ __forceinline__ unsigned long long cta_excl_scan64(unsigned long long v, unsigned long long *s_warp, unsigned long long *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long incl = warp_incl_scan64(v, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long lower = 0, sum = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) {
        const unsigned long long t = s_warp[w];
        if (w < warp) lower += t;
        sum += t;
    }
    __syncthreads();
    *total = sum;
    return lower + incl - v;
}

// One CTA per bucket (slot buckets first, then key buckets): sorted and de-duplicated in shared memory, written back
// to the front of the bucket's region, counted.
//
// The sort is by counting: an entry goes to position #{entries smaller than it}; equal keys land on the same position
// (they are the same value) and leave holes behind them, which is the de-duplication.  To keep the comparisons far
// below n^2 the bucket is first split into kSub sub-buckets by the next bits of the scaled id (a counting scatter in
// shared memory); a warp then ranks 32 consecutive entries against the sub-buckets they span only, every comparand
// one broadcast load for the warp.
__global__ void __launch_bounds__(kSortThreads)
k_mtx_sort(XHeader *hdr, XBuckets slots, XBuckets keys) {
    __shared__ uint64_t s_tmp[kBucketCap];
    __shared__ uint64_t s_out[kBucketCap];
    __shared__ uint32_t s_sub[kSub + 1], s_cur[kSub];
    __shared__ unsigned long long s_warp[kSortThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int job = blockIdx.x;
    const bool is_key = job >= slots.count;
    const XBuckets &cls = is_key ? keys : slots;
    const int b = is_key ? job - slots.count : job;
    const uint32_t filled = cls.cursors[b];
    uint32_t n = filled;
    if (n > (uint32_t)kBucketCap) {  // reported; the bucket is skipped and the caller takes another path
        if (threadIdx.x == 0) hdr->overflow = 1u;
        n = 0;
    }
    if (n == 0) {
        if (threadIdx.x == 0) cls.counts[b] = 0ull;
        return;
    }
    if (is_key && threadIdx.x == 0) atomicAdd(&hdr->ne, (unsigned long long)filled);
    uint64_t *region = cls.regions + (size_t)b * kBucketCap;
    const uint64_t scale = cls.scale;
    const int id_shift = is_key ? 32 : 4;  // key: min id << 32 | max id; slot entry: tet << 4 | code
    auto sub_of = [&](uint64_t e) { return (uint32_t)(((e >> id_shift) * scale) >> (32 - kSubBits)) & (uint32_t)(kSub - 1); };
    if (threadIdx.x <= kSub) s_sub[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += kSortThreads) atomicAdd(&s_sub[sub_of(region[i]) + 1], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k <= kSub; ++k) s_sub[k] += s_sub[k - 1];  // s_sub[k] = first position of sub-bucket k
    }
    __syncthreads();
    if (threadIdx.x < kSub) s_cur[threadIdx.x] = s_sub[threadIdx.x];
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += kSortThreads) {
        const uint64_t e = region[i];
        s_tmp[atomicAdd(&s_cur[sub_of(e)], 1u)] = e;
        s_out[i] = kHole;
    }
    __syncthreads();
    for (uint32_t base = warp * 32; base < n; base += kSortThreads) {
        const uint32_t i = base + lane;
        const uint64_t e = s_tmp[i < n ? i : n - 1];
        const uint32_t sub = sub_of(e);
        const uint32_t from = s_sub[__shfl_sync(kFull, sub, 0)], to = s_sub[__shfl_sync(kFull, sub, 31) + 1];
        uint32_t rank = 0;
        for (uint32_t j = from; j < to; ++j) rank += s_tmp[j] < e ? 1u : 0u;
        if (i < n) s_out[from + rank] = e;
    }
    __syncthreads();
    const uint64_t *sorted = s_out;  // sorted entries with holes where duplicates were
    // a contiguous run of sorted entries per thread; the survivors go back to the front of the region, in order
    const uint32_t per = (n + kSortThreads - 1) / kSortThreads;
    const uint32_t i0 = threadIdx.x * per;
    unsigned long long mine = 0, kinds = 0;
    for (uint32_t j = 0; j < per; ++j) {
        const uint32_t i = i0 + j;
        if (i >= n) break;
        const uint64_t e = sorted[i];
        if (e == kHole) continue;
        ++mine;
        if (!is_key) kinds += num_tri((uint32_t)e & 15u) == 1u ? 1ull : (1ull << 31);
    }
    unsigned long long agg, kagg = 0;
    const unsigned long long excl = cta_excl_scan64(mine, s_warp, &agg);
    if (!is_key) cta_excl_scan64(kinds, s_warp, &kagg);
    uint32_t r = (uint32_t)excl;
    for (uint32_t j = 0; j < per; ++j) {
        const uint32_t i = i0 + j;
        if (i >= n) break;
        const uint64_t e = sorted[i];
        if (e != kHole) region[r++] = e;
    }
    if (threadIdx.x == 0) {
        const unsigned long long c = is_key ? agg : kagg;
        cls.counts[b] = c;
        atomicAdd(cls.group_sum + b / kGroup, c);
    }
}

// One CTA per bucket again, now that every count is known: the sum of the counts before the bucket (whole groups,
// then the buckets of its own group) numbers its entries globally.  Key buckets write their vertices, slot buckets
// place their tets in the two face-order lists.
__global__ void __launch_bounds__(kSortThreads)
k_mtx_emit(const float *__restrict__ pts, const float *__restrict__ sdf, XHeader *hdr, XBuckets slots, XBuckets keys,
           uint32_t *__restrict__ ubase, float *__restrict__ verts, int64_t *__restrict__ edges, int64_t vcap,
           uint64_t *__restrict__ list1, uint64_t *__restrict__ list2, int64_t slot_cap) {
    __shared__ unsigned long long s_warp[kSortThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int job = blockIdx.x;
    const bool is_key = job >= slots.count;
    const XBuckets &cls = is_key ? keys : slots;
    const int b = is_key ? job - slots.count : job;
    const int g = b / kGroup;
    static_assert(kSortThreads >= kGroup && (1 << kMaxKeyBucketBits) / kGroup <= kSortThreads, "one load per thread");
    unsigned long long before = (int)threadIdx.x < g ? cls.group_sum[threadIdx.x] : 0ull;
    if (g * kGroup + (int)threadIdx.x < b) before += cls.counts[g * kGroup + threadIdx.x];
    const unsigned long long count = cls.counts[b];
    before = warp_sum64(before);
    if (lane == 0) s_warp[warp] = before;
    __syncthreads();
    before = 0;
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) before += s_warp[w];
    __syncthreads();
    const uint64_t *region = cls.regions + (size_t)b * kBucketCap;
    if (is_key) {
        if (threadIdx.x == 0) {
            ubase[b] = (uint32_t)before;
            if (b == keys.count - 1) {
                ubase[keys.count] = (uint32_t)(before + count);
                hdr->num_unique = before + count;
            }
        }
        for (uint32_t i = threadIdx.x; i < (uint32_t)count; i += kSortThreads) {
            const int64_t v = (int64_t)(before + i);
            if (v < vcap) emit_vertex(pts, sdf, region[i], v, verts, edges);
        }
    } else {
        const uint32_t n = (uint32_t)(count & 0x7fffffffull) + (uint32_t)(count >> 31);
        if (n == 0) return;
        const uint32_t per = (n + kSortThreads - 1) / kSortThreads;
        const uint32_t i0 = threadIdx.x * per;
        unsigned long long kinds = 0;
        for (uint32_t j = 0; j < per; ++j)
            if (i0 + j < n) kinds += num_tri((uint32_t)region[i0 + j] & 15u) == 1u ? 1ull : (1ull << 31);
        unsigned long long kagg;
        const unsigned long long pos = before + cta_excl_scan64(kinds, s_warp, &kagg);
        int64_t p1 = (int64_t)(pos & 0x7fffffffull), p2 = (int64_t)(pos >> 31);
        for (uint32_t j = 0; j < per; ++j) {
            if (i0 + j >= n) break;
            const uint64_t e = region[i0 + j];
            const uint64_t entry = ((e >> 4) & 0xffffffffull) | ((e & 15ull) << 32);  // tet | code << 32
            if (num_tri((uint32_t)e & 15u) == 2u) {
                if (p2 < slot_cap) list2[p2] = entry;
                ++p2;
            } else {
                if (p1 < slot_cap) list1[p1] = entry;
                ++p1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kXThreads)
k_mtx_faces(const int64_t *__restrict__ tets, const XHeader *__restrict__ hdr, const uint64_t *__restrict__ list1,
            const uint64_t *__restrict__ list2, int64_t slot_cap, const uint64_t *__restrict__ buckets,
            const uint32_t *__restrict__ ubase, uint64_t scale, int64_t fcap, int64_t *__restrict__ faces,
            int64_t *__restrict__ tet_idx) {
    const int64_t n1 = (int64_t)hdr->n1, n2 = (int64_t)hdr->n2;
    const int64_t F = n1 + 2 * n2;
    if (n1 > slot_cap || n2 > slot_cap || F > fcap || hdr->overflow) return;  // incomplete lists: the caller runs again
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= F) return;
    uint64_t entry;
    int tri = 0;
    if (s < n1) {
        entry = list1[s];
    } else {
        entry = list2[(s - n1) >> 1];
        tri = (int)((s - n1) & 1);
    }
    const int64_t t = (int64_t)(entry & 0xffffffffull);
    const uint32_t code = (uint32_t)(entry >> 32);
    const longlong2 lo = reinterpret_cast<const longlong2 *>(tets)[2 * t];
    const longlong2 hi = reinterpret_cast<const longlong2 *>(tets)[2 * t + 1];
    const int64_t id[4] = {lo.x, lo.y, hi.x, hi.y};
    const uint32_t row = c_tri_rows[code] >> (12 * tri);
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
        const uint64_t key = make_key(ia, ib);
        const uint32_t bucket = bucket_of(key >> 32, scale);
        const uint32_t u0 = ubase[bucket], u1 = ubase[bucket + 1];
        const uint64_t *region = buckets + (size_t)bucket * kBucketCap;
        uint32_t a0 = 0, a1 = u1 - u0;  // first index with region[idx] >= key
        while (a0 < a1) {
            const uint32_t mid = (a0 + a1) >> 1;
            if (__ldg(region + mid) < key) a0 = mid + 1;
            else a1 = mid;
        }
        faces[3 * s + k] = (int64_t)u0 + a0;
    }
    if (tet_idx) tet_idx[s] = t;
}

struct XLayout {
    size_t header, cursors_slots, cursors_keys, groups_slots, groups_keys, zeroed, counts_slots, counts_keys, packed, ubase, list1, list2, regions_slots, regions_keys, total;
    int slot_buckets, key_buckets;
    uint64_t slot_scale, key_scale;
};

int bucket_count(int64_t expected, int max_bits, int64_t ids) {
    int bits = 0;
    while (bits < max_bits && ((int64_t)kBucketMean << bits) < expected) ++bits;
    while (bits > 0 && ((int64_t)1 << bits) > ids) --bits;  // never more buckets than ids
    return 1 << bits;
}

// slot_cap: valid tets expected (and the size of each of the two lists); key_cap: crossing-edge instances expected
bool x_layout(int64_t T, int64_t P, int64_t slot_cap, int64_t key_cap, XLayout *l) {
    l->slot_buckets = bucket_count(slot_cap, kMaxSlotBucketBits, T);
    l->key_buckets = bucket_count(key_cap, kMaxKeyBucketBits, P);
    l->slot_scale = T > 0 ? (((uint64_t)l->slot_buckets << 32) / (uint64_t)T) : 0;
    l->key_scale = P > 0 ? (((uint64_t)l->key_buckets << 32) / (uint64_t)P) : 0;
    size_t off = 0;
    l->header = off;        off += up256(sizeof(XHeader));
    l->cursors_slots = off; off += up256((size_t)l->slot_buckets * 4);
    l->cursors_keys = off;  off += up256((size_t)l->key_buckets * 4);
    l->groups_slots = off;  off += up256((size_t)(l->slot_buckets / kGroup + 1) * 8);
    l->groups_keys = off;   off += up256((size_t)(l->key_buckets / kGroup + 1) * 8);
    l->zeroed = off;
    l->counts_slots = off;  off += up256((size_t)l->slot_buckets * 8);
    l->counts_keys = off;   off += up256((size_t)l->key_buckets * 8);
    l->ubase = off;         off += up256((size_t)(l->key_buckets + 1) * 4);
    l->packed = off;        off += up256(P3D_MTX_PACK ? (size_t)P * 16 : 0);
    l->list1 = off;         off += up256((size_t)slot_cap * 8);
    l->list2 = off;         off += up256((size_t)slot_cap * 8);
    l->regions_slots = off; off += (size_t)l->slot_buckets * kBucketCap * 8;
    l->regions_keys = off;  off += (size_t)l->key_buckets * kBucketCap * 8;
    l->total = off;
    // more entries expected than the largest layout takes at a quarter of its capacity: such inputs belong to the
    // staged path (fewer buckets because there are few ids is no reason: a bucket that overflows says so itself)
    return key_cap <= ((int64_t)(kBucketCap / 4) << kMaxKeyBucketBits) && slot_cap <= ((int64_t)(kBucketCap / 4) << kMaxSlotBucketBits);
}

int64_t *x_pinned() {
    thread_local int64_t *buf = nullptr;
    if (!buf && cudaHostAlloc(reinterpret_cast<void **>(&buf), sizeof(XHeader), cudaHostAllocPortable) != cudaSuccess) buf = nullptr;
    return buf;
}

}  // namespace
}  // namespace p3d

using namespace p3d;

#define MTX_FAIL(st, msg) return p3d::set_error(st, msg)
#define MTX_CUDA(expr)                                                                                                    \
    do {                                                                                                                  \
        cudaError_t e_ = (expr);                                                                                          \
        if (e_ != cudaSuccess) MTX_FAIL(P3D_ERR_CUDA, std::string(#expr " failed: ") + cudaGetErrorString(e_));           \
    } while (0)

extern "C" {

size_t p3d_mt_extract_workspace_bytes(int64_t num_tets, int64_t num_points, int64_t slot_capacity, int64_t key_capacity) {
    if (num_tets < 0 || num_points < 0 || slot_capacity < 0 || key_capacity < 0) return 0;
    XLayout l;
    x_layout(num_tets, num_points, slot_capacity, key_capacity, &l);
    return l.total;
}

p3d_status p3d_mt_extract(const float *points, int64_t num_points, int64_t *tets, int64_t num_tets, const float *sdf,
                          int oriented, void *workspace, size_t workspace_bytes, int64_t slot_capacity, int64_t key_capacity,
                          float *verts, int64_t *edges, int64_t vertex_capacity, int64_t *faces, int64_t *tet_idx,
                          int64_t face_capacity, int64_t *counts_host, void *stream) {
    if (num_tets < 0 || num_points < 0 || slot_capacity < 0 || key_capacity < 0 || vertex_capacity < 0 || face_capacity < 0 || !counts_host)
        MTX_FAIL(P3D_ERR_INVALID, "p3d_mt_extract: invalid argument");
    for (int i = 0; i < 5; ++i) counts_host[i] = 0;
    if (num_tets == 0) return P3D_OK;
    if (!points || !tets || !sdf || !workspace) MTX_FAIL(P3D_ERR_INVALID, "p3d_mt_extract: null pointer");
    if ((vertex_capacity && !verts) || (face_capacity && !faces)) MTX_FAIL(P3D_ERR_INVALID, "p3d_mt_extract: null output with a non-zero capacity");
    if (num_tets > (int64_t)INT32_MAX || num_points > ((int64_t)1 << 32) - 1) MTX_FAIL(P3D_ERR_OVERFLOW, "p3d_mt_extract: more than 2^31 tets or 2^32 points");
    if (reinterpret_cast<uintptr_t>(tets) & 15) MTX_FAIL(P3D_ERR_INVALID, "p3d_mt_extract: tets must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(workspace) & 255) MTX_FAIL(P3D_ERR_INVALID, "p3d_mt_extract: workspace must be 256-byte aligned");
    XLayout l;
    const bool fits = x_layout(num_tets, num_points, slot_capacity, key_capacity, &l);
    if (fits && workspace_bytes < l.total) MTX_FAIL(P3D_ERR_WORKSPACE, "p3d_mt_extract: workspace too small");
    if (!fits) {
        counts_host[4] = 3;  // more valid tets / crossing edges expected than the bucket layout takes: staged calls
        return P3D_OK;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    char *base = static_cast<char *>(workspace);
    XHeader *hdr = reinterpret_cast<XHeader *>(base + l.header);
    XBuckets slots, keys;
    slots.cursors = reinterpret_cast<uint32_t *>(base + l.cursors_slots);
    slots.regions = reinterpret_cast<uint64_t *>(base + l.regions_slots);
    slots.counts = reinterpret_cast<unsigned long long *>(base + l.counts_slots);
    slots.group_sum = reinterpret_cast<unsigned long long *>(base + l.groups_slots);
    slots.scale = l.slot_scale;
    slots.count = l.slot_buckets;
    keys.cursors = reinterpret_cast<uint32_t *>(base + l.cursors_keys);
    keys.regions = reinterpret_cast<uint64_t *>(base + l.regions_keys);
    keys.counts = reinterpret_cast<unsigned long long *>(base + l.counts_keys);
    keys.group_sum = reinterpret_cast<unsigned long long *>(base + l.groups_keys);
    keys.scale = l.key_scale;
    keys.count = l.key_buckets;
    uint32_t *ubase = reinterpret_cast<uint32_t *>(base + l.ubase);
    uint64_t *list1 = reinterpret_cast<uint64_t *>(base + l.list1);
    uint64_t *list2 = reinterpret_cast<uint64_t *>(base + l.list2);

    MTX_CUDA(cudaMemsetAsync(base, 0, l.zeroed, s));
    const int sms = sm_count();
    {
        const int64_t want = (num_tets + 2 * kXThreads - 1) / (2 * kXThreads);
        const int64_t cap = (int64_t)sms * P3D_MTX_GRID;
        float4 *packed = reinterpret_cast<float4 *>(base + l.packed);
        if (P3D_MTX_PACK) k_mtx_pack<<<(unsigned)((num_points + kXThreads - 1) / kXThreads), kXThreads, 0, s>>>(points, sdf, num_points, packed);
        const unsigned grid = (unsigned)(want < cap ? want : cap);
        if ((reinterpret_cast<uintptr_t>(tets) & 31) == 0 && P3D_MTX_WIDE)
            k_mtx_classify<true><<<grid, kXThreads, 0, s>>>(points, (uint32_t)num_points, tets, num_tets, sdf, packed, hdr, slots, keys, oriented);
        else
            k_mtx_classify<false><<<grid, kXThreads, 0, s>>>(points, (uint32_t)num_points, tets, num_tets, sdf, packed, hdr, slots, keys, oriented);
    }
    const int jobs = l.slot_buckets + l.key_buckets;
    k_mtx_sort<<<jobs, kSortThreads, 0, s>>>(hdr, slots, keys);
    k_mtx_emit<<<jobs, kSortThreads, 0, s>>>(points, sdf, hdr, slots, keys, ubase, verts, edges, vertex_capacity, list1, list2, slot_capacity);
    if (face_capacity > 0)
        k_mtx_faces<<<(unsigned)((face_capacity + kXThreads - 1) / kXThreads), kXThreads, 0, s>>>(
            tets, hdr, list1, list2, slot_capacity, keys.regions, ubase, keys.scale, face_capacity, faces, tet_idx);
    MTX_CUDA(cudaGetLastError());
    int64_t *pin = x_pinned();
    XHeader host;
    void *dst = pin ? static_cast<void *>(pin) : static_cast<void *>(&host);
    MTX_CUDA(cudaMemcpyAsync(dst, hdr, sizeof(XHeader), cudaMemcpyDeviceToHost, s));
    MTX_CUDA(cudaStreamSynchronize(s));
    const XHeader h = *static_cast<const XHeader *>(dst);
    if (h.bad != 0) MTX_FAIL(P3D_ERR_INVALID, "p3d_mt_extract: " + std::to_string(h.bad) + " tets name a point outside [0, num_points)");
    const int64_t n1 = (int64_t)h.n1, n2 = (int64_t)h.n2;
    counts_host[0] = n1;
    counts_host[1] = n2;
    counts_host[2] = (int64_t)h.ne;
    counts_host[3] = (int64_t)h.num_unique;
    if (h.overflow) counts_host[4] = 2;
    else if (n1 > slot_capacity || n2 > slot_capacity || (int64_t)h.num_unique > vertex_capacity || n1 + 2 * n2 > face_capacity) counts_host[4] = 1;
    return P3D_OK;
}

}  // extern "C"
