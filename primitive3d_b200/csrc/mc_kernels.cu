// primitive3d_b200/csrc/mc_kernels.cu -- see mc_kernels.cuh for the design.
// Reference citations are into /root/reference/src/prim3d/Utility/marching_cubes.cu.
#include "mc_kernels.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "mc_case_table.h"
#include "scan_utils.cuh"

#ifndef P3D_TILE_PIN
#define P3D_TILE_PIN 0
#endif
#ifndef P3D_TILE_QUEUE
#define P3D_TILE_QUEUE 4
#endif
#ifndef P3D_TILE_RING
#define P3D_TILE_RING 1408
#endif


namespace p3d {

// Bourke case table, one packed word per case (nibble i = i-th edge index, nibble 15 =
// #triangles).  Lives in constant memory and is staged into shared memory once per
// persistent CTA because the per-cell lookups are lane-divergent.
__constant__ uint64_t c_case_table[256] = P3D_MC_CASE_TABLE_INIT;

// Per corner code in staging order a0 a1 b0 b1 c0 c1 d0 d1 (a = (x,y), b = (x+1,y), c = (x+1,y+1), d = (x,y+1); 0 = sample
// z, 1 = sample z+1): the case's correction  #triangles - (#crossed edges - 2), 0 for the two cases without triangles.
// Built at compile time from the case table (corner numbering of marching_cubes.cu:50-57, edges of :178-192).
struct NtriCorrection {
    int8_t v[256];
};
constexpr NtriCorrection make_ntri_correction() {
    constexpr uint64_t table[256] = P3D_MC_CASE_TABLE_INIT;
    constexpr int ea[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3}, eb[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
    NtriCorrection r{};
    for (unsigned c = 0; c < 256; ++c) {
        // staged bit0 = corner 0, bit1 = corner 4, bit2 = corner 1, bit3 = corner 5, bit4 = corner 2, bit5 = corner 6,
        // bit6 = corner 3, bit7 = corner 7
        const unsigned cs = (c & 1u) | ((c >> 1 & 1u) << 4) | ((c >> 2 & 1u) << 1) | ((c >> 3 & 1u) << 5) | ((c >> 4 & 1u) << 2) |
                            ((c >> 5 & 1u) << 6) | ((c >> 6 & 1u) << 3) | ((c >> 7 & 1u) << 7);
        const int nt = (int)(table[cs] >> 60);
        int ne = 0;
        for (int e = 0; e < 12; ++e) ne += (int)(((cs >> ea[e]) ^ (cs >> eb[e])) & 1u);
        r.v[c] = (int8_t)(ne ? nt - (ne - 2) : 0);
    }
    return r;
}
__constant__ NtriCorrection c_ntri_correction = make_ntri_correction();

// Bits of word w (32 samples from z = 32*w) with z + 1 < rz: samples that own a +z edge / a cell.
__device__ __forceinline__ uint32_t low_mask(int64_t n) {
    return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << (int)n) - 1u));
}

// One grid sample as float32: the conversion `density_grid.to(torch.float32)` performs
// (prim3d/utility/marching_cubes.py:86-87), round to nearest even.
__device__ __forceinline__ float sample_f32(float v) { return v; }
__device__ __forceinline__ float sample_f32(__half v) { return __half2float(v); }
__device__ __forceinline__ float sample_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float sample_f32(double v) { return __double2float_rn(v); }
__device__ __forceinline__ float sample_f32(long long v) { return __ll2float_rn(v); }
__device__ __forceinline__ float sample_f32(int v) { return __int2float_rn(v); }
__device__ __forceinline__ float sample_f32(short v) { return (float)v; }
__device__ __forceinline__ float sample_f32(unsigned char v) { return (float)v; }

// inside bits (value > thresh, marching_cubes.cu:25) of 32 consecutive staged samples, by ONE thread: eight
// 16-byte shared-memory reads, then a compare and a predicated OR per sample (NaN compares false, like the
// reference's `>`).  Four accumulators keep the OR chains short.
__device__ __forceinline__ uint32_t inside_word(const float *src, float thresh) {
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 f = *reinterpret_cast<const float4 *>(src + 4 * j);
        asm("{\n.reg .pred p;\nsetp.gt.f32 p, %1, %2;\n@p or.b32 %0, %0, %3;\n}" : "+r"(w0) : "f"(f.x), "f"(thresh), "r"(1u << (4 * j)));
        asm("{\n.reg .pred p;\nsetp.gt.f32 p, %1, %2;\n@p or.b32 %0, %0, %3;\n}" : "+r"(w1) : "f"(f.y), "f"(thresh), "r"(2u << (4 * j)));
        asm("{\n.reg .pred p;\nsetp.gt.f32 p, %1, %2;\n@p or.b32 %0, %0, %3;\n}" : "+r"(w2) : "f"(f.z), "f"(thresh), "r"(4u << (4 * j)));
        asm("{\n.reg .pred p;\nsetp.gt.f32 p, %1, %2;\n@p or.b32 %0, %0, %3;\n}" : "+r"(w3) : "f"(f.w), "f"(thresh), "r"(8u << (4 * j)));
    }
    return (w0 | w1) | (w2 | w3);
}

// ---------------------------------------------------------------------------------------------
// TMA / mbarrier plumbing (raw PTX; sm_100a).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// 3-D tiled load global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// Two-level single-pass scan (no CUB / thrust), used for the first vertex id of a tile (k_tile) and the
// first face index of a chunk (k_faces).
//
// Items are grouped into rounds of 256 consecutive ids.  An item publishes its count in its status
// word and adds it to its round's accumulator; whichever item completes a round (the 256th arrival)
// publishes the exclusive prefix of the NEXT round.  An item's exclusive prefix is then
//     prefix[round] + sum of the status words of the items before it in its round
// -- one batch of <= 8 independent loads per lane, however many items are in flight (with ~600
// resident tiles a classic 32-wide look-back walks ~20 dependent windows).  Every wait is on an item
// with a smaller id; ids are handed out in increasing order to running CTAs / warps, which publish
// without waiting for anyone, so the waits end.  The result does not depend on arrival order:
// numbering is deterministic.
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long kPublished = 1ull << 63;
constexpr unsigned long long kRoundSumMask = (1ull << 48) - 1;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    *reinterpret_cast<volatile unsigned long long *>(p) = v;
}

// One lane of the owner of `item`: publish its count (< 2^39, so that a round's sum stays below 2^48).
__device__ __forceinline__ void scan_publish(const McScan &sc, uint32_t item, unsigned long long count, uint32_t nitems) {
    const uint32_t k = item / kRoundTiles;
    st_status(sc.status + item, kPublished | count);
    const uint32_t left = nitems - k * kRoundTiles, members = left < (uint32_t)kRoundTiles ? left : (uint32_t)kRoundTiles;
    const unsigned long long old = atomicAdd(sc.round_acc + k, (1ull << 48) | count);
    if ((uint32_t)(old >> 48) == members - 1) {  // this item completes round k
        unsigned long long base = kPublished;
        if (k) do base = ld_status(sc.round_prefix + k); while (!(base & kPublished));
        st_status(sc.round_prefix + k + 1, base + (old & kRoundSumMask) + count);
    }
}

// One full warp: exclusive prefix of `item`.  Blocking: waits for the items before it in its round and for the
// round's prefix.  Non-blocking: returns false if any of them is not published yet.  Also valid as a pure
// re-read after the scan has completed (vertices-only pass).
template <bool BLOCK>
__device__ __forceinline__ bool scan_prefix(const McScan &sc, uint32_t item, int lane, unsigned long long &first) {
    const uint32_t k = item / kRoundTiles, j = item % kRoundTiles;
    unsigned long long s[kRoundTiles / 32];
    const unsigned long long *st = sc.status + (item - j);
#pragma unroll
    for (int i = 0; i < (int)kRoundTiles / 32; ++i) s[i] = (uint32_t)(lane + 32 * i) < j ? ld_status(st + lane + 32 * i) : kPublished;
    unsigned long long acc = kPublished, ready = kPublished;
    if (lane == 0 && k) {
        acc = ld_status(sc.round_prefix + k);
        if (BLOCK) while (!(acc & kPublished)) acc = ld_status(sc.round_prefix + k);
        ready = acc;
    }
#pragma unroll
    for (int i = 0; i < (int)kRoundTiles / 32; ++i) {
        if (BLOCK) while (!(s[i] & kPublished)) s[i] = ld_status(st + lane + 32 * i);
        ready &= s[i];
        acc += s[i] & ~kPublished;
    }
    if (!BLOCK && !__all_sync(kFull, (ready & kPublished) != 0ull)) return false;
    first = warp_sum64(acc & ~kPublished);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Pass A: k_tile.  Four CTAs per SM, each with ONE staged tile: while a CTA waits for its TMA load or
// for its first vertex id, the other three compute.
//
// Shared memory of a CTA:
//   stage      fp32 samples of a tile: rows (xi, yi) of 0..8 x 0..8, kBoxZ samples each (TMA box)
//   sbits      [81][8] words: inside bits of every staged row (word 4, bit 0 = the halo sample)
//   piece, nfp [64] vertex / triangle count of each owned (row, piece)
//   ent, dt    ring of the pending crossing edges (axis<<13 | row<<7 | z) and their interpolation parameters:
//              the tiles whose first vertex id is not known yet (a warp's 8 rows are a contiguous range)
//   q, prel    the pending tiles and their table entries, relative to the tile
// A thread owns bit word (row r = tid>>2, word w = tid&3) of the tile in the count phase; a warp owns the 8
// rows of one plane, whose vertices are a contiguous id range of the tile.
//
// Per tile: bits -> counts -> tile scan -> publish count -> per warp: compact edges, interpolate dt from the
// staged samples -> the stage is free: the NEXT tile's TMA load is issued -> wait for the tile's first
// vertex id (asked for without blocking, one iteration later) -> write vertices (position = integer corner + dt
// on one axis) and the table entries.  The load of the next tile and the wait for the scan overlap.
// ---------------------------------------------------------------------------------------------
constexpr int kRing = P3D_TILE_RING;  // pending crossing edges of a CTA (6 bytes each): ~4.5 tiles of the gyroid case
constexpr int kQueue = P3D_TILE_QUEUE;  // pending tiles of a CTA
constexpr int kSbitsStride = 8;
constexpr int kRowPitch = kTileY + 1;  // staged rows per plane
// position in the ring of entry `pos` (< 2 * kRing): a mask when the ring is a power of two
__device__ __forceinline__ uint32_t ring_wrap(uint32_t pos) {
    if ((kRing & (kRing - 1)) == 0) return pos & (uint32_t)(kRing - 1);
    return pos >= (uint32_t)kRing ? pos - (uint32_t)kRing : pos;
}

struct TileCoord {          // a tile and where its side products go
    int x0, y0, p, tile;
    long long piece_base;   // index of the (row, piece) pair (x0, y0, p) in the per-piece arrays; its bit words start at 4 x that
    long long pad;
};

struct PendingTile {
    int x0, y0, p, tile;
    uint32_t start, count;  // its entries in the ring
    long long piece_base;
};

struct TileSmem {
    uint32_t sbits[kBoxRows * kSbitsStride];
    uint32_t piece[kTileX * kTileY];
    uint16_t nfp[kTileX * kTileY];  // triangles of each owned (row, piece)
    float dt[kRing];       // pending vertices: interpolation parameter ...
    uint16_t ent[kRing];   // ... and edge (axis<<13 | row<<7 | z)
    int8_t ntri[256];      // indexed by the corner bits in staging order a0 a1 b0 b1 c0 c1 d0 d1:
                           // the case's correction  #triangles - (#crossed edges - 2)
    PendingTile q[kQueue];
    uint2 prel[kQueue][kTileX * kTileY];  // table entries of the pending tiles, relative to the tile: {vx | vy << 16, vz | nf << 16}
    unsigned long long bar;
    unsigned long long base;       // result of warp 0's non-blocking look-back at the top of an iteration
    unsigned long long base_wait;  // result of a blocking look-back
    uint32_t base_ok;
    TileCoord coord[2];    // the tile of iteration it, by parity
};
constexpr int kTileSmemBytes = kStageBytes + (int)sizeof(TileSmem) + 128;
static_assert(4 * (kTileSmemBytes + 1024) <= 228 * 1024, "k_tile is tuned for four CTAs per SM");

// (1 << n) - 1 for n in [0, 32]; PTX shl.b32 gives 0 for a shift of 32 (the C++ operator is undefined there)
__device__ __forceinline__ uint32_t low_mask_clamped(int n) {
    uint32_t r;
    const int m = n < 0 ? 0 : (n > 32 ? 32 : n);
    asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(m));
    return r - 1u;
}

// SPARSE: the block-sparse form (g.tile_list names the tiles to visit); a template parameter so that the dense
// instances carry none of it
template <bool TMA, typename T, bool SPARSE>
__global__ void __launch_bounds__(kTileThreads, 4)
    k_tile(const __grid_constant__ CUtensorMap tmap, const T *__restrict__ grid, McGeom g, McWorkspace ws,
           McEmitParams prm, float *__restrict__ verts, unsigned long long vcap, int mode) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    float *tf = reinterpret_cast<float *>(sm);
    TileSmem &S = *reinterpret_cast<TileSmem *>(sm + kStageBytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ntiles = (uint32_t)g.ntiles;
    const int rx = (int)g.rx, ry = (int)g.ry, rz = (int)g.rz, ox = (int)g.owned_x, np = g.np;
    const float thresh = prm.thresh;
    const int xg0 = (int)prm.x_origin;  // global index of local plane 0 (< 2^31)

    S.ntri[tid] = c_ntri_correction.v[tid];

    // Tile id -> coordinates.  Tiles are ordered band by band (a band = `band` y-blocks over all x), inside a
    // band x-block major, then y-block, then piece: the x halo plane of a block is re-read from L2, not HBM.
    auto locate = [&](uint32_t t) {
        TileCoord c;
        c.x0 = c.y0 = c.p = 0, c.tile = (int)t, c.piece_base = 0, c.pad = 0;
        if (SPARSE && t < ntiles) {
            // block-sparse form: the caller's tile id, x-block major; an id past the grid is a tile without samples
            const uint32_t id = __ldg(g.tile_list + t);
            const uint32_t per_x = (uint32_t)g.nyb * (uint32_t)np;
            const uint32_t xb = id / per_x, rem = id - xb * per_x;
            const uint32_t yb = rem / np, p = rem - yb * np;
            c.x0 = (int)(xb < (uint32_t)g.nxb ? xb * kTileX : (uint32_t)g.nxb * kTileX);
            c.y0 = (int)(yb * kTileY), c.p = (int)p;
            c.piece_base = ((long long)c.x0 * ry + c.y0) * np + c.p;
        } else if (t < ntiles) {
            const uint32_t per_band = (uint32_t)g.nxb * (uint32_t)g.band * (uint32_t)np;
            const uint32_t bi = t / per_band, rem = t - bi * per_band;
            const uint32_t left = (uint32_t)g.nyb - bi * (uint32_t)g.band, cur = left < (uint32_t)g.band ? left : (uint32_t)g.band;
            const uint32_t xb = rem / (cur * np), rem2 = rem - xb * (cur * np);
            const uint32_t yb = rem2 / np, p = rem2 - yb * np;
            c.x0 = (int)(xb * kTileX), c.y0 = (int)((bi * g.band + yb) * kTileY), c.p = (int)p;
            c.piece_base = ((long long)c.x0 * ry + c.y0) * np + c.p;
        }
        return c;
    };
    auto issue = [&](const TileCoord &c) {  // one thread, once the stage is free
        if (TMA && (uint32_t)c.tile < ntiles) {
            const uint32_t bar = smem_u32(&S.bar);
            mbar_expect_tx(bar, kBoxRows * kBoxZ * 4);
            tma_load_3d(smem_u32(tf), &tmap, bar, c.p * kTileZ, c.y0, c.x0);
        }
    };
    if (tid == 0) {
        if (TMA) {
            mbar_init(smem_u32(&S.bar), 1);
            mbar_fence_init();
        }
        const TileCoord c = locate(atomicAdd(&ws.header->ticket, 1u));
        S.coord[0] = c;
        issue(c);
    }
    __syncthreads();

    // my word of the tile (count phase): row r = (xi, yi), word w
    const int r = tid >> 2, w = tid & 3;
    const int xi = r >> 3, yi = r & 7;
    int sa_index = (xi * kRowPitch + yi) * kSbitsStride + w;
    // my (row, piece) pair relative to the tile's first, in the per-piece arrays
    long long my_piece_off = ((long long)xi * ry + yi) * np;
    int zlim = rz - 1 - 32 * w;  // samples of my word with z + 1 < rz: zlim - z0 of them
#if P3D_TILE_PIN
    // per-thread constants of the tile loop: opaque to the compiler, which otherwise recomputes them from threadIdx
    // for every tile to stay under a register count it does not need to stay under
    asm volatile("" : "+r"(sa_index), "+l"(my_piece_off), "+r"(zlim));
#endif
    const uint32_t *sa = &S.sbits[sa_index];

    // one crossing edge -> its vertex (gen_vertices_kernel :70-138 and the epilogue :298)
    auto edge_dt = [&](uint32_t ent) {
        const uint32_t ax = ent >> 13, er = (ent >> 7) & 63u, ez = ent & 127u;
        const float *src = tf + ((er >> 3) * kRowPitch + (er & 7u)) * kBoxZ + ez;
        const float d0 = src[0];
        const float d1 = src[ax == 0 ? kRowPitch * kBoxZ : (ax == 1 ? kBoxZ : 1)];
        // dt = (thresh - d_self) / (d_next - d_self), IEEE fp32, no contraction (:105)
        return __fdiv_rn(__fsub_rn(thresh, d0), __fsub_rn(d1, d0));
    };
    // position = float(voxel) + dt on the edge's axis (:107; adding +0.0f to the other two changes nothing: they are
    // >= 0), then vertices * scale + offset as two separately rounded ops (:298)
    auto put_vertex = [&](float *out, uint32_t ent, float dt, int x0g, int y0, int z0) {
        const uint32_t ax = ent >> 13;
        const float px = __fadd_rn((float)(x0g + (int)((ent >> 10) & 7u)), ax == 0 ? dt : 0.0f);
        const float py = __fadd_rn((float)(y0 + (int)((ent >> 7) & 7u)), ax == 1 ? dt : 0.0f);
        const float pz = __fadd_rn((float)(z0 + (int)(ent & 127u)), ax == 2 ? dt : 0.0f);
        out[0] = __fadd_rn(__fmul_rn(px, prm.scale[0]), prm.offset[0]);
        out[1] = __fadd_rn(__fmul_rn(py, prm.scale[1]), prm.offset[1]);
        out[2] = __fadd_rn(__fmul_rn(pz, prm.scale[2]), prm.offset[2]);
    };

    // Emission of a tile is deferred: its first vertex id needs the counts of every tile before it.  The
    // crossing edges of a counted tile (edge code + interpolation parameter, 6 bytes) wait in a ring in shared
    // memory; the oldest pending tile is retired when a non-blocking look-back finds its first id, or -- only
    // when the ring is full -- after a blocking one.  A CTA therefore (almost) never waits for another CTA.
    // Queue state, identical in every thread:
    uint32_t q_head = 0, q_count = 0, ring_used = 0, ring_tail = 0;

    // The table entries of a tile {first x-/y-/z-edge vertex id, triangle count} per (row, piece) are known relative
    // to the tile in the count phase; they wait in shared memory with the tile's vertices and are written once,
    // absolute, when the tile's first vertex id is known: the face pass needs no per-tile indirection and the
    // table is never read back.  Thread `row` (< 64) writes the entry of row `row`.
    auto write_entries = [&](int x0, int y0, long long piece_base, uint32_t tile, const uint2 *rel, unsigned long long base,
                             uint32_t count) {
        if (mode == 0 && tid < kTileX * kTileY) {
            const int ex = tid >> 3, ey = tid & 7;
            if (x0 + ex < ox && y0 + ey < ry) {
                const uint2 e = rel[tid];
                ws.ptab[piece_base + ((long long)ex * ry + ey) * np] =
                    make_uint4((e.x & 0xffffu) + (uint32_t)base, (e.x >> 16) + (uint32_t)base, (e.y & 0xffffu) + (uint32_t)base, e.y >> 16);
            }
        }
        if (mode == 0 && tid == 0 && tile == ntiles - 1) ws.header->total_v = base + count;
    };
    // vertices of the pending tile in queue slot `slot`, whose first vertex id is `base`
    auto retire = [&](uint32_t slot, unsigned long long base) {
        const PendingTile &q = S.q[slot];
        const uint32_t start = q.start, count = q.count;
        const int x0g = xg0 + q.x0, y0 = q.y0, z0 = q.p * kTileZ;
        write_entries(q.x0, y0, q.piece_base, (uint32_t)q.tile, S.prel[slot], base, count);
        float *const out0 = verts + base * 3ull;
        if (base + count <= vcap) {
            for (uint32_t k = tid; k < count; k += kTileThreads) {
                const uint32_t idx = ring_wrap(start + k);
                put_vertex(out0 + 3u * k, S.ent[idx], S.dt[idx], x0g, y0, z0);
            }
        } else {  // the speculative buffer ends inside this tile
            for (uint32_t k = tid; k < count; k += kTileThreads) {
                const uint32_t idx = ring_wrap(start + k);
                if (base + k < vcap) put_vertex(out0 + 3u * k, S.ent[idx], S.dt[idx], x0g, y0, z0);
            }
        }
        ring_used -= count;
        q_head = (q_head + 1) % kQueue;
        --q_count;
    };
    // blocking look-back for `tile` by warp 0, result to every thread (two barriers)
    auto wait_base = [&](uint32_t t) {
        if (warp == 0) {
            unsigned long long tb = 0;
            scan_prefix<true>(ws.vscan, t, lane, tb);
            if (lane == 0) S.base_wait = tb;
        }
        __syncthreads();
        const unsigned long long tb = S.base_wait;
        __syncthreads();
        return tb;
    };

    for (uint32_t it = 0;; ++it) {
        const TileCoord &tc = S.coord[it & 1u];
        const uint32_t tile = (uint32_t)tc.tile;
        if (tile >= ntiles) break;
        const int x0 = tc.x0, y0 = tc.y0, p = tc.p, z0 = p * kTileZ;
        const long long piece_base = tc.piece_base;

        // ticket of this CTA's next tile, asked for a whole tile ahead of its use.  Tiles are handed out in
        // increasing order to running CTAs: every tile before a tile is finished or in flight, whatever the
        // residency of the grid, and a CTA that got a light tile simply takes the next one sooner.
        uint32_t next_tile = 0;
        if (tid == 32) next_tile = atomicAdd(&ws.header->ticket, 1u);

        // first vertex id of the oldest pending tile, asked for while this tile's load is in flight: the round
        // trip to L2 hides behind the TMA wait.  Non-blocking.
        if (q_count && warp == 0) {
            unsigned long long tb = 0;
            const bool ok = scan_prefix<false>(ws.vscan, (uint32_t)S.q[q_head].tile, lane, tb);
            if (lane == 0) {
                S.base_ok = ok ? 1u : 0u;
                S.base = tb;
            }
        }

        if (TMA) {
            mbar_wait(smem_u32(&S.bar), it & 1u);
        } else {
            // a warp per staged row, consecutive lanes on consecutive samples; converted to float32 here
            for (int rr = warp; rr < kBoxRows; rr += kTileThreads / 32) {
                const int bx = rr / kRowPitch, by = rr - bx * kRowPitch;
                const int gx = x0 + bx, gy = y0 + by;
                const bool row_ok = gx < rx && gy < ry;
                const T *src = grid + ((int64_t)(row_ok ? gx : 0) * ry + (row_ok ? gy : 0)) * rz;
                for (int c = lane; c < kBoxZ; c += 32) {
                    const int gz = z0 + c;
                    tf[rr * kBoxZ + c] = (row_ok && gz < rz) ? sample_f32(__ldg(src + gz)) : 0.0f;
                }
            }
            __syncthreads();
        }

        // ---- phase 1: inside bits of the 81 staged rows.  inside = value > thresh (:25,31,37,43,50-57).
        // Samples outside the grid are staged as 0.0f; their bits take part in no mask that is not cut by a
        // validity test below, so they need no cleaning here. ----
        {
            // a thread per (staged row, word): the eight lanes of a quarter warp read eight consecutive rows at the
            // same word, i.e. eight different 16-byte bank groups (a staged row is 33 x 16 bytes)
            const int brow = warp * 8 + (lane & 7), bword = lane >> 3;
            S.sbits[brow * kSbitsStride + bword] = inside_word(tf + brow * kBoxZ + 32 * bword, thresh);
            if (warp < 3) {  // staged rows 64 .. 80
                const int row2 = 64 + brow;
                if (row2 < kBoxRows) S.sbits[row2 * kSbitsStride + bword] = inside_word(tf + row2 * kBoxZ + 32 * bword, thresh);
            } else if (warp < 6) {  // the halo sample (z0 + 128) of every staged row: a lane per row
                const int row = (warp - 3) * 32 + lane;
                if (row < kBoxRows) S.sbits[row * kSbitsStride + 4] = tf[row * kBoxZ + kTileZ] > thresh ? 1u : 0u;
            }
        }
        __syncthreads();  // [bits]

        // ---- phase 2: crossing masks and counts of my word ----
        const int x = x0 + xi, y = y0 + yi;
        const bool own = (x < ox) && (y < ry);
        const bool hx = own && (x + 1 < rx), hy = own && (y + 1 < ry), hc = hx && hy;
        const uint32_t A = sa[0], An = sa[1];
        const uint32_t B = sa[kRowPitch * kSbitsStride], Bn = sa[kRowPitch * kSbitsStride + 1];
        const uint32_t D = sa[kSbitsStride], Dn = sa[kSbitsStride + 1];
        const uint32_t C = sa[(kRowPitch + 1) * kSbitsStride], Cn = sa[(kRowPitch + 1) * kSbitsStride + 1];
        const uint32_t A2 = __funnelshift_r(A, An, 1);
        const uint32_t zv = low_mask_clamped(zlim - z0);  // samples with z + 1 < rz
        const uint32_t m0 = hx ? (A ^ B) : 0u;            // +x edges, :29-33 / :100-111
        const uint32_t m1 = hy ? (A ^ D) : 0u;            // +y edges, :35-39 / :113-124
        const uint32_t m2 = own ? ((A ^ A2) & zv) : 0u;   // +z edges, :41-45 / :126-137
        uint32_t nf = 0;
        // Triangles of my 32 cells without walking them.  Every row of the case table triangulates the closed
        // loops the surface cuts through the cell, a loop over k crossed edges into k - 2 triangles, so a cell
        // with E crossed edges and L loops has E - 2 L triangles (checked for all 256 cases, tests/test_tables.py).
        // E summed over the word is twelve popcounts; L = 1 unless the inside or the outside corners fall apart,
        // which needs an ambiguous face (all four edges of a face crossed) or two isolated opposite corners:
        // only those cells are looked up, for their correction  #triangles - (E - 2).
        if (hc && mode == 0) {                            // cells :48-66; bit i = cell at sample z0 + 32 w + i
            const uint32_t B2 = __funnelshift_r(B, Bn, 1), C2 = __funnelshift_r(C, Cn, 1), D2 = __funnelshift_r(D, Dn, 1);
            const uint32_t xa0 = (A ^ B) & zv, xa1 = (A2 ^ B2) & zv, xd0 = (D ^ C) & zv, xd1 = (D2 ^ C2) & zv;  // x edges
            const uint32_t ya0 = (A ^ D) & zv, ya1 = (A2 ^ D2) & zv, yb0 = (B ^ C) & zv, yb1 = (B2 ^ C2) & zv;  // y edges
            const uint32_t za = (A ^ A2) & zv, zb = (B ^ B2) & zv, zc = (C ^ C2) & zv, zd = (D ^ D2) & zv;      // z edges
            const uint32_t act = xa0 | xa1 | xd0 | xd1 | ya0 | ya1 | yb0 | yb1 | za | zb | zc | zd;  // mixed corners
            const int edges = __popc(xa0) + __popc(xa1) + __popc(xd0) + __popc(xd1) + __popc(ya0) + __popc(ya1) +
                              __popc(yb0) + __popc(yb1) + __popc(za) + __popc(zb) + __popc(zc) + __popc(zd);
            // three crossed edges of a face imply the fourth (parity around the face)
            const uint32_t amb = (xa0 & yb0 & xd0) | (xa1 & yb1 & xd1) | (xa0 & zb & xa1) | (yb0 & zc & yb1) |
                                 (xd0 & zd & xd1) | (ya0 & za & ya1);
            const uint32_t pac = za & zc, pbd = zb & zd;
            const uint32_t diag = (pac & ((xa0 & ya0 & yb1 & xd1) | (yb0 & xd0 & xa1 & ya1))) |   // a0 + c1, c0 + a1 isolated
                                  (pbd & ((xa0 & yb0 & xd1 & ya1) | (xd0 & ya0 & xa1 & yb1)));    // b0 + d1, d0 + b1 isolated
            int total = edges - 2 * __popc(act);
            for (uint32_t rem = amb | diag; rem;) {
                const int i = __ffs(rem) - 1;
                rem &= rem - 1;
                const uint32_t code = (__funnelshift_r(A, An, i) & 3u) | ((__funnelshift_r(B, Bn, i) & 3u) << 2) |
                                      ((__funnelshift_r(C, Cn, i) & 3u) << 4) | ((__funnelshift_r(D, Dn, i) & 3u) << 6);
                total += S.ntri[code];
            }
            nf = (uint32_t)total;
        }
        // my (row, piece) = 4 adjacent lanes: packed {nx, ny, nz} (8-bit fields, <= 128 each)
        const uint32_t cx = (uint32_t)__popc(m0), cy = (uint32_t)__popc(m1), cz = (uint32_t)__popc(m2);
        const uint32_t cnt = cx | (cy << 8) | (cz << 16);
        uint32_t inc = cnt;
        {
            uint32_t t = __shfl_up_sync(kFull, inc, 1);
            if (w >= 1) inc += t;
            t = __shfl_up_sync(kFull, inc, 2);
            if (w >= 2) inc += t;
        }
        const uint32_t tot = __shfl_sync(kFull, inc, lane | 3);
        const uint32_t exw = inc - cnt;
        // triangles of the four words of my (row, piece), one byte each (<= 160): what the face pass reads per word
        uint32_t nfw = nf << (8 * w);
        nfw += __shfl_xor_sync(kFull, nfw, 1);
        nfw += __shfl_xor_sync(kFull, nfw, 2);
        nf = __dp4a(nfw, 0x01010101u, 0u);  // of the piece
        if (w == 0) {
            S.piece[r] = __dp4a(tot, 0x01010101u, 0u);
            S.nfp[r] = (uint16_t)nf;  // <= 640
        }

        if (mode == 0) {
            const long long mine = piece_base + my_piece_off;  // my (row, piece) in the per-piece arrays
            if (x < rx && y < ry) ws.bits[4 * mine + w] = A;
            if (own && w == 0) ws.nf[mine] = nfw;
            // the halo plane of a slab sits one past the last x-block when owned_x is a multiple of 8
            if (tid < 32 && x0 + kTileX == ox && ox < rx && y0 + (tid >> 2) < ry)
                ws.bits[((int64_t)ox * ry + y0 + (tid >> 2)) * (4 * (int64_t)np) + 4 * p + (tid & 3)] =
                    S.sbits[(kTileX * kRowPitch + (tid >> 2)) * kSbitsStride + (tid & 3)];
            if (SPARSE) {
                // block-sparse form: the tiles next to this one may not be visited, yet the cells on this tile's +x /
                // +y / +z faces read their first rows / samples.  OR the staged halo bits into the (zeroed) bit words:
                // a neighbour that is visited writes the same bits and more.
                if (tid < 4 * 17) {                      // halo rows: xi = 8 (nine of them), yi = 8 (eight more), 4 words
                    const int h = tid >> 2, hw = tid & 3;
                    const int hx_ = h < 9 ? kTileX : h - 9, hy_ = h < 9 ? h : kTileY;
                    const int gx = x0 + hx_, gy = y0 + hy_;
                    const uint32_t v = S.sbits[(hx_ * kRowPitch + hy_) * kSbitsStride + hw];
                    if (gx < rx && gy < ry && v) atomicOr(ws.bits + ((int64_t)gx * ry + gy) * (4 * (int64_t)np) + 4 * p + hw, v);
                } else if (tid >= 128 && tid < 128 + kBoxRows && p + 1 < np) {   // sample z0 + 128 of every staged row
                    const int h = tid - 128, hx_ = h / kRowPitch, hy_ = h - hx_ * kRowPitch;
                    const int gx = x0 + hx_, gy = y0 + hy_;
                    if (gx < rx && gy < ry && (S.sbits[h * kSbitsStride + 4] & 1u))
                        atomicOr(ws.bits + ((int64_t)gx * ry + gy) * (4 * (int64_t)np) + 4 * (p + 1), 1u);
                }
            }
        }
        // the look-back result of the top of this iteration: written by warp 0 before [bits], read here, and not
        // overwritten before warp 0 has passed [count]
        bool probe_ok = q_count && S.base_ok;
        const unsigned long long probe_base = S.base;
        __syncthreads();  // [count]

        // ---- tile scan (every warp redundantly): first vertex of each (row, piece), relative to the tile ----
        uint32_t vt, pe, wbase, wcount;
        {
            const uint2 v = *reinterpret_cast<const uint2 *>(&S.piece[2 * lane]);
            const uint32_t s2 = v.x + v.y, incl = warp_incl_scan(s2, lane);
            vt = __shfl_sync(kFull, incl, 31);
            const uint32_t e0 = incl - s2;
            const uint32_t g0 = __shfl_sync(kFull, e0, r >> 1), g1 = __shfl_sync(kFull, v.x, r >> 1);
            pe = g0 + ((r & 1) ? g1 : 0u);
            // vertices of warp k's 8 rows: [e0 of lane 4k, e0 of lane 4k+4)
            const uint32_t nxt = __shfl_down_sync(kFull, e0, 4);
            const uint32_t seg = ((lane & 3) == 0) ? (lane == 28 ? vt : nxt) - e0 : 0u;
            wbase = __shfl_sync(kFull, e0, 4 * warp);
            wcount = __shfl_sync(kFull, seg, 4 * warp);
        }
        // face offsets of the face pass: triangle counts summed per chunk of 128 consecutive (row, piece) pairs.
        // Warp 7, a lane per x-row of the tile (its 8 rows are one RED when they fall into one chunk).
        if (mode == 0 && warp == 7) {
            const uint32_t v2 = *reinterpret_cast<const uint32_t *>(&S.nfp[2 * lane]);
            uint32_t t = (v2 & 0xffffu) + (v2 >> 16);
            t += __shfl_xor_sync(kFull, t, 1);
            t += __shfl_xor_sync(kFull, t, 2);  // triangles of x-row lane >> 2
            uint32_t tile_nf = t;
            tile_nf += __shfl_xor_sync(kFull, tile_nf, 4);
            tile_nf += __shfl_xor_sync(kFull, tile_nf, 8);
            tile_nf += __shfl_xor_sync(kFull, tile_nf, 16);
            if (lane == 0 && tile_nf) atomicAdd(&ws.header->total_f, (unsigned long long)tile_nf);  // F = sum of the table-row lengths, :66
            if ((lane & 3) == 0 && t) {
                constexpr int64_t kChunkPieces = kFacePieces * kFaceChunk;
                const int xr = lane >> 2, ylast = (y0 + kTileY <= ry ? y0 + kTileY : ry) - 1;
                const int64_t gfirst = piece_base + (int64_t)xr * ry * np, glast = gfirst + (int64_t)(ylast - y0) * np;
                if (gfirst / kChunkPieces == glast / kChunkPieces) {
                    atomicAdd(ws.chunk_sum + gfirst / kChunkPieces, t);
                } else {
                    for (int k = 0; k < kTileY; ++k) {
                        const uint32_t nk = S.nfp[xr * kTileY + k];
                        if (nk) atomicAdd(ws.chunk_sum + (gfirst + (int64_t)k * np) / kChunkPieces, nk);
                    }
                }
            }
        }
        if (mode == 0 && tid == 0) {
            scan_publish(ws.vscan, tile, vt, ntiles);
        }
        const uint32_t vx_rel = pe, vy_rel = pe + (tot & 255u), vz_rel = vy_rel + ((tot >> 8) & 255u);
        // table entry of my (row, piece), relative to the tile (every field < 2^16: a tile has <= 24576 vertices)
        const uint2 my_entry = make_uint2(vx_rel | (vy_rel << 16), vz_rel | (nf << 16));
        // one past the last entry of my word's crossings, per axis (relative to the tile)
        const uint32_t last[3] = {vx_rel + (exw & 255u) + cx, vy_rel + ((exw >> 8) & 255u) + cy, vz_rel + (exw >> 16) + cz};
        const uint32_t wmask[3] = {m0, m1, m2};
        const uint32_t ecode = (uint32_t)((r << 7) | (w << 5));
        if (vt <= (uint32_t)kRing) {
            // ---- make room in the ring (rare): retire the oldest pending tiles, waiting for their ids ----
            while (q_count && (ring_used + vt > (uint32_t)kRing || q_count == (uint32_t)kQueue)) {
                unsigned long long tb = probe_base;
                if (!probe_ok) tb = wait_base((uint32_t)S.q[q_head].tile);
                probe_ok = false;
                retire(q_head, tb);
                __syncthreads();  // its entries may be overwritten now
            }
            // ---- my crossing edges -> ring (last to first); my warp interpolates its own (contiguous) range ----
            const uint32_t start = ring_tail;
            if (w == 0) S.prel[(q_head + q_count) % kQueue][r] = my_entry;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                uint32_t pos = start + last[ax];
                const uint32_t code = (uint32_t)(ax << 13) | ecode;
                for (uint32_t rem = wmask[ax]; rem;) {
                    const int i = 31 - __clz(rem);
                    rem ^= 1u << i;
                    --pos;
                    S.ent[ring_wrap(pos)] = (uint16_t)(code | i);
                }
            }
            __syncwarp();
            for (uint32_t k = lane; k < wcount; k += 32) {
                const uint32_t idx = ring_wrap(start + wbase + k);
                S.dt[idx] = edge_dt(S.ent[idx]);
            }
            if (tid == 0) {
                PendingTile &q = S.q[(q_head + q_count) % kQueue];
                q.x0 = x0, q.y0 = y0, q.p = p, q.tile = (int)tile;
                q.start = start;
                q.count = vt;
                q.piece_base = piece_base;
            }
            ++q_count;
            ring_used += vt;
            ring_tail = ring_wrap(start + vt);
        } else {
            // ---- more crossings than the ring holds (noise-like data): retire everything pending, wait for
            // this tile's first id, and emit it chunk by chunk from the stage ----
            while (q_count) {
                unsigned long long tb = probe_base;
                if (!probe_ok) tb = wait_base((uint32_t)S.q[q_head].tile);
                probe_ok = false;
                retire(q_head, tb);
            }
            const unsigned long long tb = wait_base(tile);
            // the entries go through the (idle) first queue slot: thread `row` writes the entry of row `row`
            if (w == 0) S.prel[0][r] = my_entry;
            __syncthreads();
            write_entries(x0, y0, piece_base, tile, S.prel[0], tb, vt);
            for (uint32_t c0 = 0; c0 < vt; c0 += kRing) {
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    uint32_t pos = last[ax] - c0;  // wraps for entries outside the chunk: filtered by the range test
                    const uint32_t code = (uint32_t)(ax << 13) | ecode;
                    for (uint32_t rem = wmask[ax]; rem;) {
                        const int i = 31 - __clz(rem);
                        rem ^= 1u << i;
                        --pos;
                        if (pos < (uint32_t)kRing) S.ent[pos] = (uint16_t)(code | i);
                    }
                }
                __syncthreads();
                const uint32_t n = vt - c0 < (uint32_t)kRing ? vt - c0 : (uint32_t)kRing;
                for (uint32_t k = tid; k < n; k += kTileThreads) {
                    const unsigned long long id = tb + c0 + k;
                    const uint32_t ent = S.ent[k];
                    if (id < vcap) put_vertex(verts + id * 3ull, ent, edge_dt(ent), xg0 + x0, y0, z0);
                }
                __syncthreads();
            }
            ring_tail = 0;
        }
        // the next tile's coordinates are published before the barrier, its load is issued after it
        if (tid == 32) S.coord[(it + 1u) & 1u] = locate(next_tile);
        __syncthreads();  // [stage free]
        if (tid == 32) issue(S.coord[(it + 1u) & 1u]);

        // ---- the oldest pending tile, if the look-back at the top of this iteration found its first id ----
        if (probe_ok) retire(q_head, probe_base);
    }
    while (q_count) {  // the tiles still pending
        const unsigned long long tb = wait_base((uint32_t)S.q[q_head].tile);
        retire(q_head, tb);
    }
}

// ---------------------------------------------------------------------------------------------
// Pass B: k_faces.  Replaces gen_faces_kernel (:140-209).
//
// A warp takes 16 consecutive (row, piece) pairs in voxel-major order; a lane owns 64 cells (half a
// piece: two bit words).  It recomputes the eight crossing masks of its words,
//   q0 (x,y) x-edges   q1 (x,y) y-edges   q2 (x,y) z-edges   q3 (x+1,y) y-edges
//   q4 (x+1,y) z-edges q5 (x,y+1) x-edges q6 (x,y+1) z-edges q7 (x+1,y+1) z-edges
// carries the id of the first crossing of each mask along the piece (table entry + popcounts), and
// parks {corner words, masks, first ids} of every word with active cells in shared memory.  The sparse
// work then runs lane-balanced: one active cell per lane (case, triangle count), one triangle per lane
// (three ranks, 12-byte store).  Cube edge e -> nibble n = q | up << 3 of a word's rank table, following the
// owner map of :178-192:
//   e: 0 1 2 3 4  5  6 7 8 9 10 11
//   n: 0 3 5 1 8 11 13 9 2 4  7  6     e4..e7 are the x-/y-edges at sample z+1 (up = 1) of masks q0, q3, q5, q1
// and one formula serves all twelve:  id = first id of mask q + popc(mask q & bits below z + up).
// All global loads of a group are issued in one batch; the triangle counts of the NEXT group (the early-out
// test) are fetched one iteration ahead.
//
// Work distribution and face offsets: a warp takes chunks of kFaceChunk consecutive groups (128 pieces) by
// ticket.  The index of a chunk's first face is the sum of the rounds before its round and of the chunks before
// it in its round (sums accumulated by k_tile: final data, one batch of loads, nobody waits for anybody); the
// groups' first indices follow from an exclusive scan of the chunk's own counts (4 pieces per lane).  There is
// no scan kernel and no per-piece offset array.
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kEdgeToEntry = (0ull << 0) | (3ull << 4) | (5ull << 8) | (1ull << 12) | (8ull << 16) | (11ull << 20) |
                                  (13ull << 24) | (9ull << 28) | (2ull << 32) | (4ull << 36) | (7ull << 40) |
                                  (6ull << 44);
constexpr int kFaceWarps = 4;
constexpr int kFaceCtasPerSm = 6;   // 79 registers x 128 threads and 34 KB of shared memory per CTA
static_assert(kFacePieces * kFaceChunk == 128, "a chunk is 4 pieces per lane");
constexpr int kFaceSlots = 64;     // bit words per warp iteration
constexpr int kCellCap = 256;
constexpr int kTriBatch = 160;     // triangles of one batch of 32 cells (<= 5 each)
constexpr int kRankStride = 9;     // uint2 per slot: 8 entries + 1 pad (bank spread)

struct FaceScratch {
    uint4 corner[kFaceSlots][2];          // per word slot: {a, b, c, d} and the words that follow them in z
    uint2 rank[kFaceSlots * kRankStride];  // per word slot and mask q: {mask, id of its first crossing (+ vertex_id_base)}
    uint16_t cell[kCellCap];              // slot<<5 | bit
    uint32_t tri[kTriBatch];              // slot<<5 | bit | three entry nibbles << 12
    unsigned long long gbase[32];         // index of the first face of each group of the chunks of the runs in flight
};
constexpr int kFaceSmemBytes = 256 * (int)sizeof(uint64_t) + kFaceWarps * (int)sizeof(FaceScratch);

// the cell's 8 corner bits in the order a0 a1 b0 b1 c0 c1 d0 d1 (x0 = sample z, x1 = sample z+1)
__device__ __forceinline__ uint32_t corner_code(const uint4 &w, const uint4 &n, int i) {
    return (__funnelshift_r(w.x, n.x, i) & 3u) | ((__funnelshift_r(w.y, n.y, i) & 3u) << 2) |
           ((__funnelshift_r(w.z, n.z, i) & 3u) << 4) | ((__funnelshift_r(w.w, n.w, i) & 3u) << 6);
}

__global__ void __launch_bounds__(kFaceWarps * 32, kFaceCtasPerSm)
    k_faces(McGeom g, McWorkspace ws, int32_t vbase, int32_t *__restrict__ faces, unsigned long long face_capacity,
            int vertex_base_from_header, int gpt, uint32_t ntickets) {
    if (vertex_base_from_header) vbase += (int32_t)ws.header->vertex_base;  // multi-GPU: computed by k_apply_exchange
    // speculative launch (p3d_mc_extract): the buffer was sized before F was known; if it is too small nothing is
    // written and the caller runs the pass again with an exact buffer
    if (ws.header->total_f > face_capacity) return;
    extern __shared__ __align__(16) unsigned char face_smem[];
    // per corner code: up to 15 entry nibbles, nibble 15 = #triangles
    uint64_t *s_table = reinterpret_cast<uint64_t *>(face_smem);
    FaceScratch *s_scratch = reinterpret_cast<FaceScratch *>(face_smem + 256 * sizeof(uint64_t));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        // c is a corner code; the case index has corner k in bit k (:168-176)
        const uint32_t cs = (c & 1u) | ((c >> 1 & 1u) << 4) | ((c >> 2 & 1u) << 1) | ((c >> 3 & 1u) << 5) | ((c >> 4 & 1u) << 2) |
                            ((c >> 5 & 1u) << 6) | ((c >> 6 & 1u) << 3) | ((c >> 7 & 1u) << 7);
        const uint64_t t = c_case_table[cs];
        const uint32_t n = (uint32_t)(t >> 60);
        uint64_t out = (uint64_t)n << 60;
        for (uint32_t j = 0; j < 3 * n; ++j) {
            const uint32_t e = (uint32_t)(t >> (4 * j)) & 15u;
            out |= ((kEdgeToEntry >> (4 * e)) & 15ull) << (4 * j);
        }
        s_table[c] = out;
    }
    __syncthreads();
    FaceScratch &sc = s_scratch[warp];

    const int np = g.np, ry = (int)g.ry, rz = (int)g.rz;
    const int64_t bstride = 4 * (int64_t)np, plane_pieces = (int64_t)ry * np;
    const int64_t ngroups = (g.npieces + kFacePieces - 1) / kFacePieces;
    const int h = lane & 1;  // which half of the piece
    const uint32_t lanes_below = (1u << lane) - 1u;

    // ---- work distribution: runs of `gpt` consecutive groups by ticket (the cost of a group follows the surface,
    // so a static round-robin leaves SMs idle at the end); the next ticket is taken a run ahead of its use.  gpt is
    // a whole chunk (8 groups) on large grids and 4, 2 or 1 on small ones, where spreading the groups over more warps
    // matters more than sharing the chunk's prefix loads.  The first face index of each group of the run's chunk is
    // parked in shared memory; up to three runs are in the pipeline at once, so four sets of slots rotate. ----
    uint32_t ticket_ahead = 0;
    auto take_ticket = [&]() {
        if (lane == 0) ticket_ahead = atomicAdd(&ws.header->ticket_faces, 1u);
    };
    take_ticket();
    uint32_t run_ticket = 0, chunk_j = 0, chunk_n = 0, chunk_slot = 0;
    bool exhausted = false;
    auto next_group = [&](uint32_t &slot) -> int64_t {
        if (chunk_j == chunk_n && !exhausted) {
            run_ticket = __shfl_sync(kFull, ticket_ahead, 0);
            if (run_ticket >= ntickets) {
                exhausted = true;
            } else {
                take_ticket();
                const int64_t run_first = (int64_t)run_ticket * gpt;
                const uint32_t chunk_id = (uint32_t)(run_first / kFaceChunk);
                chunk_j = 0;
                chunk_n = (uint32_t)(run_first + gpt < ngroups ? gpt : ngroups - run_first);
                chunk_slot = ((chunk_slot & ~7u) + 8u) & 31u;            // next set of eight slots ...
                chunk_slot += (uint32_t)(run_first % kFaceChunk);        // ... at the run's first group in its chunk
                // everything in one batch of loads: my four counts, the rounds before, the chunks before in the round
                const int64_t p0 = (int64_t)chunk_id * (kFaceChunk * kFacePieces) + 4 * lane;
                uint4 n4 = make_uint4(0u, 0u, 0u, 0u);
                if (p0 + 4 <= g.npieces) {
                    n4 = __ldg(reinterpret_cast<const uint4 *>(ws.nf + p0));
                } else {
                    if (p0 < g.npieces) n4.x = __ldg(ws.nf + p0);
                    if (p0 + 1 < g.npieces) n4.y = __ldg(ws.nf + p0 + 1);
                    if (p0 + 2 < g.npieces) n4.z = __ldg(ws.nf + p0 + 2);
                }
                const uint32_t round = chunk_id / kRoundTiles, in_round = chunk_id % kRoundTiles;
                unsigned long long acc = 0;
                const uint32_t *cs = ws.chunk_sum + (chunk_id - in_round);
#pragma unroll
                for (int i = 0; i < kRoundTiles / 32; ++i)
                    if ((uint32_t)(lane + 32 * i) < in_round) acc += __ldg(cs + lane + 32 * i);
#pragma unroll 4
                for (uint32_t r = lane; r < round; r += 32) acc += __ldg(ws.fround_sum + r);
                const unsigned long long chunk_face = warp_sum64(acc);
                const uint32_t mine = __dp4a(n4.x, 0x01010101u, 0u) + __dp4a(n4.y, 0x01010101u, 0u) +
                                      __dp4a(n4.z, 0x01010101u, 0u) + __dp4a(n4.w, 0x01010101u, 0u);
                const uint32_t excl = warp_incl_scan(mine, lane) - mine;
                // group j of the chunk = pieces 16 j .. 16 j + 15 = lanes 4 j .. 4 j + 3
                if ((lane & 3) == 0) sc.gbase[(chunk_slot & ~7u) + (lane >> 2)] = chunk_face + excl;
                __syncwarp();
            }
        }
        if (exhausted) return -1;
        slot = chunk_slot + chunk_j;
        return (int64_t)run_ticket * gpt + chunk_j++;
    };
    auto group_counts = [&](int64_t gr) {  // triangle count of my piece of group gr (the tile pass stores a byte per word)
        const int64_t i = gr * kFacePieces + (lane >> 1);
        return (gr >= 0 && i < g.npieces) ? __dp4a(__ldg(ws.nf + i), 0x01010101u, 0u) : 0u;
    };

    // ---- everything a group reads from global memory, issued as one batch (a group ahead of its use) ----
    uint4 ta, tb, td, tcc;            // table entries of my piece in rows a (x,y), b (x+1,y), d (x,y+1), c (x+1,y+1)
    uint32_t A[3], B[3], C[3], D[3];  // my two bit words of the four rows and the word after them
    uint32_t nax, nay, nby, ndx;      // lane 31: entries of the piece after mine (for the cell at bit 127)
    int p_ld;                         // my piece index within its row
    auto issue_loads = [&](int64_t gr, uint32_t nf) {
        ta = make_uint4(0, 0, 0, 0), tb = ta, td = ta, tcc = ta;
#pragma unroll
        for (int i = 0; i < 3; ++i) A[i] = B[i] = C[i] = D[i] = 0u;
        nax = nay = nby = ndx = 0u;
        p_ld = 0;
        if (!__any_sync(kFull, nf != 0u)) return;
        const int64_t gi = gr * kFacePieces + (lane >> 1);
        int64_t row;
        int p;
        if (g.npieces <= 0x7fffffffll) {
            const uint32_t rw = np == 1 ? (uint32_t)gi : (uint32_t)__umul64hi((unsigned long long)gi, g.magic_np);
            p = (int)((uint32_t)gi - rw * (uint32_t)np);
            row = rw;
        } else {
            row = gi / np;
            p = (int)(gi - row * np);
        }
        p_ld = p;
        // The piece after an active piece in its row is loaded too: the cell at bit 127 has its z+1 edges there.
        const uint32_t nf_prev = __shfl_up_sync(kFull, nf, 2);
        if (nf != 0u || (lane >= 2 && p >= 1 && nf_prev != 0u)) {
            const uint4 *e = ws.ptab + gi;
            ta = __ldg(e), tb = __ldg(e + plane_pieces), td = __ldg(e + np), tcc = __ldg(e + plane_pieces + np);
            if (lane == 31 && nf && p + 1 < np) {
                const uint4 t0 = __ldg(e + 1);
                nax = t0.x, nay = t0.y;
                nby = __ldg(e + 1 + plane_pieces).y;
                ndx = __ldg(e + 1 + np).x;
            }
        }
        if (nf) {
            const uint32_t *pa = ws.bits + row * bstride + 4 * p + 2 * h;
            const uint32_t *pb = pa + ry * bstride, *pd = pa + bstride, *pc = pb + bstride;
            const uint2 a2 = __ldg(reinterpret_cast<const uint2 *>(pa)), b2 = __ldg(reinterpret_cast<const uint2 *>(pb));
            const uint2 c2 = __ldg(reinterpret_cast<const uint2 *>(pc)), d2 = __ldg(reinterpret_cast<const uint2 *>(pd));
            const bool more = h == 0 || p + 1 < np;  // the word after mine exists in the row
            A[0] = a2.x, A[1] = a2.y, A[2] = more ? __ldg(pa + 2) : 0u;
            B[0] = b2.x, B[1] = b2.y, B[2] = more ? __ldg(pb + 2) : 0u;
            C[0] = c2.x, C[1] = c2.y, C[2] = more ? __ldg(pc + 2) : 0u;
            D[0] = d2.x, D[1] = d2.y, D[2] = more ? __ldg(pd + 2) : 0u;
        }
    };

    uint32_t fs0 = 0, fs1 = 0, fs2 = 0;  // gbase slots of groups g0, g1, g2
    int64_t g0 = next_group(fs0), g1 = next_group(fs1);
    uint32_t nf0 = group_counts(g0), nf1 = group_counts(g1);  // != 0 only for rows with x + 1 < rx and y + 1 < ry (k_tile)
    issue_loads(g0, nf0);

    while (g0 >= 0) {
        const int64_t g2 = next_group(fs2);
        const uint32_t nf2 = group_counts(g2);
        const uint32_t nf = nf0;
        const bool active = __any_sync(kFull, nf != 0u);

        // ---- dense phase: masks, first ids, scratch of the words with active cells ----
        uint32_t actw[2] = {0u, 0u};
        uint32_t nact = 0, cincl = 0, ncell = 0, finc = 0;
        uint32_t xax = 0, xay = 0, xby = 0, xdx = 0;
        unsigned long long fbase = 0;
        if (active) {
            // entries of the next piece of the same rows (lane 31 loaded its own)
            {
                const uint32_t s0 = __shfl_down_sync(kFull, ta.x, 2), s1 = __shfl_down_sync(kFull, ta.y, 2);
                const uint32_t s2 = __shfl_down_sync(kFull, tb.y, 2), s3 = __shfl_down_sync(kFull, td.x, 2);
                xax = lane < 30 ? s0 : nax, xay = lane < 30 ? s1 : nay, xby = lane < 30 ? s2 : nby, xdx = lane < 30 ? s3 : ndx;
            }
            fbase = sc.gbase[fs0];
            finc = warp_incl_scan(h ? 0u : nf, lane);  // faces of the pieces up to and including mine

            // the eight masks of my word w (samples outside the grid were staged as 0.0f in every row, so the x/y
            // masks are zero there, and a z crossing cut by zv can only sit above every valid cell of the row)
            auto masks = [&](int w, uint32_t (&m)[8], uint32_t &act) {
                const uint32_t A2 = __funnelshift_r(A[w], A[w + 1], 1), B2 = __funnelshift_r(B[w], B[w + 1], 1);
                const uint32_t C2 = __funnelshift_r(C[w], C[w + 1], 1), D2 = __funnelshift_r(D[w], D[w + 1], 1);
                const uint32_t zv = low_mask(rz - 1 - (p_ld * kTileZ + 32 * (2 * h + w)));
                m[0] = A[w] ^ B[w], m[1] = A[w] ^ D[w], m[2] = (A[w] ^ A2) & zv, m[3] = B[w] ^ C[w];
                m[4] = (B[w] ^ B2) & zv, m[5] = D[w] ^ C[w], m[6] = (D[w] ^ D2) & zv, m[7] = (C[w] ^ C2) & zv;
                // mixed corners (:154,168-176) <=> one of the bottom x/y edges or of the four z edges is crossed
                act = nf ? (((m[0] | m[1] | m[2]) | (m[3] | m[4] | m[5]) | (m[6] | m[7])) & zv) : 0u;
            };
            uint32_t run[8] = {ta.x, ta.y, ta.z, tb.y, tb.z, td.x, td.z, tcc.z};
            // crossings of the first half's two words per mask (8-bit fields): the second half starts after them
            uint32_t half_lo = 0, half_hi = 0;
            uint32_t m0[8], m1[8];
            masks(0, m0, actw[0]);
            masks(1, m1, actw[1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                half_lo += (uint32_t)(__popc(m0[q]) + __popc(m1[q])) << (8 * q);
                half_hi += (uint32_t)(__popc(m0[q + 4]) + __popc(m1[q + 4])) << (8 * q);
            }
            const uint32_t prev_lo = __shfl_up_sync(kFull, half_lo, 1), prev_hi = __shfl_up_sync(kFull, half_hi, 1);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t prev = ((q < 4 ? prev_lo : prev_hi) >> (8 * (q & 3))) & 255u;
                run[q] += (uint32_t)vbase + (h ? prev : 0u);
            }
            if (actw[0]) {
                const int slot = lane * 2;
                sc.corner[slot][0] = make_uint4(A[0], B[0], C[0], D[0]);
                sc.corner[slot][1] = make_uint4(A[1], B[1], C[1], D[1]);
                uint2 *rk = &sc.rank[slot * kRankStride];
#pragma unroll
                for (int q = 0; q < 8; ++q) rk[q] = make_uint2(m0[q], run[q]);
            }
            if (actw[1]) {
                const int slot = lane * 2 + 1;
                sc.corner[slot][0] = make_uint4(A[1], B[1], C[1], D[1]);
                sc.corner[slot][1] = make_uint4(A[2], B[2], C[2], D[2]);
                uint2 *rk = &sc.rank[slot * kRankStride];
#pragma unroll
                for (int q = 0; q < 8; ++q) rk[q] = make_uint2(m1[q], run[q] + __popc(m0[q]));
            }
            nact = __popc(actw[0]) + __popc(actw[1]);
            cincl = warp_incl_scan(nact, lane);
            ncell = __shfl_sync(kFull, cincl, 31);
        }

        // ---- the next group's loads fly while this group's sparse phase runs from shared memory ----
        issue_loads(g1, nf1);
        __syncwarp();

        if (active) {
            unsigned long long frun = fbase;
            for (uint32_t c0 = 0; c0 < ncell; c0 += kCellCap) {
                {
                    uint32_t pos = cincl - nact - c0;  // wraps below the chunk: filtered by the range test
#pragma unroll
                    for (int w = 0; w < 2; ++w)
                        for (uint32_t rem = actw[w]; rem; ++pos) {
                            const int i = __ffs(rem) - 1;
                            rem &= rem - 1;
                            if (pos < (uint32_t)kCellCap) sc.cell[pos] = (uint16_t)(((lane * 2 + w) << 5) | i);
                        }
                }
                __syncwarp();
                const uint32_t n = ncell - c0 < (uint32_t)kCellCap ? ncell - c0 : (uint32_t)kCellCap;
                for (uint32_t k0 = 0; k0 < n; k0 += 32) {
                    // one cell per lane: case, triangle count, and one list entry per triangle
                    const uint32_t k = k0 + lane;
                    uint32_t nt = 0, e = 0;
                    uint64_t tt = 0;
                    if (k < n) {
                        e = sc.cell[k];
                        tt = s_table[corner_code(sc.corner[e >> 5][0], sc.corner[e >> 5][1], e & 31u)];
                        nt = (uint32_t)(tt >> 60);
                    }
                    // exclusive prefix of nt (<= 5: three bit planes) by ballots: no shuffle chain
                    const uint32_t b0 = __ballot_sync(kFull, nt & 1u), b1 = __ballot_sync(kFull, nt & 2u), b2 = __ballot_sync(kFull, nt & 4u);
                    const uint32_t tp = __popc(b0 & lanes_below) + 2u * __popc(b1 & lanes_below) + 4u * __popc(b2 & lanes_below);
                    const uint32_t btot = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
                    {
                        const uint32_t lo = (uint32_t)tt, hi = (uint32_t)(tt >> 32);
                        uint32_t *dst = &sc.tri[tp];
                        if (nt > 0) dst[0] = e | ((lo & 0xfffu) << 12);
                        if (nt > 1) dst[1] = e | (((lo >> 12) & 0xfffu) << 12);
                        if (nt > 2) dst[2] = e | ((__funnelshift_r(lo, hi, 24) & 0xfffu) << 12);
                        if (nt > 3) dst[3] = e | (((hi >> 4) & 0xfffu) << 12);
                        if (nt > 4) dst[4] = e | (((hi >> 16) & 0xfffu) << 12);
                    }
                    __syncwarp();
                    // one triangle per lane: rank its three edges, 12-byte stores (:194-208)
                    int32_t *const out0 = faces + frun * 3ull;
                    for (uint32_t j = lane; j < btot; j += 32) {
                        const uint32_t ent = sc.tri[j];
                        const uint32_t i = ent & 31u;
                        const uint32_t below0 = (1u << i) - 1u, below1 = (2u << i) - 1u;  // bits below z / below z+1
                        const uint2 *rk = &sc.rank[((ent >> 5) & 63u) * kRankStride];
                        int32_t *out = out0 + j * 3u;
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            const uint32_t nib = ent >> (12 + 4 * cc);
                            const uint2 en = rk[nib & 7u];
                            out[cc] = (int32_t)(en.y + __popc(en.x & ((nib & 8u) ? below1 : below0)));
                        }
                    }
                    __syncwarp();
                    frun += btot;
                }
            }

            // the last cell of a piece (bit 127) has its z+1 x-/y-edges in the NEXT piece, which is numbered by
            // another tile: overwrite those indices with that piece's table entries (its bit 0 is rank 0)
            if (h && (actw[1] >> 31)) {
                const int slot = lane * 2 + 1;
                uint64_t tt = s_table[corner_code(sc.corner[slot][0], sc.corner[slot][1], 31)];
                const uint32_t nt = (uint32_t)(tt >> 60);
                int32_t *out = faces + (fbase + finc - nt) * 3ull;
                for (uint32_t t = 0; t < nt; ++t)
                    for (int cc = 0; cc < 3; ++cc, tt >>= 4) {
                        const uint32_t n = (uint32_t)tt & 15u;
                        if (n >= 8u) out[t * 3 + cc] = vbase + (int32_t)(n == 8u ? xax : (n == 9u ? xay : (n == 11u ? xby : xdx)));
                    }
            }
            __syncwarp();
        }
        g0 = g1, nf0 = nf1, fs0 = fs1;
        g1 = g2, nf1 = nf2, fs1 = fs2;
    }
}

// ---------------------------------------------------------------------------------------------
// Host side.
// ---------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

thread_local const char *g_tile_error = nullptr;

// P3D_MC_LOADER = generic forces the non-TMA staging path (A/B runs, tests)
bool force_generic() {
    static const bool v = [] {
        const char *e = getenv("P3D_MC_LOADER");
        return e && e[0] == 'g';
    }();
    return v;
}

template <bool TMA, typename T, bool SPARSE = false>
void launch_tile_kernel(const CUtensorMap &map, const void *grid, const McGeom &g, const McWorkspace &ws,
                        const McEmitParams &p, float *verts, int64_t vcap, int mode, cudaStream_t s) {
    // every CTA must be resident: a CTA waits for tiles with lower ids, which running CTAs hold
    static int cache[kMaxDevices];
    const int per_sm = per_device(cache, [] {
        cudaFuncSetAttribute(k_tile<TMA, T, SPARSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes);
        int n = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_tile<TMA, T, SPARSE>, kTileThreads, kTileSmemBytes);
        return n > 0 ? n : 1;
    });
    const int64_t cap = (int64_t)sm_count() * per_sm;
    const unsigned blocks = (unsigned)(g.ntiles < cap ? g.ntiles : cap);
    k_tile<TMA, T, SPARSE><<<blocks, kTileThreads, kTileSmemBytes, s>>>(map, static_cast<const T *>(grid), g, ws, p, verts,
                                                                        (unsigned long long)(vcap > 0 ? vcap : 0), mode);
}

}  // namespace

const char *tile_pass_error() { return g_tile_error; }

void launch_tile_pass(const void *grid, int dtype, const McGeom &g, const McWorkspace &ws, const McEmitParams &p,
                      float *verts, int64_t vertex_capacity, int mode, cudaStream_t s) {
    g_tile_error = nullptr;
    if (g.ntiles <= 0) return;
    if (!verts) vertex_capacity = 0;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    // TMA needs a 16-byte aligned base and 16-byte multiples as row / plane strides
    bool tma = dtype == 0 && !force_generic() && (g.rz % 4 == 0) && ((reinterpret_cast<uintptr_t>(grid) & 15) == 0);
    if (tma) {
        EncodeTiledFn enc = encode_tiled();
        if (!enc) {
            tma = false;
        } else {
            const cuuint64_t dims[3] = {(cuuint64_t)g.rz, (cuuint64_t)g.ry, (cuuint64_t)g.rx};
            const cuuint64_t strides[2] = {(cuuint64_t)g.rz * 4, (cuuint64_t)g.ry * (cuuint64_t)g.rz * 4};
            const cuuint32_t box[3] = {kBoxZ, kTileY + 1, kTileX + 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(grid), dims, strides, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                g_tile_error = "cuTensorMapEncodeTiled failed";
                return;
            }
        }
    }
    if (g.tile_list) {  // block-sparse form: float32 grids
        if (dtype != 0) g_tile_error = "the block-sparse form takes float32 grids";
        else if (tma) launch_tile_kernel<true, float, true>(map, grid, g, ws, p, verts, vertex_capacity, mode, s);
        else launch_tile_kernel<false, float, true>(map, grid, g, ws, p, verts, vertex_capacity, mode, s);
        return;
    }
    if (tma) {
        launch_tile_kernel<true, float>(map, grid, g, ws, p, verts, vertex_capacity, mode, s);
        return;
    }
    switch (dtype) {  // p3d_dtype, include/prim3d_b200.h
        case 0: launch_tile_kernel<false, float>(map, grid, g, ws, p, verts, vertex_capacity, mode, s); break;
        case 1: launch_tile_kernel<false, __half>(map, grid, g, ws, p, verts, vertex_capacity, mode, s); break;
        case 2: launch_tile_kernel<false, __nv_bfloat16>(map, grid, g, ws, p, verts, vertex_capacity, mode, s); break;
        case 3: launch_tile_kernel<false, double>(map, grid, g, ws, p, verts, vertex_capacity, mode, s); break;
        case 4: launch_tile_kernel<false, long long>(map, grid, g, ws, p, verts, vertex_capacity, mode, s); break;
        case 5: launch_tile_kernel<false, int>(map, grid, g, ws, p, verts, vertex_capacity, mode, s); break;
        case 6: launch_tile_kernel<false, short>(map, grid, g, ws, p, verts, vertex_capacity, mode, s); break;
        case 7: launch_tile_kernel<false, unsigned char>(map, grid, g, ws, p, verts, vertex_capacity, mode, s); break;
        default: g_tile_error = "unknown grid dtype"; break;
    }
}

// triangles of each round of 256 chunks (a CTA per round)
__global__ void __launch_bounds__(kRoundTiles) k_round_sums(McGeom g, McWorkspace ws) {
    __shared__ unsigned long long s_warp[kRoundTiles / 32];
    const int64_t c = (int64_t)blockIdx.x * kRoundTiles + threadIdx.x;
    const unsigned long long v = warp_sum64(c < g.nchunks ? ws.chunk_sum[c] : 0u);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
#pragma unroll
        for (int i = 0; i < kRoundTiles / 32; ++i) t += s_warp[i];
        ws.fround_sum[blockIdx.x] = t;
        if (blockIdx.x == 0) ws.header->ticket_faces = 0u;  // the face pass's ticket counter (saves a memset node)
    }
}

void launch_faces(const McGeom &g, const McWorkspace &ws, const McEmitParams &p, int32_t *faces, int64_t face_capacity,
                  bool vertex_base_from_header, cudaStream_t s) {
    if (g.npieces <= 0) return;
    k_round_sums<<<(unsigned)g.nfrounds, kRoundTiles, 0, s>>>(g, ws);
    if (faces_rows_applicable(g)) {
        launch_faces_rows(g, ws, p, faces, face_capacity, vertex_base_from_header, s);
        return;
    }
    const int64_t groups = (g.npieces + kFacePieces - 1) / kFacePieces;
    const int64_t cap = (int64_t)sm_count() * kFaceCtasPerSm;
    // groups per ticket: a whole chunk when there are at least two runs per resident warp, else fewer (small grids)
    int gpt = kFaceChunk;
    while (gpt > 1 && groups / gpt < 2 * cap * kFaceWarps) gpt /= 2;
    const int64_t want = ((groups + gpt - 1) / gpt + kFaceWarps - 1) / kFaceWarps;
    static int cache[kMaxDevices];
    per_device(cache, [] {
        cudaFuncSetAttribute(k_faces, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaceSmemBytes);
        return 1;
    });
    k_faces<<<(unsigned)(want < cap ? want : cap), kFaceWarps * 32, kFaceSmemBytes, s>>>(g, ws, p.vertex_id_base, faces,
                                                                                         (unsigned long long)face_capacity,
                                                                                         vertex_base_from_header ? 1 : 0, gpt,
                                                                                         (uint32_t)((groups + gpt - 1) / gpt));
}

// Multi-GPU.  Export: the first plane's table entries (shard-local ids).  Import: install the next shard's
// first-plane entries, shifted by this shard's vertex count, as this shard's halo-plane numbering.
__global__ void k_shift_plane(uint4 *__restrict__ dst, const uint4 *__restrict__ src, int64_t n, uint32_t delta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint4 t = src[i];
        dst[i] = make_uint4(t.x + delta, t.y + delta, t.z + delta, 0u);
    }
}

// Device-side exchange (no host round trip between the two passes).  Export: first-plane table entries followed by
// {V, F}.  Apply: vertex_base = sum of the lower shards' V; the next shard's first-plane entries, shifted by this
// shard's V, become this shard's halo-plane numbering.
__global__ void k_export_exchange(uint4 *__restrict__ dst, const uint4 *__restrict__ src, int64_t n, const McHeader *header) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint4 t = src[i];
        dst[i] = make_uint4(t.x, t.y, t.z, 0u);
    } else if (i == n) {
        const unsigned long long v = header->total_v, f = header->total_f;
        dst[n] = make_uint4((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)f, (uint32_t)(f >> 32));
    }
}

__global__ void k_apply_exchange(McWorkspace ws, uint4 *__restrict__ halo, const uint4 *__restrict__ gathered, int64_t n,
                                 int rank, int world) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        unsigned long long base = 0;
        for (int r = 0; r < rank; ++r) {
            const uint4 c = gathered[(int64_t)r * (n + 1) + n];
            base += (unsigned long long)c.x | ((unsigned long long)c.y << 32);
        }
        ws.header->vertex_base = base;
    }
    if (i < n && rank + 1 < world) {
        const uint32_t delta = (uint32_t)ws.header->total_v;
        const uint4 t = gathered[(int64_t)(rank + 1) * (n + 1) + i];
        halo[i] = make_uint4(t.x + delta, t.y + delta, t.z + delta, 0u);
    }
}

void launch_export_exchange(uint32_t *out, const McGeom &g, const McWorkspace &ws, cudaStream_t s) {
    const int64_t n = g.ry * (int64_t)g.np;
    k_export_exchange<<<(unsigned)((n + 1 + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint4 *>(out), ws.ptab, n, ws.header);
}

void launch_apply_exchange(const McGeom &g, const McWorkspace &ws, const uint32_t *gathered, int rank, int world,
                           cudaStream_t s) {
    const int64_t n = g.ry * (int64_t)g.np;
    k_apply_exchange<<<(unsigned)((n + 255) / 256 > 0 ? (n + 255) / 256 : 1), 256, 0, s>>>(
        ws, ws.ptab + g.owned_x * n, reinterpret_cast<const uint4 *>(gathered), n, rank, world);
}

void launch_export_plane(uint32_t *table_out, const McGeom &g, const McWorkspace &ws, cudaStream_t s) {
    const int64_t n = g.ry * (int64_t)g.np;
    if (n <= 0) return;
    k_shift_plane<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint4 *>(table_out), ws.ptab, n, 0u);
}

void launch_import_halo(const McGeom &g, const McWorkspace &ws, const uint32_t *table_in, uint32_t delta, cudaStream_t s) {
    const int64_t n = g.ry * (int64_t)g.np;
    if (n <= 0) return;
    k_shift_plane<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ws.ptab + g.owned_x * n, reinterpret_cast<const uint4 *>(table_in), n, delta);
}

}  // namespace p3d
