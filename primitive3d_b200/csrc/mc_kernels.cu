// primitive3d_b200/csrc/mc_kernels.cu -- see mc_kernels.cuh for the design.
// Reference citations are into /root/reference/src/prim3d/Utility/marching_cubes.cu.
#include "mc_kernels.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime)

#include <cstdlib>
#include <cstring>

#include "mc_case_table.h"
#include "scan_utils.cuh"

namespace p3d {

// Bourke case table, one packed word per case (nibble i = i-th edge index, nibble 15 =
// #triangles).  Lives in constant memory and is staged into shared memory once per
// persistent CTA because the per-cell lookups are lane-divergent.
__constant__ uint64_t c_case_table[256] = P3D_MC_CASE_TABLE_INIT;

// Bits of word w (32 samples from z = 32*w) with z + 1 < rz: samples that own a +z edge / a cell.
__device__ __forceinline__ uint32_t low_mask(int64_t n) {
    return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << (int)n) - 1u));
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
// Pass A: k_tile.
//
// Shared memory of a CTA:
//   stage[2]   fp32 samples of a tile: rows (xi, yi) of 0..8 x 0..8, kBoxZ samples each
//   sbits      [81][8] words: inside bits of every staged row (word 4, bit 0 = the halo sample)
//   s_piece    [64] vertex count of each owned (row, piece)
//   s_list     compacted crossing edges of the tile: axis<<13 | row<<7 | z
// A thread owns bit word (row r = tid>>2, word w = tid&3) of the tile in the count phase.
// ---------------------------------------------------------------------------------------------
constexpr int kListCap = 2048;
constexpr int kSbitsStride = 8;
constexpr int kRowPitch = kTileY + 1;  // staged rows per plane

struct TileSmem {
    uint32_t sbits[kBoxRows * kSbitsStride];
    uint32_t piece[kTileX * kTileY];
    uint16_t list[kListCap];
    uint8_t ntri[256];  // indexed by the corner bits in staging order a0 a1 b0 b1 c0 c1 d0 d1
    unsigned long long bar[2];
    unsigned long long tile_base;
    int4 coord[2];      // {x0, y0, piece, -}
    uint32_t tile[2];
};
constexpr int kTileSmemBytes = 2 * kStageBytes + (int)sizeof(TileSmem) + 128;

template <bool TMA>
__global__ void __launch_bounds__(kTileThreads, 2)
    k_tile(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ grid, McGeom g, McWorkspace ws,
           McEmitParams prm, float *__restrict__ verts, unsigned long long vcap, int mode) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    float *stage0 = reinterpret_cast<float *>(sm);
    TileSmem &S = *reinterpret_cast<TileSmem *>(sm + 2 * kStageBytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ntiles = (uint32_t)g.ntiles;

    // ntri by staged corner order: bit0 = corner 0 (a, z), bit1 = corner 4 (a, z+1), bit2 = corner 1 (b, z),
    // bit3 = corner 5, bit4 = corner 2 (c, z), bit5 = corner 6, bit6 = corner 3 (d, z), bit7 = corner 7
    // (corner numbering of marching_cubes.cu:50-57)
    {
        const uint32_t c = tid;
        const uint32_t cs = (c & 1u) | ((c >> 1 & 1u) << 4) | ((c >> 2 & 1u) << 1) | ((c >> 3 & 1u) << 5) | ((c >> 4 & 1u) << 2) |
                            ((c >> 5 & 1u) << 6) | ((c >> 6 & 1u) << 3) | ((c >> 7 & 1u) << 7);
        S.ntri[c] = (uint8_t)(c_case_table[cs] >> 60);
    }

    // tile id -> coordinates; tiles are ordered band by band (a band = `band` y-blocks over all x), inside a
    // band x-block major, then y-block, then piece: the x halo plane of a block is re-read from L2, not HBM
    auto fetch = [&](int st) {
        const uint32_t t = atomicAdd(&ws.header->ticket, 1u);
        S.tile[st] = t;
        if (t >= ntiles) return;
        const int64_t per_band = (int64_t)g.nxb * g.band * g.np;
        const int64_t bi = (int64_t)t / per_band, rem = (int64_t)t - bi * per_band;
        const int64_t left = g.nyb - bi * g.band, cur = left < g.band ? left : g.band;
        const int64_t xb = rem / (cur * g.np), rem2 = rem - xb * (cur * g.np);
        const int64_t yi = rem2 / g.np, p = rem2 - yi * g.np;
        const int x0 = (int)(xb * kTileX), y0 = (int)((bi * g.band + yi) * kTileY);
        S.coord[st] = make_int4(x0, y0, (int)p, 0);
        if (TMA) {
            const uint32_t bar = smem_u32(&S.bar[st]);
            mbar_expect_tx(bar, kBoxRows * kBoxZ * 4);
            tma_load_3d(smem_u32(sm + st * kStageBytes), &tmap, bar, (int)p * kTileZ, y0, x0);
        }
    };

    if (tid == 0) {
        if (TMA) {
            mbar_init(smem_u32(&S.bar[0]), 1);
            mbar_init(smem_u32(&S.bar[1]), 1);
            mbar_fence_init();
        }
        fetch(0);
        fetch(1);
    }
    __syncthreads();

    const int64_t bstride = 4 * (int64_t)g.np;  // bit words per row
    const float thresh = prm.thresh;

    for (uint32_t it = 0;; ++it) {
        const int st = it & 1;
        const uint32_t tile = S.tile[st];
        if (tile >= ntiles) break;
        const int4 tc = S.coord[st];
        const int x0 = tc.x, y0 = tc.y, p = tc.z;
        const int64_t z0 = (int64_t)p * kTileZ;
        float *tf = TMA ? reinterpret_cast<float *>(sm + st * kStageBytes) : stage0;

        if (TMA) {
            mbar_wait(smem_u32(&S.bar[st]), (it >> 1) & 1u);
        } else {
            for (int idx = tid; idx < kBoxRows * kBoxZ; idx += kTileThreads) {
                const int r = idx / kBoxZ, c = idx - r * kBoxZ;
                const int xi = r / kRowPitch, yi = r - xi * kRowPitch;
                const int64_t gx = x0 + xi, gy = y0 + yi, gz = z0 + c;
                tf[idx] = (gx < g.rx && gy < g.ry && gz < g.rz) ? __ldg(grid + (gx * g.ry + gy) * g.rz + gz) : 0.0f;
            }
            __syncthreads();
        }

        // ---- phase 1: inside bits of the 81 staged rows.  inside = value > thresh (:25,31,37,43,50-57) ----
        {
            const int64_t nz = g.rz - z0;  // samples of this piece inside the grid (>= 1)
            const bool ztail = nz < kTileZ;
            const bool halo_ok = nz > kTileZ;
            for (int r = warp; r < kBoxRows; r += kTileThreads / 32) {
                const int xi = r / kRowPitch, yi = r - xi * kRowPitch;
                const bool rowok = (x0 + xi < g.rx) && (y0 + yi < g.ry);
                const float *src = tf + r * kBoxZ;
                const float f0 = src[lane], f1 = src[lane + 32], f2 = src[lane + 64], f3 = src[lane + 96];
                uint32_t b0 = __ballot_sync(kFull, f0 > thresh), b1 = __ballot_sync(kFull, f1 > thresh);
                uint32_t b2 = __ballot_sync(kFull, f2 > thresh), b3 = __ballot_sync(kFull, f3 > thresh);
                if (ztail) {
                    b0 &= low_mask(nz);
                    b1 &= low_mask(nz - 32);
                    b2 &= low_mask(nz - 64);
                    b3 &= low_mask(nz - 96);
                }
                if (!rowok) b0 = b1 = b2 = b3 = 0u;
                if (lane == 0) {
                    *reinterpret_cast<uint4 *>(&S.sbits[r * kSbitsStride]) = make_uint4(b0, b1, b2, b3);
                    S.sbits[r * kSbitsStride + 4] = (rowok && halo_ok && src[kTileZ] > thresh) ? 1u : 0u;
                }
            }
        }
        __syncthreads();

        // ---- phase 2: crossing masks and counts of my word ----
        const int r = tid >> 2, w = tid & 3;
        const int xi = r >> 3, yi = r & 7;
        const int ra = xi * kRowPitch + yi;
        const int64_t x = x0 + xi, y = y0 + yi;
        const bool inrow = (x < g.rx) && (y < g.ry);
        const bool own = (x < g.owned_x) && (y < g.ry);
        const bool hx = own && (x + 1 < g.rx), hy = own && (y + 1 < g.ry), hc = hx && hy;
        const uint32_t *sa = &S.sbits[ra * kSbitsStride + w];
        const uint32_t A = sa[0], An = sa[1];
        const uint32_t B = sa[kRowPitch * kSbitsStride], Bn = sa[kRowPitch * kSbitsStride + 1];
        const uint32_t D = sa[kSbitsStride], Dn = sa[kSbitsStride + 1];
        const uint32_t C = sa[(kRowPitch + 1) * kSbitsStride], Cn = sa[(kRowPitch + 1) * kSbitsStride + 1];
        const uint32_t A2 = __funnelshift_r(A, An, 1), B2 = __funnelshift_r(B, Bn, 1);
        const uint32_t C2 = __funnelshift_r(C, Cn, 1), D2 = __funnelshift_r(D, Dn, 1);
        const uint32_t zv = low_mask(g.rz - 1 - (z0 + 32 * w));
        const uint32_t m0 = hx ? (A ^ B) : 0u;            // +x edges, :29-33 / :100-111
        const uint32_t m1 = hy ? (A ^ D) : 0u;            // +y edges, :35-39 / :113-124
        const uint32_t m2 = own ? ((A ^ A2) & zv) : 0u;   // +z edges, :41-45 / :126-137
        uint32_t act = 0u;                                // cells with mixed corners, :48-66
        if (hc) {
            const uint32_t any = A | B | C | D | A2 | B2 | C2 | D2, all = A & B & C & D & A2 & B2 & C2 & D2;
            act = (any & ~all) & zv;
        }
        uint32_t nf = 0;
        for (uint32_t rem = act; rem;) {
            const int i = __ffs(rem) - 1;
            rem &= rem - 1;
            const uint32_t code = (__funnelshift_r(A, An, i) & 3u) | ((__funnelshift_r(B, Bn, i) & 3u) << 2) |
                                  ((__funnelshift_r(C, Cn, i) & 3u) << 4) | ((__funnelshift_r(D, Dn, i) & 3u) << 6);
            nf += S.ntri[code];
        }
        // my (row, piece) = 4 adjacent lanes: packed {nx, ny, nz} (8-bit fields, <= 128 each)
        const uint32_t cnt = (uint32_t)__popc(m0) | ((uint32_t)__popc(m1) << 8) | ((uint32_t)__popc(m2) << 16);
        uint32_t inc = cnt;
        {
            uint32_t t = __shfl_up_sync(kFull, inc, 1);
            if (w >= 1) inc += t;
            t = __shfl_up_sync(kFull, inc, 2);
            if (w >= 2) inc += t;
        }
        const uint32_t tot = __shfl_sync(kFull, inc, lane | 3);
        const uint32_t exw = inc - cnt;
        nf += __shfl_xor_sync(kFull, nf, 1);
        nf += __shfl_xor_sync(kFull, nf, 2);
        if (w == 0) S.piece[r] = (tot & 255u) + ((tot >> 8) & 255u) + (tot >> 16);

        if (mode == 0) {
            if (inrow) ws.bits[(x * g.ry + y) * bstride + 4 * p + w] = A;
            if (own && w == 0) ws.nf[(x * g.ry + y) * g.np + p] = nf;
            // the halo plane of a slab sits one past the last x-block when owned_x is a multiple of 8
            if (tid < 32 && x0 + kTileX == g.owned_x && g.owned_x < g.rx) {
                const int hy_ = tid >> 2, hw = tid & 3;
                if (y0 + hy_ < g.ry)
                    ws.bits[(g.owned_x * g.ry + y0 + hy_) * bstride + 4 * p + hw] =
                        S.sbits[(kTileX * kRowPitch + hy_) * kSbitsStride + hw];
            }
        }
        __syncthreads();

        // ---- tile scan (every warp redundantly): first vertex of each (row, piece), relative to the tile ----
        uint32_t vt, pe;
        {
            const uint32_t v0 = S.piece[2 * lane], v1 = S.piece[2 * lane + 1];
            const uint32_t s2 = v0 + v1, incl = warp_incl_scan(s2, lane);
            vt = __shfl_sync(kFull, incl, 31);
            const uint32_t e0 = incl - s2, e1 = e0 + v0;
            const uint32_t g0 = __shfl_sync(kFull, e0, r >> 1), g1 = __shfl_sync(kFull, e1, r >> 1);
            pe = (r & 1) ? g1 : g0;
        }
        const uint32_t vx_rel = pe, vy_rel = pe + (tot & 255u), vz_rel = vy_rel + ((tot >> 8) & 255u);
        const uint32_t wfirst[3] = {vx_rel + (exw & 255u), vy_rel + ((exw >> 8) & 255u), vz_rel + (exw >> 16)};
        const uint32_t wmask[3] = {m0, m1, m2};

        // ---- first vertex id of the tile: decoupled look-back over tiles (warp 0) ----
        if (warp == 0) {
            unsigned long long tb;
            if (mode == 0) {
                tb = lookback(ws.status, (int64_t)tile, (unsigned long long)vt, lane);
                if (lane == 0 && tile == ntiles - 1) ws.header->total_v = tb + vt;
            } else {
                tb = tile ? (ws.status[tile - 1] & kValueMask) : 0ull;
            }
            if (lane == 0) S.tile_base = tb;
        }

        // ---- vertices: compact the crossing edges, then one edge per thread (gen_vertices_kernel :70-138) ----
        for (uint32_t c0 = 0; c0 < vt || c0 == 0; c0 += kListCap) {
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                uint32_t pos = wfirst[ax] - c0;  // wraps for entries below the chunk: filtered by the range test
                for (uint32_t rem = wmask[ax]; rem; ++pos) {
                    const int i = __ffs(rem) - 1;
                    rem &= rem - 1;
                    if (pos < (uint32_t)kListCap) S.list[pos] = (uint16_t)((ax << 13) | (r << 7) | (w << 5) | i);
                }
            }
            __syncthreads();  // list complete; S.tile_base visible
            const unsigned long long tb = S.tile_base;
            if (c0 == 0 && mode == 0 && own && w == 0)
                ws.ptab[(x * g.ry + y) * g.np + p] =
                    make_uint4((uint32_t)tb + vx_rel, (uint32_t)tb + vy_rel, (uint32_t)tb + vz_rel, nf);
            const uint32_t n = vt - c0 < (uint32_t)kListCap ? vt - c0 : (uint32_t)kListCap;
            for (uint32_t k = tid; k < n && vt > c0; k += kTileThreads) {
                const unsigned long long id = tb + c0 + k;
                if (id >= vcap) continue;
                const uint32_t ent = S.list[k];
                const uint32_t ax = ent >> 13, er = (ent >> 7) & 63u, ez = ent & 127u;
                const uint32_t exi = er >> 3, eyi = er & 7u;
                const float *src = tf + (exi * kRowPitch + eyi) * kBoxZ + ez;
                const float d0 = src[0];
                const float d1 = src[ax == 0 ? kRowPitch * kBoxZ : (ax == 1 ? kBoxZ : 1)];
                // dt = (thresh - d_self) / (d_next - d_self), IEEE fp32, no contraction (:105)
                const float dt = __fdiv_rn(__fsub_rn(thresh, d0), __fsub_rn(d1, d0));
                float px = (float)(prm.x_origin + x0 + (int)exi);  // static_cast<float>(x), :107
                float py = (float)(y0 + (int)eyi);
                float pz = (float)(z0 + ez);
                if (ax == 0) px = __fadd_rn(px, dt);
                if (ax == 1) py = __fadd_rn(py, dt);
                if (ax == 2) pz = __fadd_rn(pz, dt);
                // vertices * scale + offset as two separately rounded ops (:298)
                float *out = verts + id * 3ull;
                out[0] = __fadd_rn(__fmul_rn(px, prm.scale[0]), prm.offset[0]);
                out[1] = __fadd_rn(__fmul_rn(py, prm.scale[1]), prm.offset[1]);
                out[2] = __fadd_rn(__fmul_rn(pz, prm.scale[2]), prm.offset[2]);
            }
            __syncthreads();  // the list (next chunk) and the stage (next tile) may be overwritten
        }

        if (tid == 0) fetch(st);
    }
}

// ---------------------------------------------------------------------------------------------
// k_fscan: exclusive scan of the per-piece triangle counts in voxel-major (row, piece) order --
// single pass, decoupled look-back over tiles of 2048 pieces.  f8[i] = index of the first face of
// piece 8*i; the grand total goes to the header.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fscan(McGeom g, McWorkspace ws) {
    __shared__ unsigned long long s_warp[8];
    __shared__ unsigned long long s_excl;
    __shared__ unsigned int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&ws.header->ticket_scan, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= g.nscan) break;
        const int64_t r0 = tile * kFscanTile + (int64_t)threadIdx.x * 8;
        uint32_t c[8];
        if (r0 + 8 <= g.npieces) {
            const uint4 lo = *reinterpret_cast<const uint4 *>(ws.nf + r0), hi = *reinterpret_cast<const uint4 *>(ws.nf + r0 + 4);
            c[0] = lo.x, c[1] = lo.y, c[2] = lo.z, c[3] = lo.w, c[4] = hi.x, c[5] = hi.y, c[6] = hi.z, c[7] = hi.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = (r0 + j < g.npieces) ? ws.nf[r0 + j] : 0u;
        }
        unsigned long long sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += c[j];
        const unsigned long long incl = warp_incl_scan64(sum, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        unsigned long long before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned long long a = s_warp[k];
            if (k < warp) before += a;
            total += a;
        }
        if (warp == 0) {
            const unsigned long long e = lookback(ws.status_f, tile, total, lane);
            if (lane == 0) {
                s_excl = e;
                if (tile == g.nscan - 1) ws.header->total_f = e + total;
            }
        }
        __syncthreads();
        if (r0 < g.npieces) ws.f8[r0 >> 3] = s_excl + before + incl - sum;
    }
}

// ---------------------------------------------------------------------------------------------
// Pass B: k_faces.  Replaces gen_faces_kernel (:140-209).
//
// A warp takes 32 consecutive (row, piece) pairs in voxel-major order; lane l owns the 128 cells of
// piece l.  It recomputes the eight crossing masks of each of its four words,
//   q0 (x,y) x-edges   q1 (x,y) y-edges   q2 (x,y) z-edges   q3 (x+1,y) y-edges
//   q4 (x+1,y) z-edges q5 (x,y+1) x-edges q6 (x,y+1) z-edges q7 (x+1,y+1) z-edges
// carries the id of the first crossing of each mask along the piece (table entry + popcounts), and
// parks {corner words, masks, first ids} of every word with active cells in shared memory.  The sparse
// work then runs lane-balanced: one active cell per lane (case, triangle count), one triangle per lane
// (three ranks, 12-byte store).  Cube edge e -> (q, dz) follows the owner map of :178-192:
//   e: 0 1 2 3 4 5 6 7 8 9 10 11
//   q: 0 3 5 1 0 3 5 1 2 4  7  6      dz = 1 for e in 4..7 (the edge sits at sample z+1)
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kEdgeToMask = (0ull << 0) | (3ull << 3) | (5ull << 6) | (1ull << 9) | (0ull << 12) | (3ull << 15) |
                                 (5ull << 18) | (1ull << 21) | (2ull << 24) | (4ull << 27) | (7ull << 30) |
                                 (6ull << 33);
constexpr int kFaceWarps = 4;
constexpr int kCellCap = 1024;
constexpr int kTriBatch = 160;  // triangles of one batch of 32 cells (<= 5 each)

struct FaceScratch {
    uint4 corner[kFaceGroup * 4][2];  // per word slot: {a, b, c, d} at z and at z+1
    uint2 rm[kFaceGroup * 4][8];      // per word slot and mask q: {crossing mask, id of its first crossing}
    uint16_t cell[kCellCap];          // slot<<5 | bit
    uint32_t tri[kTriBatch];          // slot<<5 | bit | three (q | dz<<3) nibbles << 12
};

constexpr int kFaceSmemBytes = 256 * (int)sizeof(uint64_t) + kFaceWarps * (int)sizeof(FaceScratch);

// 8-bit cube case of the cell at bit i (corner order of :168-176)
__device__ __forceinline__ uint32_t cube_case_at(const uint4 &w, const uint4 &w2, int i) {
    return ((w.x >> i) & 1u) | (((w.y >> i) & 1u) << 1) | (((w.z >> i) & 1u) << 2) | (((w.w >> i) & 1u) << 3) |
           (((w2.x >> i) & 1u) << 4) | (((w2.y >> i) & 1u) << 5) | (((w2.z >> i) & 1u) << 6) | (((w2.w >> i) & 1u) << 7);
}

__global__ void __launch_bounds__(kFaceWarps * 32, 3)
    k_faces(McGeom g, McWorkspace ws, int32_t vbase, int32_t *__restrict__ faces) {
    extern __shared__ __align__(16) unsigned char face_smem[];
    uint64_t *s_table = reinterpret_cast<uint64_t *>(face_smem);  // per case: up to 15 nibbles (q | dz<<3), nibble 15 = #triangles
    FaceScratch *s_scratch = reinterpret_cast<FaceScratch *>(face_smem + 256 * sizeof(uint64_t));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        const uint64_t t = c_case_table[c];
        const uint32_t n = (uint32_t)(t >> 60);
        uint64_t out = (uint64_t)n << 60;
        for (uint32_t j = 0; j < 3 * n; ++j) {
            const uint32_t e = (uint32_t)(t >> (4 * j)) & 15u;
            out |= (((kEdgeToMask >> (3 * e)) & 7ull) | ((e & 12u) == 4u ? 8ull : 0ull)) << (4 * j);
        }
        s_table[c] = out;
    }
    __syncthreads();
    FaceScratch &sc = s_scratch[warp];

    const int64_t np = g.np, bstride = 4 * np, plane_pieces = g.ry * np;
    const int64_t ngroups = (g.npieces + kFaceGroup - 1) / kFaceGroup;
    const int64_t nwarps = (int64_t)gridDim.x * kFaceWarps;

    for (int64_t grp = (int64_t)blockIdx.x * kFaceWarps + warp; grp < ngroups; grp += nwarps) {
        const int64_t gi = grp * kFaceGroup + lane;
        const bool valid = gi < g.npieces;
        int64_t row, x, y;
        int p;
        if (g.npieces <= 0x7fffffffll) {
            const uint32_t rw = (uint32_t)gi / (uint32_t)np;
            p = (int)((uint32_t)gi - rw * (uint32_t)np);
            const uint32_t xx = rw / (uint32_t)g.ry;
            row = rw, x = xx, y = rw - xx * (uint32_t)g.ry;
        } else {
            row = gi / np;
            p = (int)(gi - row * np);
            x = row / g.ry;
            y = row - x * g.ry;
        }
        const bool hc = valid && (x + 1 < g.rx) && (y + 1 < g.ry);
        uint4 ta = make_uint4(0, 0, 0, 0), tb = ta, td = ta, tcc = ta;
        if (hc) {
            ta = ws.ptab[gi];
            tb = ws.ptab[gi + plane_pieces];
            td = ws.ptab[gi + np];
            tcc = ws.ptab[gi + plane_pieces + np];
        }
        const uint32_t nf = hc ? ta.w : 0u;
        if (!__any_sync(kFull, nf != 0u)) continue;

        // entries of the next piece of the same rows: the cell at bit 127 reads its z+1 edges there
        uint4 nxa = make_uint4(__shfl_down_sync(kFull, ta.x, 1), __shfl_down_sync(kFull, ta.y, 1), 0, 0);
        uint32_t tbn_y = __shfl_down_sync(kFull, tb.y, 1), tdn_x = __shfl_down_sync(kFull, td.x, 1);
        if (lane == 31 && hc && p + 1 < np) {
            const uint4 t0 = ws.ptab[gi + 1];
            nxa.x = t0.x, nxa.y = t0.y;
            tbn_y = ws.ptab[gi + 1 + plane_pieces].y;
            tdn_x = ws.ptab[gi + 1 + np].x;
        }
        unsigned long long fbase = 0;
        if (lane == 0) fbase = ws.f8[grp * (kFaceGroup / 8)];
        fbase = __shfl_sync(kFull, fbase, 0);
        const uint32_t finc = warp_incl_scan(nf, lane);

        uint32_t actw[4] = {0, 0, 0, 0};
        uint32_t nact = 0;
        uint4 last_w = make_uint4(0, 0, 0, 0), last_w2 = last_w;  // corner words of word 3 (patch below)
        if (nf) {
            const uint32_t *pa = ws.bits + row * bstride + 4 * p;
            const uint32_t *pb = pa + g.ry * bstride, *pd = pa + bstride, *pc = pb + bstride;
            const uint4 a4 = __ldg(reinterpret_cast<const uint4 *>(pa)), b4 = __ldg(reinterpret_cast<const uint4 *>(pb));
            const uint4 c4 = __ldg(reinterpret_cast<const uint4 *>(pc)), d4 = __ldg(reinterpret_cast<const uint4 *>(pd));
            const bool more = p + 1 < np;
            const uint32_t a[5] = {a4.x, a4.y, a4.z, a4.w, more ? __ldg(pa + 4) : 0u};
            const uint32_t b[5] = {b4.x, b4.y, b4.z, b4.w, more ? __ldg(pb + 4) : 0u};
            const uint32_t c[5] = {c4.x, c4.y, c4.z, c4.w, more ? __ldg(pc + 4) : 0u};
            const uint32_t d[5] = {d4.x, d4.y, d4.z, d4.w, more ? __ldg(pd + 4) : 0u};
            uint32_t run[8] = {ta.x, ta.y, ta.z, tb.y, tb.z, td.x, td.z, tcc.z};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t A = a[w], B = b[w], C = c[w], D = d[w];
                const uint32_t A2 = __funnelshift_r(A, a[w + 1], 1), B2 = __funnelshift_r(B, b[w + 1], 1);
                const uint32_t C2 = __funnelshift_r(C, c[w + 1], 1), D2 = __funnelshift_r(D, d[w + 1], 1);
                const uint32_t zv = low_mask(g.rz - 1 - ((int64_t)p * kTileZ + 32 * w));
                // pad bits are zero, and a masked-out crossing can only sit above every valid cell of the row,
                // so the masks need no z-validity here (the ids were assigned with it, k_tile)
                const uint32_t m[8] = {A ^ B, A ^ D, (A ^ A2) & zv, B ^ C, (B ^ B2) & zv, D ^ C, (D ^ D2) & zv, (C ^ C2) & zv};
                const uint32_t any = A | B | C | D | A2 | B2 | C2 | D2, all = A & B & C & D & A2 & B2 & C2 & D2;
                const uint32_t act = (any & ~all) & zv;  // :154,168-176
                if (act) {
                    const int slot = lane * 4 + w;
                    sc.corner[slot][0] = make_uint4(A, B, C, D);
                    sc.corner[slot][1] = make_uint4(A2, B2, C2, D2);
#pragma unroll
                    for (int q = 0; q < 8; q += 2)
                        *reinterpret_cast<uint4 *>(&sc.rm[slot][q]) = make_uint4(m[q], run[q], m[q + 1], run[q + 1]);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) run[q] += __popc(m[q]);
                actw[w] = act;
                nact += __popc(act);
                if (w == 3) {
                    last_w = make_uint4(A, B, C, D);
                    last_w2 = make_uint4(A2, B2, C2, D2);
                }
            }
        }
        const uint32_t cincl = warp_incl_scan(nact, lane);
        const uint32_t ncell = __shfl_sync(kFull, cincl, 31);
        unsigned long long frun = fbase;
        __syncwarp();

        for (uint32_t c0 = 0; c0 < ncell; c0 += kCellCap) {
            {
                uint32_t pos = cincl - nact - c0;  // wraps below the chunk: filtered by the range test
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    for (uint32_t rem = actw[w]; rem; ++pos) {
                        const int i = __ffs(rem) - 1;
                        rem &= rem - 1;
                        if (pos < (uint32_t)kCellCap) sc.cell[pos] = (uint16_t)(((lane * 4 + w) << 5) | i);
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
                    tt = s_table[cube_case_at(sc.corner[e >> 5][0], sc.corner[e >> 5][1], e & 31u)];
                    nt = (uint32_t)(tt >> 60);
                }
                const uint32_t tincl = warp_incl_scan(nt, lane);
                const uint32_t btot = __shfl_sync(kFull, tincl, 31);
                uint32_t tp = tincl - nt;
                for (uint32_t t = 0; t < nt; ++t, tt >>= 12) sc.tri[tp++] = e | (((uint32_t)tt & 0xfffu) << 12);
                __syncwarp();
                // one triangle per lane: rank its three edges, 12-byte stores (:194-208)
                for (uint32_t j = lane; j < btot; j += 32) {
                    const uint32_t ent = sc.tri[j];
                    const uint32_t slot = (ent >> 5) & 127u, i = ent & 31u;
                    const uint32_t lt = (1u << i) - 1u;
                    int32_t *out = faces + (frun + j) * 3ull;
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        const uint32_t nib = (ent >> (12 + 4 * cc)) & 15u;
                        // crossings strictly below sample z (+dz): for dz = 1 the bit at z counts too; at i = 31
                        // that makes the whole word count, i.e. the first id of the next word
                        const uint2 rk = sc.rm[slot][nib & 7u];
                        out[cc] = vbase + (int32_t)(rk.y + __popc(rk.x & (lt | ((nib >> 3) << i))));
                    }
                }
                __syncwarp();
                frun += btot;
            }
        }

        // the last cell of a piece (bit 127) has its z+1 x-/y-edges in the NEXT piece, which is numbered by
        // another tile: overwrite those indices with that piece's table entries (its bit 0 is rank 0)
        if (actw[3] >> 31) {
            const uint64_t tt0 = s_table[cube_case_at(last_w, last_w2, 31)];
            const uint32_t nt = (uint32_t)(tt0 >> 60);
            uint64_t tt = tt0;
            int32_t *out = faces + (fbase + finc - nt) * 3ull;
            for (uint32_t t = 0; t < nt; ++t)
                for (int cc = 0; cc < 3; ++cc, tt >>= 4) {
                    const uint32_t nib = (uint32_t)tt & 15u;
                    if (nib & 8u) {
                        const uint32_t q = nib & 7u;
                        out[t * 3 + cc] = vbase + (int32_t)(q == 0 ? nxa.x : (q == 1 ? nxa.y : (q == 3 ? tbn_y : tdn_x)));
                    }
                }
        }
        __syncwarp();
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

template <bool TMA>
void launch_tile_kernel(const CUtensorMap &map, const float *grid, const McGeom &g, const McWorkspace &ws,
                        const McEmitParams &p, float *verts, int64_t vcap, int mode, cudaStream_t s) {
    static const bool attr = [] {
        cudaFuncSetAttribute(k_tile<TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileSmemBytes);
        return true;
    }();
    (void)attr;
    const int64_t cap = (int64_t)sm_count() * 2;
    const unsigned blocks = (unsigned)(g.ntiles < cap ? g.ntiles : cap);
    k_tile<TMA><<<blocks, kTileThreads, kTileSmemBytes, s>>>(map, grid, g, ws, p, verts,
                                                             (unsigned long long)(vcap > 0 ? vcap : 0), mode);
}

}  // namespace

const char *tile_pass_error() { return g_tile_error; }

void launch_tile_pass(const float *grid, const McGeom &g, const McWorkspace &ws, const McEmitParams &p, float *verts,
                      int64_t vertex_capacity, int mode, cudaStream_t s) {
    g_tile_error = nullptr;
    if (g.ntiles <= 0) return;
    if (!verts) vertex_capacity = 0;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    // TMA needs a 16-byte aligned base and 16-byte multiples as row / plane strides
    bool tma = !force_generic() && (g.rz % 4 == 0) && ((reinterpret_cast<uintptr_t>(grid) & 15) == 0);
    if (tma) {
        EncodeTiledFn enc = encode_tiled();
        if (!enc) {
            tma = false;
        } else {
            const cuuint64_t dims[3] = {(cuuint64_t)g.rz, (cuuint64_t)g.ry, (cuuint64_t)g.rx};
            const cuuint64_t strides[2] = {(cuuint64_t)g.rz * 4, (cuuint64_t)g.ry * (cuuint64_t)g.rz * 4};
            const cuuint32_t box[3] = {kBoxZ, kTileY + 1, kTileX + 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(grid), dims, strides, box,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                g_tile_error = "cuTensorMapEncodeTiled failed";
                return;
            }
        }
    }
    if (tma)
        launch_tile_kernel<true>(map, grid, g, ws, p, verts, vertex_capacity, mode, s);
    else
        launch_tile_kernel<false>(map, grid, g, ws, p, verts, vertex_capacity, mode, s);
}

void launch_face_scan(const McGeom &g, const McWorkspace &ws, cudaStream_t s) {
    if (g.nscan <= 0) return;
    const int64_t cap = (int64_t)sm_count() * 4;
    k_fscan<<<(unsigned)(g.nscan < cap ? g.nscan : cap), 256, 0, s>>>(g, ws);
}

void launch_faces(const McGeom &g, const McWorkspace &ws, const McEmitParams &p, int32_t *faces, cudaStream_t s) {
    if (g.npieces <= 0) return;
    const int64_t groups = (g.npieces + kFaceGroup - 1) / kFaceGroup;
    const int64_t want = (groups + kFaceWarps - 1) / kFaceWarps, cap = (int64_t)sm_count() * 3;
    static const bool attr = [] {
        cudaFuncSetAttribute(k_faces, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaceSmemBytes);
        return true;
    }();
    (void)attr;
    k_faces<<<(unsigned)(want < cap ? want : cap), kFaceWarps * 32, kFaceSmemBytes, s>>>(g, ws, p.vertex_id_base, faces);
}

// Multi-GPU: install the next shard's first-plane table as this shard's halo-plane numbering.
__global__ void k_import_halo(uint4 *__restrict__ dst, const uint4 *__restrict__ src, int64_t n, uint32_t delta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint4 t = src[i];
        dst[i] = make_uint4(t.x + delta, t.y + delta, t.z + delta, t.w);
    }
}

void launch_import_halo(uint4 *halo_entries, const uint32_t *table_in, int64_t n, uint32_t delta, cudaStream_t s) {
    if (n <= 0) return;
    k_import_halo<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(halo_entries, reinterpret_cast<const uint4 *>(table_in), n, delta);
}

}  // namespace p3d
