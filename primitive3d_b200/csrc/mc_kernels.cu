// primitive3d_b200/csrc/mc_kernels.cu -- see mc_kernels.cuh for the design.
// Reference citations are into /root/reference/src/prim3d/Utility/marching_cubes.cu.
#include "mc_kernels.cuh"

#include <cstdlib>

#include "mc_case_table.h"
#include "scan_utils.cuh"

namespace p3d {

// Bourke case table, one packed word per case (nibble i = i-th edge index, nibble 15 =
// #triangles).  Lives in constant memory and is staged into shared memory once per
// persistent CTA because the per-cell lookups are lane-divergent.
__constant__ uint64_t c_case_table[256] = P3D_MC_CASE_TABLE_INIT;

// Bits z of word w that own a +z edge / a cell: z + 1 < rz.
__device__ __forceinline__ uint32_t zvalid_mask(int w, int64_t rz) {
    const int64_t n = rz - 1 - 32 * (int64_t)w;
    return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << (int)n) - 1u));
}

// ---------------------------------------------------------------------------------------------
// K1: classify.  inside = value > thresh  (marching_cubes.cu:25,31,37,43,50-57).
// ---------------------------------------------------------------------------------------------

// Flat fast path (rz % 32 == 0, 16-byte aligned grid): a warp turns 1024 consecutive samples
// (eight coalesced 512-byte float4 loads, all in flight together) into 32 bit words and
// stores them as one 128-byte line.
__global__ void __launch_bounds__(256) k_classify_flat(const float4 *__restrict__ g4, uint32_t *__restrict__ bits,
                                                       int64_t nchunks, float thresh) {
    const int lane = threadIdx.x & 31;
    const int sub = lane & 7;  // position inside the 8-lane group that assembles one word
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < nchunks; c += nwarps) {
        const float4 *p = g4 + c * 256 + lane;
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + u * 32);
        uint32_t mine = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t w = (v[u].x > thresh ? 1u : 0u) | (v[u].y > thresh ? 2u : 0u) | (v[u].z > thresh ? 4u : 0u) |
                         (v[u].w > thresh ? 8u : 0u);
            w <<= 4 * sub;
            w |= __shfl_xor_sync(kFull, w, 1);
            w |= __shfl_xor_sync(kFull, w, 2);
            w |= __shfl_xor_sync(kFull, w, 4);
            if (sub == u) mine = w;  // lane keeps word (u = sub, group = lane>>3)
        }
        bits[c * 32 + sub * 4 + (lane >> 3)] = mine;
    }
}

// General path (any rz / alignment): a warp handles 32 words of one row with scalar coalesced
// loads and ballots; samples past the end of the row read as outside (pad bits are 0).
__global__ void __launch_bounds__(256) k_classify_rows(const float *__restrict__ grid, uint32_t *__restrict__ bits,
                                                       int64_t nrows, int64_t rz, int wz, int pieces, float thresh) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t items = nrows * pieces;
    for (int64_t it = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < items; it += nwarps) {
        const int64_t row = it / pieces;
        const int p = (int)(it - row * pieces);
        const float *src = grid + row * rz;
        uint32_t mine = 0;
#pragma unroll 1
        for (int j0 = 0; j0 < 32; j0 += 8) {
            bool in[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int64_t z = ((int64_t)(p * 32 + j0 + u) << 5) + lane;
                in[u] = (z < rz) && (__ldcs(src + (z < rz ? z : 0)) > thresh);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t b = __ballot_sync(kFull, in[u]);
                if (lane == j0 + u) mine = b;
            }
        }
        const int w = p * 32 + lane;
        if (w < wz) bits[row * wz + w] = mine;
    }
}

void launch_classify(const float *grid, const McGeom &g, float thresh, uint32_t *bits, cudaStream_t s) {
    const int sms = sm_count();
    const int64_t nrows = g.rx * g.ry;
    const int64_t n = nrows * g.rz;
    const bool flat = (g.rz % 32 == 0) && ((reinterpret_cast<uintptr_t>(grid) & 15) == 0);
    if (flat) {
        const int64_t nchunks = n / 1024;
        if (nchunks > 0) {
            const int64_t want = (nchunks + 7) / 8;
            const int blocks = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
            k_classify_flat<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4 *>(grid), bits, nchunks, thresh);
        }
        const int64_t rem = n - nchunks * 1024;  // a multiple of 32 samples, < 1024
        if (rem > 0)
            k_classify_rows<<<1, 32, 0, s>>>(grid + nchunks * 1024, bits + nchunks * 32, 1, rem, (int)(rem / 32), 1,
                                             thresh);
    } else {
        const int64_t items = nrows * g.pieces;
        const int64_t want = (items + 7) / 8;
        const int blocks = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
        k_classify_rows<<<blocks, 256, 0, s>>>(grid, bits, nrows, g.rz, g.wz, g.pieces, thresh);
    }
}

// ---------------------------------------------------------------------------------------------
// Shared row machinery of K2 and K3.
// A warp owns row r = (x,y).  Lane l holds word w = 32*piece + l of the four bit rows
//   a = (x,y)  b = (x+1,y)  c = (x+1,y+1)  d = (x,y+1)
// i.e. the reference's cube corners 0,1,2,3 at z and 4,5,6,7 at z+1 (marching_cubes.cu:50-57).
// ---------------------------------------------------------------------------------------------
struct RowPtrs {
    const uint32_t *a, *b, *c, *d;
    bool has_x, has_y;  // x+1 < rx, y+1 < ry
    int64_t x, y;
};

__device__ __forceinline__ RowPtrs row_ptrs(const uint32_t *bits, const McGeom &g, int64_t row) {
    RowPtrs r;
    int64_t x, y;
    if (g.rx * g.ry <= 0x7fffffffll) {
        x = (uint32_t)row / (uint32_t)g.ry;
        y = (uint32_t)row - (uint32_t)x * (uint32_t)g.ry;
    } else {
        x = row / g.ry;
        y = row - x * g.ry;
    }
    r.has_x = x + 1 < g.rx;
    r.has_y = y + 1 < g.ry;
    r.a = bits + row * g.wz;
    r.b = r.a + (r.has_x ? g.ry * (int64_t)g.wz : 0);
    r.d = r.a + (r.has_y ? g.wz : 0);
    r.c = r.b + (r.has_y ? g.wz : 0);
    r.x = x;
    r.y = y;
    return r;
}

struct Piece {
    uint32_t a, b, c, d;      // this lane's words
    uint32_t a2, b2, c2, d2;  // the same rows shifted down by one sample: bit i = sample z+1
    uint32_t an, bn, cn, dn;  // next words (lane+1's; lane 31 loads them)
};

__device__ __forceinline__ uint32_t next_word(const uint32_t *row, uint32_t mine, int w, int wz, int lane) {
    uint32_t n = __shfl_down_sync(kFull, mine, 1);
    if (lane == 31) n = (w + 1 < wz) ? __ldg(row + w + 1) : 0u;
    return n;
}

__device__ __forceinline__ Piece load_piece(const RowPtrs &r, int w, int wz, int lane) {
    Piece p;
    const bool in = w < wz;
    p.a = in ? __ldg(r.a + w) : 0u;
    p.b = in ? __ldg(r.b + w) : 0u;
    p.c = in ? __ldg(r.c + w) : 0u;
    p.d = in ? __ldg(r.d + w) : 0u;
    p.an = next_word(r.a, p.a, w, wz, lane);
    p.bn = next_word(r.b, p.b, w, wz, lane);
    p.cn = next_word(r.c, p.c, w, wz, lane);
    p.dn = next_word(r.d, p.d, w, wz, lane);
    p.a2 = (p.a >> 1) | (p.an << 31);
    p.b2 = (p.b >> 1) | (p.bn << 31);
    p.c2 = (p.c >> 1) | (p.cn << 31);
    p.d2 = (p.d >> 1) | (p.dn << 31);
    return p;
}

// Cells of the row whose 8 corners are not all equal (z+1 < rz, x+1 < rx, y+1 < ry).
__device__ __forceinline__ uint32_t active_cells(const Piece &p, uint32_t zv, bool cells) {
    const uint32_t any = p.a | p.b | p.c | p.d | p.a2 | p.b2 | p.c2 | p.d2;
    const uint32_t all = p.a & p.b & p.c & p.d & p.a2 & p.b2 & p.c2 & p.d2;
    return cells ? ((any & ~all) & zv) : 0u;
}

// 8-bit cube case of the cell at bit i (corner order of marching_cubes.cu:168-176) from the packed
// words {a,b,c,d} and their one-sample-shifted copies
__device__ __forceinline__ uint32_t cube_case_at(const uint4 &w, const uint4 &w2, int i) {
    return ((w.x >> i) & 1u) | (((w.y >> i) & 1u) << 1) | (((w.z >> i) & 1u) << 2) | (((w.w >> i) & 1u) << 3) |
           (((w2.x >> i) & 1u) << 4) | (((w2.y >> i) & 1u) << 5) | (((w2.z >> i) & 1u) << 6) | (((w2.w >> i) & 1u) << 7);
}

// ---------------------------------------------------------------------------------------------
// K2a: per-row counts from the bit words.  Replaces count_vertices_faces_kernel (:4-68) and its
// two global atomic counters.  rowv[row] = {nx, ny, nz, nf} (turned into offsets by K2b).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_count(McGeom g, McWorkspace ws) {
    __shared__ uint8_t s_ntri[256];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = (uint8_t)(c_case_table[i] >> 60);
    __syncthreads();
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < g.owned_rows; row += nwarps) {
        const RowPtrs r = row_ptrs(ws.bits, g, row);
        const bool cells = r.has_x && r.has_y;
        uint32_t nx = 0, ny = 0, nz = 0, nf = 0;
        for (int pc = 0; pc < g.pieces; ++pc) {
            const int w = pc * kPieceWords + lane;
            const Piece p = load_piece(r, w, g.wz, lane);
            const uint32_t zv = zvalid_mask(w, g.rz);
            if (r.has_x) nx += __popc(p.a ^ p.b);       // :29-33
            if (r.has_y) ny += __popc(p.a ^ p.d);       // :35-39
            nz += __popc((p.a ^ p.a2) & zv);            // :41-45
            uint32_t act = active_cells(p, zv, cells);  // :48-66
            const uint4 w4 = make_uint4(p.a, p.b, p.c, p.d), w42 = make_uint4(p.a2, p.b2, p.c2, p.d2);
            while (act) {
                const int i = __ffs(act) - 1;
                act &= act - 1;
                nf += s_ntri[cube_case_at(w4, w42, i)];
            }
        }
        // one packed reduction: nx,ny,nz <= rz (< 2^31 / 32 per lane) ... keep them separate but cheap
        const unsigned long long lo = warp_sum64(((unsigned long long)ny << 32) | nx);
        const unsigned long long hi = warp_sum64(((unsigned long long)nf << 32) | nz);
        if (lane == 0) ws.rowv[row] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
    }
}

// ---------------------------------------------------------------------------------------------
// K2b: exclusive scan of the per-row counts -- single pass, decoupled look-back over tiles of
// 2048 rows (no CUB / thrust).  rowv[row] becomes {vx, vy, vz, nf}: first vertex id of the row's
// x-, y- and z-edge groups; rowf[row] = first face of the row.  Totals go to the header.
// ---------------------------------------------------------------------------------------------
constexpr int kScanRowsPerThread = 8;
constexpr int kScanTile = 256 * kScanRowsPerThread;

__global__ void __launch_bounds__(256) k_row_scan(McGeom g, McWorkspace ws, int64_t num_scan_tiles) {
    __shared__ unsigned long long s_warp[2][8];
    __shared__ unsigned long long s_excl[2];
    __shared__ unsigned int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&ws.header->ticket, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= num_scan_tiles) break;
        const int64_t r0 = tile * kScanTile + (int64_t)threadIdx.x * kScanRowsPerThread;
        uint4 c[kScanRowsPerThread];
        unsigned long long sv = 0, sf = 0;
#pragma unroll
        for (int j = 0; j < kScanRowsPerThread; ++j) {
            c[j] = (r0 + j < g.owned_rows) ? ws.rowv[r0 + j] : make_uint4(0, 0, 0, 0);
            sv += (unsigned long long)c[j].x + c[j].y + c[j].z;
            sf += c[j].w;
        }
        const unsigned long long iv = warp_incl_scan64(sv, lane), jf = warp_incl_scan64(sf, lane);
        if (lane == 31) {
            s_warp[0][warp] = iv;
            s_warp[1][warp] = jf;
        }
        __syncthreads();
        unsigned long long bv = 0, bf = 0, tv = 0, tf = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const unsigned long long a = s_warp[0][w], b = s_warp[1][w];
            if (w < warp) {
                bv += a;
                bf += b;
            }
            tv += a;
            tf += b;
        }
        if (warp < 2) {  // warp 0 looks back over vertex aggregates, warp 1 over face aggregates
            const unsigned long long agg = warp == 0 ? tv : tf;
            const unsigned long long e = lookback(warp == 0 ? ws.status_v : ws.status_f, tile, agg, lane);
            if (lane == 0) {
                s_excl[warp] = e;
                if (tile == num_scan_tiles - 1) (warp == 0 ? ws.header->total_v : ws.header->total_f) = e + agg;
            }
        }
        __syncthreads();
        unsigned long long v = s_excl[0] + bv + iv - sv, f = s_excl[1] + bf + jf - sf;
#pragma unroll
        for (int j = 0; j < kScanRowsPerThread; ++j) {
            if (r0 + j < g.owned_rows) {
                const uint32_t vx = (uint32_t)v;
                ws.rowv[r0 + j] = make_uint4(vx, vx + c[j].x, vx + c[j].x + c[j].y, c[j].w);
                ws.rowf[r0 + j] = f;
            }
            v += (unsigned long long)c[j].x + c[j].y + c[j].z;
            f += c[j].w;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// K3: emit vertices and faces.  Replaces gen_vertices_kernel (:70-138), gen_faces_kernel
// (:140-209) and the two ATen passes of the bounding-box epilogue (:298).
//
// A warp owns a row; lane l owns word l of the current 1024-sample piece.  Two packed 64-bit
// warp scans rank all eight crossing masks at once; the sparse work (one vertex per crossing
// edge, one triangle per table entry) is compacted into shared-memory lists so that it runs one
// item per lane with coalesced output stores.
// ---------------------------------------------------------------------------------------------

// Mask ids q used for edge ranking: which row's crossing mask numbers the edge.
//   q0 (x,y) x-edges   q1 (x,y) y-edges   q2 (x,y) z-edges   q3 (x+1,y) y-edges
//   q4 (x+1,y) z-edges q5 (x,y+1) x-edges q6 (x,y+1) z-edges q7 (x+1,y+1) z-edges
// Cube edge e -> (q, dz) following the owner map of marching_cubes.cu:178-192:
//   e: 0 1 2 3 4 5 6 7 8 9 10 11
//   q: 0 3 5 1 0 3 5 1 2 4  7  6      dz = 1 for e in 4..7 (the edge sits at sample z+1)
constexpr uint64_t kEdgeToMask = (0ull << 0) | (3ull << 3) | (5ull << 6) | (1ull << 9) | (0ull << 12) | (3ull << 15) |
                                 (5ull << 18) | (1ull << 21) | (2ull << 24) | (4ull << 27) | (7ull << 30) |
                                 (6ull << 33);

#ifndef P3D_EMIT_MINBLOCKS
#define P3D_EMIT_MINBLOCKS 4
#endif
#ifndef P3D_STRIP_MINBLOCKS
#define P3D_STRIP_MINBLOCKS 3
#endif
constexpr int kTriBatch = 160;  // triangles of one batch of 32 cells (<= 5 each)

struct WarpScratch {
    uint4 words[kPieceWords];       // {a, b, c, d}
    uint4 words2[kPieceWords];      // the same rows shifted by one sample (bit i = sample z+1)
    uint2 rank[8][kPieceWords];     // per mask q and word: {crossing mask, id of its first crossing}
    uint16_t list[kPieceWords * 32];  // compacted work items: vertex (z | axis<<10) or cell (z)
    uint32_t tri[kTriBatch];        // z | three (q | dz<<3) nibbles << 10
};

__global__ void __launch_bounds__(kRowsPerTile * 32, P3D_EMIT_MINBLOCKS) k_emit(const float *__restrict__ grid, McGeom g, McWorkspace ws,
                                                              McEmitParams prm, float *__restrict__ verts,
                                                              int32_t *__restrict__ faces) {
    __shared__ uint64_t s_table[256];  // per case: up to 15 nibbles (q | dz<<3), nibble 15 = #triangles
    __shared__ WarpScratch s_scratch[kRowsPerTile];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < 256; c += blockDim.x) {
        const uint64_t t = c_case_table[c];
        const uint32_t n = (uint32_t)(t >> 60);
        uint64_t out = (uint64_t)n << 60;
        for (uint32_t j = 0; j < 3 * n; ++j) {
            const uint32_t e = (uint32_t)(t >> (4 * j)) & 15u;
            const uint64_t nib = ((kEdgeToMask >> (3 * e)) & 7ull) | ((e & 12u) == 4u ? 8ull : 0ull);
            out |= nib << (4 * j);
        }
        s_table[c] = out;
    }
    __syncthreads();
    WarpScratch &sc = s_scratch[warp];
    const int64_t plane = g.ry * g.rz;
    const int64_t nwarps = (int64_t)gridDim.x * kRowsPerTile;

    for (int64_t row = (int64_t)blockIdx.x * kRowsPerTile + warp; row < g.owned_rows; row += nwarps) {
        const RowPtrs r = row_ptrs(ws.bits, g, row);
        const bool cells = r.has_x && r.has_y;
        const float fx = (float)(prm.x_origin + r.x);  // static_cast<float>(x), :107
        const float fy = (float)r.y;
        const float *grow = grid + row * g.rz;

        // first ids of the four rows' vertex groups (row table written by K2b / halo import)
        uint32_t run[8];
        {
            const uint4 t00 = ws.rowv[row];
            run[0] = t00.x;
            run[1] = t00.y;
            run[2] = t00.z;
            run[3] = run[4] = run[5] = run[6] = run[7] = 0;
            if (cells) {
                const uint4 t10 = ws.rowv[row + g.ry], t01 = ws.rowv[row + 1], t11 = ws.rowv[row + g.ry + 1];
                run[3] = t10.y;
                run[4] = t10.z;
                run[5] = t01.x;
                run[6] = t01.z;
                run[7] = t11.z;
            }
        }
        unsigned long long frun = ws.rowf[row];

        for (int pc = 0; pc < g.pieces; ++pc) {
            const int w = pc * kPieceWords + lane;
            const Piece p = load_piece(r, w, g.wz, lane);
            const uint32_t zv = zvalid_mask(w, g.rz);
            uint32_t m[8];
            m[0] = r.has_x ? (p.a ^ p.b) : 0u;
            m[1] = r.has_y ? (p.a ^ p.d) : 0u;
            m[2] = (p.a ^ p.a2) & zv;
            m[3] = cells ? (p.b ^ p.c) : 0u;
            m[4] = cells ? ((p.b ^ p.b2) & zv) : 0u;
            m[5] = cells ? (p.d ^ p.c) : 0u;
            m[6] = cells ? ((p.d ^ p.d2) & zv) : 0u;
            m[7] = cells ? ((p.c ^ p.c2) & zv) : 0u;
            const uint32_t act = active_cells(p, zv, cells);

            // two packed scans (12-bit fields; a field's inclusive sum is <= 1024)
            const unsigned long long ca = (unsigned long long)__popc(m[0]) | ((unsigned long long)__popc(m[1]) << 12) |
                                          ((unsigned long long)__popc(m[2]) << 24) | ((unsigned long long)__popc(m[3]) << 36) |
                                          ((unsigned long long)__popc(m[4]) << 48);
            const unsigned long long cb = (unsigned long long)__popc(m[5]) | ((unsigned long long)__popc(m[6]) << 12) |
                                          ((unsigned long long)__popc(m[7]) << 24) | ((unsigned long long)__popc(act) << 36);
            const unsigned long long ia = warp_incl_scan64(ca, lane), ib = warp_incl_scan64(cb, lane);
            const unsigned long long ta = __shfl_sync(kFull, ia, 31), tb = __shfl_sync(kFull, ib, 31);
            const unsigned long long ea = ia - ca, eb = ib - cb;

            __syncwarp();  // the previous piece's readers are done with the scratch
            sc.words[lane] = make_uint4(p.a, p.b, p.c, p.d);
            sc.words2[lane] = make_uint4(p.a2, p.b2, p.c2, p.d2);
            uint32_t first[3], tot[3], ex[3];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned long long e = q < 5 ? ea : eb, t = q < 5 ? ta : tb;
                const int sh = 12 * (q < 5 ? q : q - 5);
                const uint32_t exq = (uint32_t)(e >> sh) & 0xfffu, totq = (uint32_t)(t >> sh) & 0xfffu;
                sc.rank[q][lane] = make_uint2(m[q], run[q] + exq);
                if (q < 3) {
                    first[q] = run[q];
                    tot[q] = totq;
                    ex[q] = exq;
                }
                run[q] += totq;  // first id of the next piece
            }
            const uint32_t ncell = (uint32_t)(tb >> 36) & 0xfffu;
            const uint32_t cell_ex = (uint32_t)(eb >> 36) & 0xfffu;

            // ---- vertices on the row's own +x / +y / +z edges (gen_vertices_kernel) ----
            // rounds: as many whole axis groups as fit the 1024-entry list (all three unless the
            // piece is nearly all crossings)
            for (int q0 = 0; q0 < 3;) {
                uint32_t start[3] = {0, 0, 0};
                uint32_t cnt = 0;
                int q1 = q0;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (q == q1 && (q == q0 || cnt + tot[q] <= (uint32_t)(kPieceWords * 32))) {
                        start[q] = cnt;
                        uint32_t mm = m[q], pos = cnt + ex[q];
                        while (mm) {
                            const int i = __ffs(mm) - 1;
                            mm &= mm - 1;
                            sc.list[pos++] = (uint16_t)((lane << 5) | i | (q << 10));
                        }
                        cnt += tot[q];
                        q1 = q + 1;
                    }
                }
                __syncwarp();
                for (uint32_t k = lane; k < cnt; k += 32) {
                    const uint32_t ent = sc.list[k];
                    const uint32_t ax = ent >> 10;
                    const int64_t z = (int64_t)pc * (kPieceWords * 32) + (ent & 1023u);
                    const int64_t stride = ax == 0 ? plane : (ax == 1 ? g.rz : 1);
                    const float d0 = __ldg(grow + z);
                    const float d1 = __ldg(grow + z + stride);
                    // dt = (thresh - d_self) / (d_next - d_self), IEEE fp32, no contraction (:105)
                    const float dt = __fdiv_rn(__fsub_rn(prm.thresh, d0), __fsub_rn(d1, d0));
                    float px = fx, py = fy, pz = (float)z;
                    if (ax == 0) px = __fadd_rn(px, dt);
                    if (ax == 1) py = __fadd_rn(py, dt);
                    if (ax == 2) pz = __fadd_rn(pz, dt);
                    const uint32_t id = (ax == 0 ? first[0] - start[0] : (ax == 1 ? first[1] - start[1] : first[2] - start[2])) + k;
                    // vertices * scale + offset as two separately rounded ops (:298)
                    float *out = verts + (int64_t)id * 3;
                    out[0] = __fadd_rn(__fmul_rn(px, prm.scale[0]), prm.offset[0]);
                    out[1] = __fadd_rn(__fmul_rn(py, prm.scale[1]), prm.offset[1]);
                    out[2] = __fadd_rn(__fmul_rn(pz, prm.scale[2]), prm.offset[2]);
                }
                __syncwarp();
                q0 = q1;
            }

            // ---- faces of the row's cells, voxel-major, table order inside a cell (gen_faces_kernel) ----
            if (ncell) {
                uint32_t mm = act, pos = cell_ex;
                while (mm) {
                    const int i = __ffs(mm) - 1;
                    mm &= mm - 1;
                    sc.list[pos++] = (uint16_t)((lane << 5) | i);
                }
                __syncwarp();
                for (uint32_t k0 = 0; k0 < ncell; k0 += 32) {
                    // one cell per lane: case, triangle count, and one list entry per triangle
                    const uint32_t k = k0 + lane;
                    uint32_t nt = 0, zl = 0;
                    uint64_t tt = 0;
                    if (k < ncell) {
                        zl = sc.list[k];
                        tt = s_table[cube_case_at(sc.words[zl >> 5], sc.words2[zl >> 5], zl & 31)];
                        nt = (uint32_t)(tt >> 60);
                    }
                    const uint32_t tincl = warp_incl_scan(nt, lane);
                    const uint32_t btot = __shfl_sync(kFull, tincl, 31);
                    uint32_t tp = tincl - nt;
                    for (uint32_t t = 0; t < nt; ++t, tt >>= 12) sc.tri[tp++] = zl | (((uint32_t)tt & 0xfffu) << 10);
                    __syncwarp();
                    // one triangle per lane: rank its three edges, 12-byte coalesced stores
                    for (uint32_t j = lane; j < btot; j += 32) {
                        const uint32_t ent = sc.tri[j];
                        const uint32_t wl = (ent >> 5) & 31u, i = ent & 31u;
                        const uint32_t lt = (1u << i) - 1u;
                        int32_t *out = faces + (frun + j) * 3ull;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const uint32_t nib = (ent >> (10 + 4 * c)) & 15u;
                            // crossings strictly below sample z (+dz): for dz = 1 the bit at z counts too; at
                            // i = 31 that makes the whole word count, i.e. the first id of the next word
                            const uint2 rk = sc.rank[nib & 7u][wl];
                            out[c] = prm.vertex_id_base + (int32_t)(rk.y + __popc(rk.x & (lt | ((nib >> 3) << i))));
                        }
                    }
                    __syncwarp();
                    frun += btot;
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Strip kernels (lane = row).  A warp owns a strip of 32 consecutive rows y0..y0+31 of one plane x
// and walks their bit words in z order.  Each lane keeps the running ranks of ITS row in registers,
// so no warp scans are needed to number crossings; the rows (x,y+1) / (x+1,y+1) a lane's cells
// touch are simply the next lane's words (one shuffle; lane 31 loads row y0+32 itself).  Sparse
// work from all 32 rows is pooled in shared memory and processed one item per lane.
//   k_strip<false>  K2a: per-row counts {nx, ny, nz, nf}
//   k_strip<true>   K3 : vertices and faces
// ---------------------------------------------------------------------------------------------
constexpr int kVRing = 128;   // vertex ring entries (a round adds <= 3 per lane, <= 31 are carried over)
constexpr int kCellPool = 128;  // a round pools <= 4 cells per lane

struct StripScratch {
    uint4 words[32];            // per row-lane: {a, b, c, d} of the current word
    uint4 words2[32];           // the same shifted by one sample
    uint2 rank[8][32];          // per mask q and row-lane: {crossing mask, id of its first crossing}
    unsigned long long fbase[32];  // per row-lane: index of the row's next face at the start of this word
    uint32_t rowtris[32];       // per row-lane: triangles emitted so far for this word
    uint2 vring[kVRing];        // pooled vertices {id, z<<7 | lane<<2 | axis}
    uint32_t tri[kTriBatch];    // lane | bit<<5 | three (q|dz<<3) nibbles<<10 | offset in row<<22
    uint16_t cell[kCellPool];   // lane<<5 | bit
};

__device__ __forceinline__ void load_group(const uint32_t *row, bool ok, int gw, int wz, bool vec, uint32_t out[4]) {
    if (ok && vec && gw + 4 <= wz) {
        const uint4 t = __ldg(reinterpret_cast<const uint4 *>(row + gw));
        out[0] = t.x;
        out[1] = t.y;
        out[2] = t.z;
        out[3] = t.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = (ok && gw + j < wz) ? __ldg(row + gw + j) : 0u;
    }
}

template <bool EMIT>
__global__ void __launch_bounds__(256, EMIT ? P3D_STRIP_MINBLOCKS : 4)
    k_strip(const float *__restrict__ grid, McGeom g, McWorkspace ws, McEmitParams prm, float *__restrict__ verts,
            int32_t *__restrict__ faces) {
    __shared__ uint64_t s_table[256];  // EMIT: nibbles (q | dz<<3), nibble 15 = #triangles; else only the counts are used
    __shared__ StripScratch s_scratch[EMIT ? 8 : 1];

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
    StripScratch &sc = s_scratch[EMIT ? warp : 0];

    const int wz = g.wz;
    const bool vec = (wz & 3) == 0;
    const int64_t plane = g.ry * g.rz;
    const uint32_t nsy = (uint32_t)((g.ry + 31) / 32);
    const int64_t nstrips = g.owned_x * (int64_t)nsy;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);

    for (int64_t strip = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; strip < nstrips; strip += nwarps) {
        int64_t x;
        int64_t y0;
        if (nstrips <= 0x7fffffffll) {
            const uint32_t xs = (uint32_t)strip / nsy;
            x = xs;
            y0 = (int64_t)((uint32_t)strip - xs * nsy) * 32;
        } else {
            x = strip / nsy;
            y0 = (strip - x * nsy) * 32;
        }
        const int64_t y = y0 + lane;
        const bool valid = y < g.ry;
        const bool has_x = x + 1 < g.rx;
        const bool has_y = valid && (y + 1 < g.ry);
        const uint32_t hx = (valid && has_x) ? 0xffffffffu : 0u;
        const uint32_t hy = has_y ? 0xffffffffu : 0u;
        const uint32_t hc = (has_x && has_y) ? 0xffffffffu : 0u;
        const int64_t row = x * g.ry + y;
        const uint32_t *pa = ws.bits + row * wz;
        const uint32_t *pb = pa + g.ry * (int64_t)wz;
        const bool halo = (lane == 31) && (y0 + 32 < g.ry);  // lane 31 also loads row y0+32

        // per-row running state
        uint32_t run[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned long long frun = 0;
        uint32_t nx = 0, ny = 0, nz = 0, nf = 0;
        if (EMIT && valid) {
            const uint4 t00 = ws.rowv[row];
            run[0] = t00.x;
            run[1] = t00.y;
            run[2] = t00.z;
            if (hc) {
                const uint4 t10 = ws.rowv[row + g.ry], t01 = ws.rowv[row + 1], t11 = ws.rowv[row + g.ry + 1];
                run[3] = t10.y;
                run[4] = t10.z;
                run[5] = t01.x;
                run[6] = t01.z;
                run[7] = t11.z;
            }
            frun = ws.rowf[row];
        }
        uint32_t vhead = 0, vcount = 0;  // vertex ring (warp-uniform)
        const float fx = (float)(prm.x_origin + x);  // static_cast<float>(x), :107
        const int64_t row0 = x * g.ry + y0;

        auto flush_vertices = [&](uint32_t n) {  // the first n (<= 32) ring entries, one per lane
            if ((uint32_t)lane < n) {
                const uint2 ent = sc.vring[(vhead + lane) & (kVRing - 1)];
                const uint32_t ax = ent.y & 3u, l = (ent.y >> 2) & 31u;
                const int64_t z = ent.y >> 7;
                const float *src = grid + (row0 + l) * g.rz + z;
                const float d0 = __ldg(src);
                const float d1 = __ldg(src + (ax == 0 ? plane : (ax == 1 ? g.rz : 1)));
                // dt = (thresh - d_self) / (d_next - d_self), IEEE fp32, no contraction (:105)
                const float dt = __fdiv_rn(__fsub_rn(prm.thresh, d0), __fsub_rn(d1, d0));
                float px = fx, py = (float)(y0 + l), pz = (float)z;
                if (ax == 0) px = __fadd_rn(px, dt);
                if (ax == 1) py = __fadd_rn(py, dt);
                if (ax == 2) pz = __fadd_rn(pz, dt);
                // vertices * scale + offset as two separately rounded ops (:298)
                float *out = verts + (int64_t)ent.x * 3;
                out[0] = __fadd_rn(__fmul_rn(px, prm.scale[0]), prm.offset[0]);
                out[1] = __fadd_rn(__fmul_rn(py, prm.scale[1]), prm.offset[1]);
                out[2] = __fadd_rn(__fmul_rn(pz, prm.scale[2]), prm.offset[2]);
            }
            vhead += n;
            vcount -= n;
        };

        uint32_t na[4], nb[4], nd[4] = {0, 0, 0, 0}, nc[4] = {0, 0, 0, 0};
        load_group(pa, valid, 0, wz, vec, na);
        load_group(pb, valid && has_x, 0, wz, vec, nb);
        if (lane == 31) {
            load_group(pa + wz, halo, 0, wz, vec, nd);
            load_group(pb + wz, halo && has_x, 0, wz, vec, nc);
        }

        for (int gw = 0; gw < wz; gw += 4) {
            uint32_t a[5], b[5], d[5], c[5];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a[j] = na[j];
                b[j] = nb[j];
                d[j] = __shfl_down_sync(kFull, na[j], 1);
                c[j] = __shfl_down_sync(kFull, nb[j], 1);
                if (lane == 31) {
                    d[j] = nd[j];
                    c[j] = nc[j];
                }
            }
            // prefetch the next group; its first word also closes this group's last word
            load_group(pa, valid, gw + 4, wz, vec, na);
            load_group(pb, valid && has_x, gw + 4, wz, vec, nb);
            if (lane == 31) {
                load_group(pa + wz, halo, gw + 4, wz, vec, nd);
                load_group(pb + wz, halo && has_x, gw + 4, wz, vec, nc);
            }
            a[4] = na[0];
            b[4] = nb[0];
            d[4] = __shfl_down_sync(kFull, na[0], 1);
            c[4] = __shfl_down_sync(kFull, nb[0], 1);
            if (lane == 31) {
                d[4] = nd[0];
                c[4] = nc[0];
            }

#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int w = gw + j;
                if (w >= wz) break;
                const uint32_t zv = zvalid_mask(w, g.rz);
                const uint32_t A = a[j], B = b[j], C = c[j], D = d[j];
                const uint32_t A2 = __funnelshift_r(A, a[j + 1], 1), B2 = __funnelshift_r(B, b[j + 1], 1);
                const uint32_t C2 = __funnelshift_r(C, c[j + 1], 1), D2 = __funnelshift_r(D, d[j + 1], 1);
                uint32_t m[8];
                m[0] = (A ^ B) & hx;        // :29-33  / :100-111
                m[1] = (A ^ D) & hy;        // :35-39  / :113-124
                m[2] = (A ^ A2) & zv;       // :41-45  / :126-137 (rows past ry hold zeros)
                const uint32_t any = A | B | C | D | A2 | B2 | C2 | D2, all = A & B & C & D & A2 & B2 & C2 & D2;
                const uint32_t act = (any & ~all) & zv & hc;  // :48 / :154
                if (!EMIT) {
                    nx += __popc(m[0]);
                    ny += __popc(m[1]);
                    nz += __popc(m[2]);
                    uint32_t rem = act;
                    const uint4 w4 = make_uint4(A, B, C, D), w42 = make_uint4(A2, B2, C2, D2);
                    while (rem) {
                        const int i = __ffs(rem) - 1;
                        rem &= rem - 1;
                        nf += (uint32_t)(s_table[cube_case_at(w4, w42, i)] >> 60);
                    }
                    continue;
                }
                m[3] = (B ^ C) & hc;
                m[4] = (B ^ B2) & zv & hc;
                m[5] = (D ^ C) & hc;
                m[6] = (D ^ D2) & zv & hc;
                m[7] = (C ^ C2) & zv & hc;
                if (__any_sync(kFull, (m[0] | m[1] | m[2] | act) != 0u)) {  // else: nothing crosses in these 32x32 samples

                // ---- vertices on the rows' own +x / +y / +z edges (gen_vertices_kernel) ----
                {
                    uint32_t r0 = m[0], r1 = m[1], r2 = m[2];
                    uint32_t i0 = run[0], i1 = run[1], i2 = run[2];
                    const uint32_t zb = (uint32_t)w << 5;
                    for (;;) {
                        const uint32_t have = __popc(r0) + __popc(r1) + __popc(r2);
                        if (!__any_sync(kFull, have != 0u)) break;
                        const uint32_t take = have < 3u ? have : 3u;
                        const uint32_t incl = warp_incl_scan(take, lane);
                        const uint32_t total = __shfl_sync(kFull, incl, 31);
                        uint32_t pos = vhead + vcount + incl - take;
                        for (uint32_t t = 0; t < take; ++t) {
                            uint32_t id, code;
                            if (r0) {
                                const uint32_t i = __ffs(r0) - 1;
                                r0 &= r0 - 1;
                                id = i0++;
                                code = ((zb + i) << 7) | (lane << 2) | 0u;
                            } else if (r1) {
                                const uint32_t i = __ffs(r1) - 1;
                                r1 &= r1 - 1;
                                id = i1++;
                                code = ((zb + i) << 7) | (lane << 2) | 1u;
                            } else {
                                const uint32_t i = __ffs(r2) - 1;
                                r2 &= r2 - 1;
                                id = i2++;
                                code = ((zb + i) << 7) | (lane << 2) | 2u;
                            }
                            sc.vring[(pos++) & (kVRing - 1)] = make_uint2(id, code);
                        }
                        vcount += total;
                        __syncwarp();
                        while (vcount >= 32u) flush_vertices(32u);
                        __syncwarp();
                    }
                }

                // ---- faces, voxel-major within each row, table order inside a cell (gen_faces_kernel) ----
                if (__any_sync(kFull, act != 0u)) {
                    sc.words[lane] = make_uint4(A, B, C, D);
                    sc.words2[lane] = make_uint4(A2, B2, C2, D2);
#pragma unroll
                    for (int q = 0; q < 8; ++q) sc.rank[q][lane] = make_uint2(m[q], run[q]);
                    sc.fbase[lane] = frun;
                    sc.rowtris[lane] = 0u;
                    __syncwarp();
                    uint32_t rem = act;
                    for (;;) {
                        const uint32_t have = __popc(rem);
                        if (!__any_sync(kFull, have != 0u)) break;
                        const uint32_t take = have < 4u ? have : 4u;
                        const uint32_t incl = warp_incl_scan(take, lane);
                        const uint32_t total = __shfl_sync(kFull, incl, 31);
                        uint32_t pos = incl - take;
                        for (uint32_t t = 0; t < take; ++t) {
                            const uint32_t i = __ffs(rem) - 1;
                            rem &= rem - 1;
                            sc.cell[pos++] = (uint16_t)((lane << 5) | i);
                        }
                        __syncwarp();
                        for (uint32_t k0 = 0; k0 < total; k0 += 32) {
                            // one cell per lane: case, triangle count, its offset among the row's triangles
                            const uint32_t k = k0 + lane;
                            const bool on = k < total;
                            uint32_t nt = 0, cl = 32u + lane, bit = 0;
                            uint64_t tt = 0;
                            if (on) {
                                const uint32_t e = sc.cell[k];
                                cl = e >> 5;
                                bit = e & 31u;
                                tt = s_table[cube_case_at(sc.words[cl], sc.words2[cl], bit)];
                                nt = (uint32_t)(tt >> 60);
                            }
                            const uint32_t tincl = warp_incl_scan(nt, lane);
                            const uint32_t btot = __shfl_sync(kFull, tincl, 31);
                            const uint32_t texcl = tincl - nt;
                            const uint32_t peers = __match_any_sync(kFull, cl);   // cells of one row are contiguous
                            const int head = __ffs(peers) - 1, tail = 31 - __clz(peers);
                            const uint32_t rel = (on ? sc.rowtris[cl] : 0u) + texcl - __shfl_sync(kFull, texcl, head);
                            __syncwarp();
                            if (on && lane == tail) sc.rowtris[cl] = rel + nt;
                            uint32_t tp = texcl;
                            for (uint32_t t = 0; t < nt; ++t, tt >>= 12)
                                sc.tri[tp++] = cl | (bit << 5) | (((uint32_t)tt & 0xfffu) << 10) | ((rel + t) << 22);
                            __syncwarp();
                            // one triangle per lane: rank its three edges, 12-byte stores
                            for (uint32_t jj = lane; jj < btot; jj += 32) {
                                const uint32_t ent = sc.tri[jj];
                                const uint32_t tl = ent & 31u, i = (ent >> 5) & 31u;
                                const uint32_t lt = (1u << i) - 1u;
                                int32_t *out = faces + (sc.fbase[tl] + (ent >> 22)) * 3ull;
#pragma unroll
                                for (int cc = 0; cc < 3; ++cc) {
                                    const uint32_t nib = (ent >> (10 + 4 * cc)) & 15u;
                                    // crossings below sample z (+dz): for dz = 1 the bit at z counts too; at i = 31
                                    // the whole word counts, i.e. the id of the first crossing of the next word
                                    const uint2 rk = sc.rank[nib & 7u][tl];
                                    out[cc] = prm.vertex_id_base + (int32_t)(rk.y + __popc(rk.x & (lt | ((nib >> 3) << i))));
                                }
                            }
                            __syncwarp();
                        }
                    }
                    frun += sc.rowtris[lane];
                    __syncwarp();
                }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) run[q] += __popc(m[q]);
            }
        }
        if (EMIT) {
            __syncwarp();
            if (vcount) flush_vertices(vcount);
            __syncwarp();
        } else if (valid) {
            ws.rowv[row] = make_uint4(nx, ny, nz, nf);
        }
    }
}

// Kernel selection (measured on B200, gyroid 1024^3, profiles/r1c_*): counting is faster with the
// lane = row strip kernel (0.35 ms vs 0.49 ms), emission with the warp-per-row kernel (1.6 ms vs
// 3.9 ms: the strip variant scatters its 12-byte output stores over 32 rows and its unrolled body
// misses the instruction cache).  P3D_MC_COUNT / P3D_MC_EMIT = row | strip override for A/B runs.
static bool pick_strip(const char *var, bool dflt) {
    const char *e = getenv(var);
    if (!e || !e[0]) return dflt;
    return e[0] == 's';
}
static bool use_strip_count() {
    static const bool v = pick_strip("P3D_MC_COUNT", true);
    return v;
}
static bool use_strip_emit() {
    static const bool v = pick_strip("P3D_MC_EMIT", false);
    return v;
}

void launch_count_scan(const McGeom &g, const McWorkspace &ws, cudaStream_t s) {
    if (g.owned_rows <= 0) return;
    const int sms = sm_count();
    {
        if (use_strip_count()) {
            const int64_t strips = g.owned_x * ((g.ry + 31) / 32), want = (strips + 7) / 8, cap = (int64_t)sms * 4;
            k_strip<false><<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(nullptr, g, ws, McEmitParams{}, nullptr, nullptr);
        } else {
            const int64_t want = (g.owned_rows + 7) / 8, cap = (int64_t)sms * 8;
            k_row_count<<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(g, ws);
        }
    }
    {
        const int64_t tiles = (g.owned_rows + kScanTile - 1) / kScanTile, cap = (int64_t)sms * 4;
        k_row_scan<<<(unsigned)(tiles < cap ? tiles : cap), 256, 0, s>>>(g, ws, tiles);
    }
}

void launch_emit(const float *grid, const McGeom &g, const McWorkspace &ws, const McEmitParams &p, float *verts,
                 int32_t *faces, cudaStream_t s) {
    if (g.owned_rows <= 0) return;
    const int sms = sm_count();
    if (use_strip_emit()) {
        const int64_t strips = g.owned_x * ((g.ry + 31) / 32), want = (strips + 7) / 8, cap = (int64_t)sms * P3D_STRIP_MINBLOCKS;
        k_strip<true><<<(unsigned)(want < cap ? want : cap), 256, 0, s>>>(grid, g, ws, p, verts, faces);
        return;
    }
    const int64_t want = (g.owned_rows + kRowsPerTile - 1) / kRowsPerTile, cap = (int64_t)sms * P3D_EMIT_MINBLOCKS;
    k_emit<<<(unsigned)(want < cap ? want : cap), kRowsPerTile * 32, 0, s>>>(grid, g, ws, p, verts, faces);
}

// Multi-GPU: install the next shard's first-plane row table as this shard's halo-plane numbering.
__global__ void k_import_halo(uint4 *__restrict__ dst, const uint4 *__restrict__ src, int64_t n, uint32_t delta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint4 t = src[i];
        dst[i] = make_uint4(t.x + delta, t.y + delta, t.z + delta, t.w);
    }
}

void launch_import_halo(uint4 *halo_rows, const uint32_t *table_in, int64_t ry, uint32_t delta, cudaStream_t s) {
    k_import_halo<<<(unsigned)((ry + 255) / 256), 256, 0, s>>>(halo_rows, reinterpret_cast<const uint4 *>(table_in), ry,
                                                             delta);
}

}  // namespace p3d
