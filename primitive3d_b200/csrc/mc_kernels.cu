// primitive3d_b200/csrc/mc_kernels.cu -- see mc_kernels.cuh for the design.
// Reference citations are into /root/reference/src/prim3d/Utility/marching_cubes.cu.
#include "mc_kernels.cuh"

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
};

__device__ __forceinline__ RowPtrs row_ptrs(const uint32_t *bits, const McGeom &g, int64_t row) {
    RowPtrs r;
    const int64_t x = row / g.ry, y = row - x * g.ry;
    r.has_x = x + 1 < g.rx;
    r.has_y = y + 1 < g.ry;
    r.a = bits + row * g.wz;
    r.b = r.a + (r.has_x ? g.ry * (int64_t)g.wz : 0);
    r.d = r.a + (r.has_y ? g.wz : 0);
    r.c = r.b + (r.has_y ? g.wz : 0);
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

// 8-bit cube case of the cell at bit i (corner order of marching_cubes.cu:168-176).
__device__ __forceinline__ uint32_t cube_case(uint32_t ra, uint32_t rb, uint32_t rc, uint32_t rd) {
    // r* = two bits of a row: bit0 = sample z, bit1 = sample z+1
    return (ra & 1u) | ((rb & 1u) << 1) | ((rc & 1u) << 2) | ((rd & 1u) << 3) | ((ra >> 1) << 4) | ((rb >> 1) << 5) |
           ((rc >> 1) << 6) | ((rd >> 1) << 7);
}

// ---------------------------------------------------------------------------------------------
// K2: per-row counts + single-pass decoupled look-back scan over CTA tiles.
// Replaces count_vertices_faces_kernel (:4-68) and its two global atomic counters.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRowsPerTile * 32) k_count_scan(McGeom g, McWorkspace ws) {
    __shared__ uint8_t s_ntri[256];
    __shared__ uint32_t s_cnt[kRowsPerTile][4];
    __shared__ unsigned long long s_rowv[kRowsPerTile], s_rowf[kRowsPerTile];
    __shared__ unsigned int s_tile;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ntri[i] = (uint8_t)(c_case_table[i] >> 60);

    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&ws.header->ticket, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= g.num_tiles) break;

        const int64_t row = tile * kRowsPerTile + warp;
        uint32_t nx = 0, ny = 0, nz = 0, nf = 0;
        if (row < g.owned_rows) {
            const RowPtrs r = row_ptrs(ws.bits, g, row);
            const bool cells = r.has_x && r.has_y;
            for (int pc = 0; pc < g.pieces; ++pc) {
                const int w = pc * kPieceWords + lane;
                const Piece p = load_piece(r, w, g.wz, lane);
                const uint32_t zv = zvalid_mask(w, g.rz);
                if (r.has_x) nx += __popc(p.a ^ p.b);       // :29-33
                if (r.has_y) ny += __popc(p.a ^ p.d);       // :35-39
                nz += __popc((p.a ^ p.a2) & zv);            // :41-45
                uint32_t act = active_cells(p, zv, cells);  // :48-66
                while (act) {
                    const int i = __ffs(act) - 1;
                    act &= act - 1;
                    nf += s_ntri[cube_case((p.a >> i) & 1u | (((p.a2 >> i) & 1u) << 1),
                                           (p.b >> i) & 1u | (((p.b2 >> i) & 1u) << 1),
                                           (p.c >> i) & 1u | (((p.c2 >> i) & 1u) << 1),
                                           (p.d >> i) & 1u | (((p.d2 >> i) & 1u) << 1))];
                }
            }
            nx = warp_sum32(nx);
            ny = warp_sum32(ny);
            nz = warp_sum32(nz);
            nf = warp_sum32(nf);
        }
        if (lane == 0) {
            s_cnt[warp][0] = nx;
            s_cnt[warp][1] = ny;
            s_cnt[warp][2] = nz;
            s_cnt[warp][3] = nf;
        }
        __syncthreads();

        if (warp < 2) {  // warp 0 scans vertex counts, warp 1 face counts
            uint32_t mine = 0;
            if (lane < kRowsPerTile) mine = warp == 0 ? s_cnt[lane][0] + s_cnt[lane][1] + s_cnt[lane][2] : s_cnt[lane][3];
            const uint32_t incl = warp_incl_scan(mine, lane);
            const unsigned long long aggregate = __shfl_sync(kFull, incl, 31);
            const unsigned long long excl =
                lookback(warp == 0 ? ws.status_v : ws.status_f, tile, aggregate, lane);
            if (lane < kRowsPerTile) (warp == 0 ? s_rowv : s_rowf)[lane] = excl + incl - mine;
            if (lane == 0 && tile == g.num_tiles - 1) {
                if (warp == 0) ws.header->total_v = excl + aggregate;
                else ws.header->total_f = excl + aggregate;
            }
        }
        __syncthreads();

        if (lane == 0 && row < g.owned_rows) {
            const uint32_t vx = (uint32_t)s_rowv[warp];
            ws.rowv[row] = make_uint4(vx, vx + nx, vx + nx + ny, nf);
            ws.rowf[row] = s_rowf[warp];
        }
    }
}

void launch_count_scan(const McGeom &g, const McWorkspace &ws, cudaStream_t s) {
    const int sms = sm_count();
    if (g.num_tiles <= 0) return;
    const int64_t cap = (int64_t)sms * 8;  // persistent CTAs; tiles are handed out by ticket
    const int blocks = (int)(g.num_tiles < cap ? g.num_tiles : cap);
    k_count_scan<<<blocks, kRowsPerTile * 32, 0, s>>>(g, ws);
}

// ---------------------------------------------------------------------------------------------
// K3: emit vertices and faces.  Replaces gen_vertices_kernel (:70-138), gen_faces_kernel
// (:140-209) and the two ATen passes of the bounding-box epilogue (:298).
// ---------------------------------------------------------------------------------------------

// Mask ids q used for edge ranking: which row's crossing mask numbers the edge.
//   q0 (x,y) x-edges   q1 (x,y) y-edges   q2 (x,y) z-edges   q3 (x+1,y) y-edges
//   q4 (x+1,y) z-edges q5 (x,y+1) x-edges q6 (x,y+1) z-edges q7 (x+1,y+1) z-edges
// Cube edge e -> (q, dz) following the owner map of marching_cubes.cu:178-192:
//   e: 0 1 2 3 4 5 6 7 8 9 10 11
//   q: 0 3 5 1 0 3 5 1 2 4  7  6      dz = 1 for e in 4..7
constexpr uint64_t kEdgeToMask = (0ull << 0) | (3ull << 3) | (5ull << 6) | (1ull << 9) | (0ull << 12) | (3ull << 15) |
                                 (5ull << 18) | (1ull << 21) | (2ull << 24) | (4ull << 27) | (7ull << 30) |
                                 (6ull << 33);

struct WarpScratch {
    uint32_t words[4][kPieceWords + 1];  // a, b, c, d (+ first word of the next piece)
    uint32_t mask[8][kPieceWords];       // crossing masks q0..q7 of this piece
    uint32_t base[8][kPieceWords + 1];   // id of the first crossing at/after word w, per mask
    uint16_t list[kPieceWords * 32];     // compacted sample positions (vertex or cell work items)
};

__global__ void __launch_bounds__(kRowsPerTile * 32) k_emit(const float *__restrict__ grid, McGeom g, McWorkspace ws,
                                                           McEmitParams prm, float *__restrict__ verts,
                                                           int32_t *__restrict__ faces) {
    __shared__ uint64_t s_table[256];
    __shared__ WarpScratch s_scratch[kRowsPerTile];
    __shared__ unsigned int s_tile;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_table[i] = c_case_table[i];
    WarpScratch &sc = s_scratch[warp];
    const int64_t plane = g.ry * g.rz;

    for (;;) {
        __syncthreads();  // s_table ready / previous s_tile consumed
        if (threadIdx.x == 0) s_tile = atomicAdd(&ws.header->ticket_emit, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= g.num_tiles) break;
        const int64_t row = tile * kRowsPerTile + warp;
        if (row >= g.owned_rows) continue;

        const RowPtrs r = row_ptrs(ws.bits, g, row);
        const bool cells = r.has_x && r.has_y;
        const int64_t x = row / g.ry, y = row - x * g.ry;
        const float fx = (float)(prm.x_origin + x);  // static_cast<float>(x), :107
        const float fy = (float)y;
        const float *grow = grid + row * g.rz;

        // first ids of the four rows' vertex groups (row table written by K2 / halo import)
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

            __syncwarp();  // the previous piece's readers are done with the scratch
            sc.words[0][lane] = p.a;
            sc.words[1][lane] = p.b;
            sc.words[2][lane] = p.c;
            sc.words[3][lane] = p.d;
            if (lane == 31) {
                sc.words[0][32] = p.an;
                sc.words[1][32] = p.bn;
                sc.words[2][32] = p.cn;
                sc.words[3][32] = p.dn;
            }
            uint32_t ex[3], tot[3];  // exclusive rank / piece total of the own-row masks q0..q2
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t cnt = __popc(m[q]);
                const uint32_t incl = warp_incl_scan(cnt, lane);
                const uint32_t total = __shfl_sync(kFull, incl, 31);
                sc.mask[q][lane] = m[q];
                sc.base[q][lane] = run[q] + incl - cnt;
                if (lane == 31) sc.base[q][32] = run[q] + total;
                if (q < 3) {
                    ex[q] = incl - cnt;
                    tot[q] = total;
                }
                run[q] += total;  // now the first id of the next piece
            }

            // ---- vertices on the row's own +x / +y / +z edges (gen_vertices_kernel) ----
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                uint32_t mm = m[q];
                uint32_t pos = ex[q];
                while (mm) {
                    const int i = __ffs(mm) - 1;
                    mm &= mm - 1;
                    sc.list[pos++] = (uint16_t)((lane << 5) | i);
                }
                __syncwarp();
                const int64_t stride = q == 0 ? plane : (q == 1 ? g.rz : 1);
                const uint32_t first = run[q] - tot[q];
                for (uint32_t k = lane; k < tot[q]; k += 32) {
                    const int64_t z = (int64_t)pc * (kPieceWords * 32) + sc.list[k];
                    const float d0 = __ldg(grow + z);
                    const float d1 = __ldg(grow + z + stride);
                    // dt = (thresh - d_self) / (d_next - d_self), IEEE fp32, no contraction (:105)
                    const float dt = __fdiv_rn(__fsub_rn(prm.thresh, d0), __fsub_rn(d1, d0));
                    float px = fx, py = fy, pz = (float)z;
                    if (q == 0) px = __fadd_rn(px, dt);
                    if (q == 1) py = __fadd_rn(py, dt);
                    if (q == 2) pz = __fadd_rn(pz, dt);
                    // vertices * scale + offset as two separately rounded ops (:298)
                    float *out = verts + (int64_t)(first + k) * 3;
                    out[0] = __fadd_rn(__fmul_rn(px, prm.scale[0]), prm.offset[0]);
                    out[1] = __fadd_rn(__fmul_rn(py, prm.scale[1]), prm.offset[1]);
                    out[2] = __fadd_rn(__fmul_rn(pz, prm.scale[2]), prm.offset[2]);
                }
                __syncwarp();
            }

            // ---- faces of the row's cells, voxel-major, table order inside a cell (gen_faces_kernel) ----
            {
                const uint32_t cnt = __popc(act);
                const uint32_t incl = warp_incl_scan(cnt, lane);
                const uint32_t ncell = __shfl_sync(kFull, incl, 31);
                uint32_t mm = act, pos = incl - cnt;
                while (mm) {
                    const int i = __ffs(mm) - 1;
                    mm &= mm - 1;
                    sc.list[pos++] = (uint16_t)((lane << 5) | i);
                }
                __syncwarp();
                for (uint32_t k0 = 0; k0 < ncell; k0 += 32) {
                    const uint32_t k = k0 + lane;
                    const bool on = k < ncell;
                    const int zl = on ? sc.list[k] : 0;
                    const int wl = zl >> 5, i = zl & 31;
                    const uint32_t ra = __funnelshift_r(sc.words[0][wl], sc.words[0][wl + 1], i) & 3u;
                    const uint32_t rb = __funnelshift_r(sc.words[1][wl], sc.words[1][wl + 1], i) & 3u;
                    const uint32_t rc = __funnelshift_r(sc.words[2][wl], sc.words[2][wl + 1], i) & 3u;
                    const uint32_t rd = __funnelshift_r(sc.words[3][wl], sc.words[3][wl + 1], i) & 3u;
                    uint64_t tt = on ? s_table[cube_case(ra, rb, rc, rd)] : 0ull;
                    const uint32_t nt = (uint32_t)(tt >> 60);
                    const uint32_t tincl = warp_incl_scan(nt, lane);
                    int32_t *out = faces + (frun + (tincl - nt)) * 3ull;
                    for (uint32_t j = 0; j < 3 * nt; ++j) {
                        const uint32_t e = (uint32_t)tt & 15u;
                        tt >>= 4;
                        const uint32_t q = (uint32_t)(kEdgeToMask >> (3 * e)) & 7u;
                        const int pz = i + ((e & 12u) == 4u ? 1 : 0);  // edges 4..7 sit at z+1
                        const int w2 = wl + (pz >> 5), b2 = pz & 31;
                        const uint32_t below = sc.mask[q][w2 & 31] & ((1u << b2) - 1u);  // b2 == 0 when w2 == 32
                        out[j] = prm.vertex_id_base + (int32_t)(sc.base[q][w2] + __popc(below));
                    }
                    frun += __shfl_sync(kFull, tincl, 31);
                }
            }
        }
    }
}

void launch_emit(const float *grid, const McGeom &g, const McWorkspace &ws, const McEmitParams &p, float *verts,
                 int32_t *faces, cudaStream_t s) {
    const int sms = sm_count();
    if (g.num_tiles <= 0) return;
    const int64_t cap = (int64_t)sms * 4;
    const int blocks = (int)(g.num_tiles < cap ? g.num_tiles : cap);
    k_emit<<<blocks, kRowsPerTile * 32, 0, s>>>(grid, g, ws, p, verts, faces);
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
