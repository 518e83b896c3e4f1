// primitive3d_b200/csrc/mc_faces_rows.cu -- pass B of the marching-cubes path in its row-streaming form.
// Replaces gen_faces_kernel, /root/reference/src/prim3d/Utility/marching_cubes.cu:140-209 (citations below are
// into that file).  See mc_kernels.cuh for the data the tile pass leaves behind (bit words, table entries,
// triangle counts per piece and per chunk).
//
// A warp takes a task = a run of consecutive (x, y) rows of the grid in voxel-major order and walks it row by
// row; lane l owns bit word l of the row -- of its WINDOW of the row when rows are longer than 32 words (1024
// samples): a task is then (run of rows, window), and what the other windows of a row contribute to the face
// offset comes from the per-piece triangle counts.  The
// eight crossing masks a cell's twelve edges live in,
//   q0 (x,y) x-edges   q1 (x,y) y-edges   q2 (x,y) z-edges   q3 (x+1,y) y-edges
//   q4 (x+1,y) z-edges q5 (x,y+1) x-edges q6 (x,y+1) z-edges q7 (x+1,y+1) z-edges
// split into what a row owns by itself (x-edge and z-edge masks of (x,y) and z-edge mask of (x+1,y), with the id
// of the first crossing of each word: table entry + popcounts along the piece) and what a row shares with the row
// after it (the two y-edge masks).  The row-own part of row y + 1 IS q5 q6 q7 of row y, so every mask and every
// rank is computed once per row: a row writes its own {mask, first id} pairs to the lane's slot of a shared-memory
// rank table once, as quad (t & 1) of the slot (t = the row's number in the task), and is read from there as the
// upper row of one pair and, untouched, as the lower row of the next (the case table is staged twice, the copy for
// odd rows with the two quads swapped).  Per row a lane loads two bit words and two table entries, where the chunk
// form of this pass (k_faces) loads four rows and recomputes all eight masks for every cell row.
//
// Triangles are placed without a cell list and without ballots: the tile pass left the triangle count of every
// 32-cell word (a byte per word, four to a piece), so an exclusive scan over the lanes gives each lane the position
// of its word's triangles in the row.  The lane then walks its active cells, last to first (case -> packed
// triangle row), and drops one entry per triangle -- (slot, bit, three pair indices) -- into a list in shared
// memory at that position; the list is consumed one triangle per lane: three  first id + popc(mask below z)
// ranks from the rank table, 12-byte store, voxel-major cell order and table order inside a cell (:194-208).
//
// Face offsets: the first face of a task is the sum of the rounds before its round, of the chunks before its
// chunk in the round and of the pieces before its first piece in the chunk (final data written by the tile
// pass and k_round_sums: one batch of loads); inside the task one running counter, because the rows of a task
// are contiguous in the output.
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <type_traits>

#include "mc_case_table.h"
#include "mc_kernels.cuh"
#include "scan_utils.cuh"

#ifndef P3D_ROWS_LATE_LOADS
#define P3D_ROWS_LATE_LOADS 0
#endif
#ifndef P3D_ROWS_WARPS
#define P3D_ROWS_WARPS 8
#endif
#ifndef P3D_ROWS_CTAS
#define P3D_ROWS_CTAS 4  // CTAs per SM of the one-word-per-lane instance (register cap 65536 / (256 * CTAs))
#endif

namespace p3d {

namespace {

__constant__ uint64_t c_case_table_rows[256] = P3D_MC_CASE_TABLE_INIT;

// A word slot of the rank table is twelve {mask, id of the mask's first crossing} pairs, one per cube edge, so that
//     id(edge e of the cell at bit i) = pair.y + popc(pair.x & ((1 << i) - 1))
// holds for all twelve: the edges at sample z + 1 (e4..e7) use the mask shifted down by one bit and the id
// advanced by the mask's bit 0.  Pair index of edge e (owner map :178-192), one nibble per edge:
//   e:     0   1  2  3  4   5  6  7  8  9 10 11
//   pair:  0  10  4  8  1  11  5  9  2  3  7  6
//   pairs 0..3   what row (x, y) owns:      e0 x-edge, e4 the same at z + 1, e8 z-edge, e9 z-edge of (x + 1, y)
//   pairs 4..7   what row (x, y + 1) owns:  e2, e6, e11, e10
//   pairs 8..11  what the two rows share:   e3 y-edge of (x, y), e7 at z + 1, e1 y-edge of (x + 1, y), e5 at z + 1
constexpr uint64_t kEdgePair = (0ull << 0) | (10ull << 4) | (4ull << 8) | (8ull << 12) | (1ull << 16) | (11ull << 20) |
                               (5ull << 24) | (9ull << 28) | (2ull << 32) | (3ull << 36) | (7ull << 40) | (6ull << 44);
constexpr int kSlotBytes = 112;    // 12 pairs + 16 bytes: 16-byte stores of eight consecutive slots hit 32 different banks

constexpr int kRowWarps = P3D_ROWS_WARPS;
constexpr int kRowTriCap = 256;    // triangles per window of a row's list
constexpr int kTaskPieces = 128;   // pieces per chunk of the tile pass's chunk sums

__device__ __forceinline__ uint32_t low_mask_rows(int n) {
    return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
}

// A slot holds the own quads of TWO consecutive rows (row t of the task writes quad t & 1) and the y quad of the
// pair being processed: the quad a row writes as the upper row of one pair is, untouched, the lower quad of the next
// pair.  The case table comes in two variants: for odd t the pair indices of the own quads are swapped (idx ^ 4).
template <int NW>
struct RowScratch {
    uint4 rank[32 * NW * kSlotBytes / 16];  // per word slot: quad 0, quad 1, y quad (two uint4 = four pairs each) + pad
    uint32_t tri[kRowTriCap];               // (slot << 5 | bit) << 12 | three pair indices
};

// what a lane keeps of a row per word; the row's masks and first ids wait in shared memory
template <int NW>
struct RowKeep {
    uint32_t a[NW], b[NW];        // inside bits of (x, y) and (x + 1, y)
    uint32_t an[NW], bn[NW];      // the words after them in the row (bit 0 = sample z + 32)
    uint32_t own[NW];             // OR of the row's own crossing masks
    uint32_t vya[NW], vyb[NW];    // id of the first y-edge vertex of the piece, rows (x, y) and (x + 1, y)
    uint32_t nfw[NW];             // triangles of the word's cells (0 for rows without cells)
    uint32_t skip;                // packed triangle counts of a piece between my window of this row and of the next row
};

// what a row reads from global memory
template <int NW>
struct RowLoad {
    uint32_t a[NW], b[NW], nfp[NW];
    uint32_t eax[NW], eay[NW], eaz[NW], eby[NW], ebz[NW];
    uint32_t an31, bn31;          // lane 31: the first words of the next window of the row (rows of several windows)
    uint32_t skip;
};

// corner bits a0 a1 b0 b1 c0 c1 d0 d1 of the cell at bit i (x0 = sample z, x1 = sample z + 1)
__device__ __forceinline__ uint32_t corner_code_rows(uint32_t a, uint32_t an, uint32_t b, uint32_t bn, uint32_t c, uint32_t cn,
                                                     uint32_t d, uint32_t dn, int i) {
    return (__funnelshift_r(a, an, i) & 3u) | ((__funnelshift_r(b, bn, i) & 3u) << 2) | ((__funnelshift_r(c, cn, i) & 3u) << 4) |
           ((__funnelshift_r(d, dn, i) & 3u) << 6);
}

// MULTI: rows of more than 32 bit words (several windows per row)
template <int NW, bool MULTI>
__global__ void __launch_bounds__(kRowWarps * 32, P3D_ROWS_CTAS)
    k_faces_rows(McGeom g, McWorkspace ws, int32_t vbase_arg, int32_t *__restrict__ faces, unsigned long long face_capacity,
                 int vertex_base_from_header, int rows_per_task, uint32_t ntasks, int windows) {
    uint32_t vbase = (uint32_t)vbase_arg;
    if (vertex_base_from_header) vbase += (uint32_t)ws.header->vertex_base;  // multi-GPU: computed by k_apply_exchange
    // speculative launch (p3d_mc_extract): nothing is written if the buffer is too small
    if (ws.header->total_f > face_capacity) return;
    extern __shared__ __align__(16) unsigned char rows_smem[];
    // per corner code: up to five triples of pair indices (12 bits each), bits 60..63 = #triangles; [256..511] = the
    // variant for odd rows
    uint2 *s_table = reinterpret_cast<uint2 *>(rows_smem);
    RowScratch<NW> *s_scratch = reinterpret_cast<RowScratch<NW> *>(rows_smem + 512 * sizeof(uint2));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < 512; c += blockDim.x) {
        // c & 255 is a corner code a0 a1 b0 b1 c0 c1 d0 d1; the case index has corner k in bit k (:168-176)
        const uint32_t cs = (c & 1u) | ((c >> 1 & 1u) << 4) | ((c >> 2 & 1u) << 1) | ((c >> 3 & 1u) << 5) | ((c >> 4 & 1u) << 2) |
                            ((c >> 5 & 1u) << 6) | ((c >> 6 & 1u) << 3) | ((c >> 7 & 1u) << 7);
        const uint64_t t = c_case_table_rows[cs];
        const uint32_t n = (uint32_t)(t >> 60);
        uint64_t out = (uint64_t)n << 60;
        for (uint32_t j = 0; j < 3 * n; ++j) {
            const uint32_t e = (uint32_t)(t >> (4 * j)) & 15u;
            uint32_t pr = (uint32_t)(kEdgePair >> (4 * e)) & 15u;
            if (c >= 256 && pr < 8u) pr ^= 4u;
            out |= (uint64_t)pr << (4 * j);
        }
        s_table[c] = make_uint2((uint32_t)out, (uint32_t)(out >> 32));
    }
    __syncthreads();
    RowScratch<NW> &sc = s_scratch[warp];

    const int np = g.np, W = 4 * np, rz = (int)g.rz;
    const int64_t nrows = g.owned_x * g.ry;  // rows whose cells this launch owns
    const int64_t allrows = g.rx * g.ry;     // rows present in the bit words / the table
    const int w4 = lane & 3;
    // word after (k, lane) in the row
    auto next_word = [&](const uint32_t (&v)[NW], int k, uint32_t of_lane31 = 0u) {
        const uint32_t t = __shfl_down_sync(kFull, v[k], 1);
        const uint32_t u = k + 1 < NW ? __shfl_sync(kFull, v[k + 1 < NW ? k + 1 : k], 0) : of_lane31;
        return lane == 31 ? u : t;
    };
    // exclusive prefix over the words of a piece (4 adjacent lanes) of 8-bit packed counts
    auto piece_prefix = [&](uint32_t c) {
        uint32_t inc = c;
        uint32_t t = __shfl_up_sync(kFull, inc, 1);
        if (w4 >= 1) inc += t;
        t = __shfl_up_sync(kFull, inc, 2);
        if (w4 >= 2) inc += t;
        return inc - c;
    };

    // ---- tasks by ticket (the cost of a row follows the surface); the next ticket is taken a task ahead ----
    uint32_t ticket_ahead = 0;
    if (lane == 0) ticket_ahead = atomicAdd(&ws.header->ticket_faces, 1u);
    for (;;) {
        const uint32_t task = __shfl_sync(kFull, ticket_ahead, 0);
        if (task >= ntasks) break;
        if (lane == 0) ticket_ahead = atomicAdd(&ws.header->ticket_faces, 1u);
        const int win = MULTI ? (int)(task % (uint32_t)windows) : 0;  // my 32 words of the rows
        const int64_t R0 = (int64_t)(MULTI ? task / (uint32_t)windows : task) * rows_per_task;
        const int nr = (int)(nrows - R0 < rows_per_task ? nrows - R0 : rows_per_task);
        const int64_t P0 = R0 * np + 8 * win;             // first piece of the task
        const int word0 = 32 * win;
        const bool more = MULTI && word0 + 32 < W;        // the rows go on after my window
        const int pw = np - 8 * win < 8 ? np - 8 * win : 8;  // pieces of my window
        const int nskip = MULTI ? np - pw : 0;            // pieces between my window of a row and of the next row (<= 31)
        bool valid[NW];
        uint32_t zv[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            valid[k] = word0 + 32 * k + lane < W;
            zv[k] = low_mask_rows(rz - 1 - 32 * (word0 + 32 * k + lane));  // samples with z + 1 < rz
        }

        // ---- everything a row reads from global memory, issued as one batch; the pointers walk down the task's rows
        // (stepped in place: an address register that is rewritten right after its load stalls on the load) ----
        const uint32_t *pa = ws.bits + R0 * W + word0 + lane, *pb = pa + g.ry * W;
        const uint4 *qa = ws.ptab + P0 + (lane >> 2), *qb = qa + g.ry * np;
        const uint32_t *pn = ws.nf + P0 + (lane >> 2), *ps = ws.nf + P0 + pw + lane;
        const int64_t want = (int64_t)nr + 2;
        const int rows_a = (int)(allrows - R0 < want ? allrows - R0 : want);
        const int rows_b = (int)(allrows - g.ry - R0 < want ? allrows - g.ry - R0 : want);
        int rload = 0;  // the row the pointers stand on
        auto load_row = [&](RowLoad<NW> &L) {
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                L.a[k] = L.b[k] = L.nfp[k] = 0u;
                L.eax[k] = L.eay[k] = L.eaz[k] = L.eby[k] = L.ebz[k] = 0u;
                L.an31 = L.bn31 = L.skip = 0u;
                if (more) {  // the word after lane 31's is in the next window
                    if (lane == 31 && rload < rows_a) L.an31 = __ldg(pa + 1);
                    if (lane == 31 && rload < rows_b) L.bn31 = __ldg(pb + 1);
                }
                // the pieces between my window of this row and my window of the next row of the task
                if (lane < nskip && rload + 1 < nr) L.skip = __ldg(ps);
                if (valid[k] && rload < rows_a) {
                    L.a[k] = __ldg(pa + 32 * k);
                    const uint4 e = __ldg(qa + 8 * k);
                    L.eax[k] = e.x, L.eay[k] = e.y, L.eaz[k] = e.z;
                }
                if (valid[k] && rload < nr) L.nfp[k] = __ldg(pn + 8 * k);
                if (valid[k] && rload < rows_b) {
                    L.b[k] = __ldg(pb + 32 * k);
                    const uint4 e = __ldg(qb + 8 * k);
                    L.eby[k] = e.y, L.ebz[k] = e.z;
                }
            }
            ++rload;
            pa += W, pb += W, qa += np, qb += np, pn += np, ps += np;
        };
        // a row's own part: x-edge and z-edge masks (:29-33, :41-45) with the ids of their first crossings, written to
        // quad t & 1 of the lane's slots (t = row of the task)
        auto make_own = [&](const RowLoad<NW> &L, int t, RowKeep<NW> &o) {
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                o.a[k] = L.a[k];
                o.b[k] = L.b[k];
            }
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                o.an[k] = next_word(o.a, k, L.an31);
                o.bn[k] = next_word(o.b, k, L.bn31);
                const uint32_t a2 = __funnelshift_r(o.a[k], o.an[k], 1), b2 = __funnelshift_r(o.b[k], o.bn[k], 1);
                const uint32_t xm = o.a[k] ^ o.b[k];  // ranks are used for rows with x + 1 < rx only
                const uint32_t za = (o.a[k] ^ a2) & zv[k], zb = (o.b[k] ^ b2) & zv[k];
                const uint32_t c = (uint32_t)__popc(xm) | ((uint32_t)__popc(za) << 8) | ((uint32_t)__popc(zb) << 16);
                const uint32_t ex = piece_prefix(c);
                const uint32_t rx = L.eax[k] + vbase + (ex & 255u);
                uint4 *rk = &sc.rank[(32 * k + lane) * (kSlotBytes / 16) + 2 * (t & 1)];
                rk[0] = make_uint4(xm, rx, xm >> 1, rx + (xm & 1u));
                rk[1] = make_uint4(za, L.eaz[k] + vbase + ((ex >> 8) & 255u), zb, L.ebz[k] + vbase + (ex >> 16));
                o.own[k] = xm | za | zb;
                o.vya[k] = L.eay[k] + vbase;
                o.vyb[k] = L.eby[k] + vbase;
                o.nfw[k] = (L.nfp[k] >> (8 * w4)) & 255u;  // 0 for rows without cells (x + 1 == rx or y + 1 == ry)
            }
            o.skip = L.skip;
        };

        RowLoad<NW> L;
        load_row(L);

        // ---- first face of the task: rounds before its round, chunks before its chunk, pieces before it in its chunk ----
        unsigned long long frun;
        {
            const int64_t chunk0 = P0 / kTaskPieces;
            const int off = (int)(P0 - chunk0 * kTaskPieces);
            const uint32_t round = (uint32_t)(chunk0 / kRoundTiles), in_round = (uint32_t)(chunk0 % kRoundTiles);
            unsigned long long acc = 0;
            const uint32_t *cs = ws.chunk_sum + (chunk0 - in_round);
#pragma unroll
            for (int i = 0; i < kRoundTiles / 32; ++i)
                if ((uint32_t)(lane + 32 * i) < in_round) acc += __ldg(cs + lane + 32 * i);
#pragma unroll 4
            for (uint32_t r = lane; r < round; r += 32) acc += __ldg(ws.fround_sum + r);
            if (off) {
                const uint32_t *pc = ws.nf + chunk0 * kTaskPieces;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (4 * lane + i < off) acc += __dp4a(__ldg(pc + 4 * lane + i), 0x01010101u, 0u);
            }
            frun = warp_sum64(acc);
        }

        RowKeep<NW> cur, nxt;
        make_own(L, 0, cur);
        load_row(L);

        for (int j = 0; j < nr; ++j) {
            make_own(L, j + 1, nxt);  // row j + 1 of the task
#if !P3D_ROWS_LATE_LOADS
            if (j + 1 < nr) load_row(L);  // row j + 2, in flight during this row's sparse phase
#endif

            // ---- what rows y and y + 1 share: y-edge masks (:35-39) and their ids; the active cells ----
            uint32_t act[NW];
            bool anyact = false;
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                const uint32_t ya = cur.a[k] ^ nxt.a[k], yb = cur.b[k] ^ nxt.b[k];
                const uint32_t ey = piece_prefix((uint32_t)__popc(ya) | ((uint32_t)__popc(yb) << 8));
                // mixed corners (:154,168-176) <=> one of the bottom x/y edges or of the four z edges is crossed
                act[k] = cur.nfw[k] ? ((cur.own[k] | ya | yb | nxt.own[k]) & zv[k]) : 0u;
                if (act[k]) {
                    const uint32_t rya = cur.vya[k] + (ey & 255u), ryb = cur.vyb[k] + (ey >> 8);
                    uint4 *rk = &sc.rank[(32 * k + lane) * (kSlotBytes / 16) + 4];
                    rk[0] = make_uint4(ya, rya, ya >> 1, rya + (ya & 1u));
                    rk[1] = make_uint4(yb, ryb, yb >> 1, ryb + (yb & 1u));
                }
                anyact |= act[k] != 0u;
            }
            __syncwarp();
            if (__any_sync(kFull, anyact)) {
                // triangles in voxel-major order: word by word along the row.  The tile pass counted them per word, so
                // every lane knows where the triangles of its words go without seeing the other lanes' cells.
                uint32_t tend[NW], ntri = 0;
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    const uint32_t incl = warp_incl_scan(cur.nfw[k], lane);
                    tend[k] = ntri + incl;  // one past my word's last triangle, relative to the row
                    ntri += __shfl_sync(kFull, incl, 31);
                }
                const uint2 *const table = s_table + 256 * (j & 1);
                // my cells, last to first: case -> packed triangle row -> one list entry per triangle
                auto list_cells = [&](uint32_t t0, auto windowed) {
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        uint32_t pos = tend[k] - t0;  // wraps outside the window: filtered by the range tests
                        const uint32_t s17 = (uint32_t)(32 * k + lane) << 17;
                        for (uint32_t rem = act[k]; rem;) {
                            const int i = 31 - __clz(rem);
                            rem ^= 1u << i;
                            const uint2 tt = table[corner_code_rows(cur.a[k], cur.an[k], cur.b[k], cur.bn[k], nxt.b[k], nxt.bn[k],
                                                                    nxt.a[k], nxt.an[k], i)];
                            const uint32_t nt = tt.y >> 28;
                            pos -= nt;
                            const uint32_t e12 = s17 | ((uint32_t)i << 12);
                            auto put = [&](uint32_t q, uint32_t v) {
                                if (!decltype(windowed)::value || pos + q < (uint32_t)kRowTriCap) sc.tri[pos + q] = e12 | v;
                            };
                            put(0, tt.x & 0xfffu);
                            if (nt > 1) {
                                put(1, (tt.x >> 12) & 0xfffu);
                                if (nt > 2) {
                                    put(2, __funnelshift_r(tt.x, tt.y, 24) & 0xfffu);
                                    if (nt > 3) put(3, (tt.y >> 4) & 0xfffu);
                                    if (nt > 4) put(4, (tt.y >> 16) & 0xfffu);
                                }
                            }
                        }
                    }
                };
                for (uint32_t t0 = 0; t0 < ntri; t0 += kRowTriCap) {
                    if (ntri <= (uint32_t)kRowTriCap) list_cells(0u, std::false_type());
                    else list_cells(t0, std::true_type());
                    __syncwarp();
#if P3D_ROWS_LATE_LOADS
                    if (t0 == 0 && j + 1 < nr) load_row(L);  // row j + 2, in flight during the triangle loop
#endif
                    // ---- one triangle per lane: rank its three edges, 12-byte stores (:194-208) ----
                    const uint32_t n = ntri - t0 < (uint32_t)kRowTriCap ? ntri - t0 : (uint32_t)kRowTriCap;
                    int32_t *const out0 = faces + (frun + t0) * 3ull;
                    for (uint32_t t = lane; t < n; t += 32) {
                        const uint32_t ent = sc.tri[t];
                        const uint32_t below = (1u << ((ent >> 12) & 31u)) - 1u;
                        const unsigned char *rk = reinterpret_cast<const unsigned char *>(sc.rank) + (ent >> 17) * kSlotBytes;
                        int32_t *out = out0 + t * 3u;
#pragma unroll
                        for (int cc = 0; cc < 3; ++cc) {
                            const uint32_t off = cc == 0 ? (ent << 3) & 0x78u : (ent >> (4 * cc - 3)) & 0x78u;
                            const uint2 en = *reinterpret_cast<const uint2 *>(rk + off);
                            out[cc] = (int32_t)(en.y + __popc(en.x & below));
                        }
                    }
                    __syncwarp();
                }

                // the last cell of a piece (bit 127) has its z+1 x-/y-edges in the NEXT piece, which is numbered by
                // another tile: overwrite those indices with that piece's first ids (its bit 0 is rank 0)
                bool fix = false;
#pragma unroll
                for (int k = 0; k < NW; ++k) fix |= w4 == 3 && (act[k] >> 31);
                if (__any_sync(kFull, fix)) {
                    const uint32_t flip = 4u * (uint32_t)(j & 1);
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        uint32_t xay = next_word(cur.vya, k), xby = next_word(cur.vyb, k);
                        if (w4 == 3 && (act[k] >> 31)) {
                            // the next word's first x-edge ids: pair 0 of its lower / upper quad
                            const uint4 *nk = &sc.rank[(32 * k + lane + 1) * (kSlotBytes / 16)];
                            uint32_t xax = nk[2 * (j & 1)].y, xdx = nk[2 * ((j + 1) & 1)].y;
                            if (lane == 31) {  // ... which another task handles: its table entries (rows j and j + 1 of mine)
                                const uint4 *e = ws.ptab + P0 + (int64_t)j * np + 8;
                                const uint4 ea = __ldg(e), ed = __ldg(e + np);
                                xax = ea.x + vbase, xay = ea.y + vbase, xdx = ed.x + vbase;
                                xby = __ldg(e + g.ry * np).y + vbase;
                            }
                            const uint2 tt = table[corner_code_rows(cur.a[k], cur.an[k], cur.b[k], cur.bn[k], nxt.b[k], nxt.bn[k],
                                                                    nxt.a[k], nxt.an[k], 31)];
                            const uint32_t nt = tt.y >> 28;
                            uint64_t trow = (uint64_t)tt.x | ((uint64_t)tt.y << 32);
                            int32_t *out = faces + (frun + tend[k] - nt) * 3ull;  // the word's last cell: its last triangles
                            for (uint32_t t = 0; t < 3 * nt; ++t, trow >>= 4) {
                                const uint32_t pr = (uint32_t)trow & 15u;
                                if (pr == (1u ^ flip)) out[t] = (int32_t)xax;       // e4
                                else if (pr == 9u) out[t] = (int32_t)xay;           // e7
                                else if (pr == 11u) out[t] = (int32_t)xby;          // e5
                                else if (pr == (5u ^ flip)) out[t] = (int32_t)xdx;  // e6
                            }
                        }
                    }
                }
                __syncwarp();
                frun += ntri;
            }
#if P3D_ROWS_LATE_LOADS
            else if (j + 1 < nr) load_row(L);
#endif
            // rows of several windows: the triangles of the other windows' pieces before my window of the next row
            if (nskip) frun += __reduce_add_sync(kFull, __dp4a(cur.skip, 0x01010101u, 0u));
            cur = nxt;
        }
    }
}

template <int NW, bool MULTI>
void launch_rows(const McGeom &g, const McWorkspace &ws, const McEmitParams &p, int32_t *faces, int64_t face_capacity,
                 bool vertex_base_from_header, cudaStream_t s) {
    constexpr int smem = 512 * (int)sizeof(uint2) + kRowWarps * (int)sizeof(RowScratch<NW>);
    static int cache[kMaxDevices];
    const int per_sm = per_device(cache, [] {
        cudaFuncSetAttribute(k_faces_rows<NW, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int n = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_faces_rows<NW, MULTI>, kRowWarps * 32, smem);
        return n > 0 ? n : 1;
    });
    const int64_t nrows = g.owned_x * g.ry;
    // rows per task: 32 (the first row of a task is read twice), fewer on small grids so that every resident warp
    // gets a few tasks
    int rpt = 32;
    if (const char *e = getenv("P3D_ROWS_PER_TASK")) rpt = atoi(e) > 0 ? atoi(e) : rpt;  // tuning runs
    const int64_t warps = (int64_t)sm_count() * per_sm * kRowWarps;
    const int windows_ = (4 * g.np + 31) / 32;
    while (rpt > 2 && nrows / rpt * windows_ < 4 * warps) rpt /= 2;
    const int windows = (4 * g.np + 31) / 32;  // tasks of a run of rows: one per 32 bit words of a row
    const int64_t ntasks = (nrows + rpt - 1) / rpt * windows;
    const int64_t want = (ntasks + kRowWarps - 1) / kRowWarps, cap = (int64_t)sm_count() * per_sm;
    k_faces_rows<NW, MULTI><<<(unsigned)(want < cap ? want : cap), kRowWarps * 32, smem, s>>>(
        g, ws, p.vertex_id_base, faces, (unsigned long long)face_capacity, vertex_base_from_header ? 1 : 0, rpt,
        (uint32_t)ntasks, windows);
}

}  // namespace

bool faces_rows_applicable(const McGeom &g) {
    // P3D_MC_FACES = chunks / rows forces one form of the face pass where both apply (A/B runs, tests)
    static const int mode = [] {
        const char *e = getenv("P3D_MC_FACES");
        return !e ? 1 : (e[0] == 'c' ? 0 : (e[0] == 'r' ? 2 : 1));
    }();
    const int W = 4 * g.np;
    // rows of 17 bit words or more, 32 words per task (the pieces a task skips between two rows are one per lane: <= 31)
    const bool can = W > 16 && g.np <= 39 && g.npieces > 0;
    // by default only rows of ONE window (rz <= 1024): on longer rows the chunk form, which gives a lane two words of
    // a row, measured faster (gyroid 2048^3 slab of 257 planes: 0.53 against 0.60 ms)
    return can && (mode == 2 || (mode == 1 && W <= 32));
}

void launch_faces_rows(const McGeom &g, const McWorkspace &ws, const McEmitParams &p, int32_t *faces, int64_t face_capacity,
                       bool vertex_base_from_header, cudaStream_t s) {
    if (4 * g.np <= 32) launch_rows<1, false>(g, ws, p, faces, face_capacity, vertex_base_from_header, s);
    else launch_rows<1, true>(g, ws, p, faces, face_capacity, vertex_base_from_header, s);
}

}  // namespace p3d
