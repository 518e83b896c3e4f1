// primitive3d_b200/csrc/mc_small.cu -- marching cubes of SMALL grids (and batches of them) in ONE kernel launch.
// Replaces, for grids of up to a few million samples, the launch chain of the tiled path (memset, k_tile,
// k_round_sums, k_faces, readback): at the reference's own example sizes (sphere 128^3, bunny 66^3;
// /root/reference/examples/sphere.py:8, bunny_sdf.py:10) that chain is bound by launch and dependency latency, not
// by bandwidth.  The reference's TODO for this regime is marching_cubes.cu:255-256.  Citations below are into
// /root/reference/src/prim3d/Utility/marching_cubes.cu.
//
// One persistent kernel (a cooperative launch: every CTA resident), three phases separated by a device-wide barrier.
// The unit of work is a bit word = 32 consecutive samples of a row, and a WARP takes a group of 2..32 consecutive words:
// lane j owns word j of the group (where it sits, its masks and counts, what it needs from the workspace), and the 32
// lanes together do the sample-level work of each word in turn, lane i = sample / cell i: every load is one coalesced
// 128-byte row segment, every mask is a ballot (the first form of this kernel, a thread per word with per-thread
// loops over 128 samples and over the bits of its masks, took 39 us at bunny 66^3 on 52 CTAs; this one 22 us).
//   1  inside bits of the four rows a word's cells touch (value > thresh, :25): four loads and four ballots per word;
//      crossing masks, vertex counts (:29-45) and triangle counts (:48-66: the owner looks up the cells with mixed
//      corners) per word, packed into one 32-bit count word; CTA totals
//   2  exclusive prefix over the words (CTA totals -> per-word first vertex id / first face index, numbering restarts
//      at every grid), a thread per word
//   3  vertices of the word's own +x / +y / +z edges, a lane per sample, interpolated in the reference's fp32 order
//      (:105-109, :298); faces of the word's cells, a lane per cell, voxel-major, table order inside a cell
//      (:194-208): the id of a cube edge is the first id of its word and axis + popc(mask below the cell), the twelve
//      {mask, first id} pairs of the word wait in shared memory
// Vertex numbering: voxel-major by (row, 32-sample word), x-edge vertices of a word first, then y, then z (a free
// choice: the reference's is atomicAdd-arbitrary).  The grid is read from L2 / HBM 4x in phase 1 (it is small).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "mc_case_table.h"
#include "mc_kernels.cuh"
#include "scan_utils.cuh"

namespace p3d {

namespace {

__constant__ uint64_t c_case_table_small[256] = P3D_MC_CASE_TABLE_INIT;

// One CTA of 32 warps per SM: the device-wide barrier is an atomic counter every CTA polls, and its cost grows with
// the number of CTAs (592 CTAs of 8 warps: 44 % of the stall samples of the launch were barrier waits).
constexpr int kSmallThreads = 1024;
constexpr int kSmallWarps = kSmallThreads / 32;
constexpr int kSmallSlab = kSmallThreads;  // words a CTA scans at a time (phase 2: a thread per word)
constexpr int kSmallMinWords = 2 * kSmallWarps;  // words per CTA the launch is sized for (two per warp)

// Ownership of the cube edges, from which phase 3 builds its per-edge {mask, first id} tables:
// cube edge e (numbering of :178-192) -> which of the cell's four rows owns it (0 a = (x,y), 1 b = (x+1,y),
// 2 c = (x+1,y+1), 3 d = (x,y+1)), its axis (0 x, 1 y, 2 z) and whether it sits at sample z + 1
//   e:    0  1  2  3  4  5  6  7  8  9 10 11
//   row:  a  b  d  a  a  b  d  a  a  b  c  d
//   axis: x  y  x  y  x  y  x  y  z  z  z  z
//   up:   0  0  0  0  1  1  1  1  0  0  0  0

__device__ __forceinline__ uint32_t low_mask_small(int n) {
    return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
}

// device-wide barrier: every CTA of the launch is resident (the host sizes the launch that way); `counter` starts at
// 0 and only grows, phase k waits for k * gridDim.x arrivals
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int phase) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int seen;
        asm volatile("atom.add.release.gpu.u32 %0, [%1], 1;" : "=r"(seen) : "l"(counter) : "memory");
        const unsigned int want = phase * gridDim.x;
        ++seen;
        while (seen < want) asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    }
    __syncthreads();
}
// The barrier words {arrivals, exits} live in memory the LIBRARY owns and that is zero between launches: the last CTA
// to leave the kernel (every CTA has passed the second barrier by then) puts them back to zero, so no memset node
// precedes the launch.
__device__ __forceinline__ void grid_barrier_release(unsigned int *sync) {
    if (threadIdx.x == 0) {
        unsigned int seen;
        asm volatile("atom.add.acq_rel.gpu.u32 %0, [%1], 1;" : "=r"(seen) : "l"(sync + 1) : "memory");
        if (seen == gridDim.x - 1) {
            sync[0] = 0u;
            sync[1] = 0u;
        }
    }
}

static_assert(kSmallWarps == 32, "the second scan level of phase 2 is one warp wide");

struct WordGeom {   // where word i of the batch sits
    int g;          // grid
    int x, y, w;    // row (x, y), word of the row
    uint32_t local; // word index within its grid (< 2^31: a grid of this path has at most a few million samples)
};

// grid of word i of the batch and the word's index within it
__device__ __forceinline__ uint32_t locate_grid(const SmallBatch &b, const SmallGrid *grids, int64_t i, int &g) {
    g = 0;
    if (b.ngrids > 1) {  // binary search over the grids' first words
        int lo = 0, hi = b.ngrids - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (grids[mid].word0 <= i) lo = mid; else hi = mid - 1;
        }
        g = lo;
    }
    return (uint32_t)(i - grids[g].word0);
}

__device__ __forceinline__ WordGeom locate_word(const SmallBatch &b, const SmallGrid *grids, int64_t i, const SmallGrid *&gr) {
    WordGeom o;
    o.local = locate_grid(b, grids, i, o.g);
    gr = grids + o.g;
    const uint32_t row = o.local / (uint32_t)gr->wpr;
    o.w = (int)(o.local - row * (uint32_t)gr->wpr);
    o.x = (int)(row / (uint32_t)gr->ry);
    o.y = (int)(row - (uint32_t)o.x * (uint32_t)gr->ry);
    return o;
}

// case index of the cell at bit i: corner k in bit k, corners (0,0,0),(1,0,0),(1,1,0),(0,1,0),(0,0,1),... (:50-57)
__device__ __forceinline__ uint32_t case_index(uint32_t A, uint32_t B, uint32_t C, uint32_t D, uint32_t A2, uint32_t B2, uint32_t C2,
                                               uint32_t D2, int i) {
    return ((A >> i) & 1u) | (((B >> i) & 1u) << 1) | (((C >> i) & 1u) << 2) | (((D >> i) & 1u) << 3) | (((A2 >> i) & 1u) << 4) |
           (((B2 >> i) & 1u) << 5) | (((C2 >> i) & 1u) << 6) | (((D2 >> i) & 1u) << 7);
}

// totals_host: optional second landing place of the batch totals {V, F}, in pinned host memory the device can
// write (the single-grid call reads its counts from there after the stream wait: no copy node behind the kernel).
// group: words a warp takes at a time (1..32, chosen by the host: few on a small grid so that every resident warp
// gets words, 32 on a batch so that the per-word work is done by 32 lanes at once).
__global__ void __launch_bounds__(kSmallThreads, 1)
    k_small(const __grid_constant__ SmallBatch b, const SmallGrid *grids, SmallWorkspace ws, unsigned long long *totals_host, int group) {
    __shared__ uint64_t s_table[256];
    __shared__ int8_t s_ntri[256];
    __shared__ unsigned long long s_warp[kSmallWarps][2];
    __shared__ unsigned long long s_carry[2];
    __shared__ uint2 s_edge[kSmallWarps][12];  // phase 3: {mask, id of the mask's first crossing} per cube edge of a warp's word
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (b.ngrids == 1) grids = &b.g0;
    if (tid < 256) {
        const uint64_t t = c_case_table_small[tid];
        s_table[tid] = t;
        s_ntri[tid] = (int8_t)(t >> 60);
    }
    __syncthreads();
    const int64_t nwords = b.nwords;
    // a CTA owns a contiguous range of words: the prefix of a word is the CTAs before + the words before in the CTA
    const int64_t per_cta = (nwords + gridDim.x - 1) / gridDim.x;
    const int64_t w_first = (int64_t)blockIdx.x * per_cta;
    const int64_t w_begin = w_first < nwords ? w_first : nwords, w_end = w_begin + per_cta < nwords ? w_begin + per_cta : nwords;
    const uint32_t below = (1u << lane) - 1u;

    // ------------------------------------------------------------------ phase 1: bits, masks, counts
    // A warp takes `group` consecutive words; lane j owns word j of them (where it sits, its masks and counts), and
    // the 32 lanes together classify the samples of each word in turn.
    unsigned long long cta_v = 0, cta_f = 0;  // of the words my lane owns
    for (int64_t c0 = w_begin + (int64_t)warp * group; c0 < w_end; c0 += (int64_t)kSmallWarps * group) {
        const int n = (int)(w_end - c0 < group ? w_end - c0 : group);
        // where my word sits: row pointer, strides to rows b and d, validity flags
        const float *my_pa = nullptr;
        int my_sx = 0, my_sy = 0, my_left = 0;   // element offsets to row (x + 1, y) / (x, y + 1) (0: the row does not exist); samples from z0 on
        float my_th = 0.0f;
        uint32_t my_zv = 0, nb = 0;
        if (lane < n) {
            const SmallGrid *gr;
            const WordGeom wg = locate_word(b, grids, c0 + lane, gr);
            my_pa = gr->grid + ((int64_t)wg.x * gr->ry + wg.y) * gr->rz + 32 * wg.w;
            my_sx = wg.x + 1 < gr->rx ? gr->ry * gr->rz : 0;
            my_sy = wg.y + 1 < gr->ry ? gr->rz : 0;
            my_left = gr->rz - 32 * wg.w;
            my_th = gr->thresh;
            my_zv = low_mask_small(my_left - 1);  // samples with z + 1 < rz
            // the sample after my word's 32 in each of the four rows (bit 0 of the next word), read by the word's owner:
            // four independent loads per lane, in flight together with the rounds below.  bit 0 = row a, 1 = b, 2 = c, 3 = d
            if (my_left > 32) {
                const float na = __ldg(my_pa + 32);
                const float nbv = my_sx ? __ldg(my_pa + my_sx + 32) : my_th;
                const float nd = my_sy ? __ldg(my_pa + my_sy + 32) : my_th;
                const float nc = my_sx && my_sy ? __ldg(my_pa + my_sx + my_sy + 32) : my_th;
                nb = (na > my_th ? 1u : 0u) | (nbv > my_th ? 2u : 0u) | (nc > my_th ? 4u : 0u) | (nd > my_th ? 8u : 0u);
            }
        }
        uint32_t A = 0, B = 0, C = 0, D = 0;
        // inside bits (value > thresh, :25) of the 32 samples of a word in the four rows its cells touch: lane i reads
        // sample i of each row.  Samples outside the grid count as outside the surface; their masks are cut by the
        // validity tests below.  All loads of a round of words are independent: the latency is paid once per round.
        const int my_pack = my_sy | ((my_left > 32 ? 32 : my_left) << 24);  // sy < 2^22 (a grid has at most 4 Mi samples)
        const bool one_grid = b.ngrids == 1;
        const float th_one = grids[0].thresh;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float *pa = reinterpret_cast<const float *>(__shfl_sync(kFull, reinterpret_cast<unsigned long long>(my_pa), j));
            const int sx = __shfl_sync(kFull, my_sx, j), pack = __shfl_sync(kFull, my_pack, j);
            const int sy = pack & 0xffffff, left = pack >> 24;
            const float th = one_grid ? th_one : __shfl_sync(kFull, my_th, j);
            float va = th, vb = th, vc = th, vd = th;
            if (lane < left) {
                va = __ldg(pa + lane);
                if (sx) vb = __ldg(pa + sx + lane);
                if (sy) vd = __ldg(pa + sy + lane);
                if (sx && sy) vc = __ldg(pa + sx + sy + lane);
            }
            const uint32_t a = __ballot_sync(kFull, va > th), bb = __ballot_sync(kFull, vb > th);
            const uint32_t cc = __ballot_sync(kFull, vc > th), d = __ballot_sync(kFull, vd > th);
            if (lane == j) A = a, B = bb, C = cc, D = d;
        }
        if (lane < n) {
            const bool xin = my_sx != 0, yin = my_sy != 0;
            const uint32_t An = nb & 1u, Bn = (nb >> 1) & 1u, Cn = (nb >> 2) & 1u, Dn = (nb >> 3) & 1u;
            const uint32_t zv = my_zv;
            const uint32_t A2 = __funnelshift_r(A, An, 1), B2 = __funnelshift_r(B, Bn, 1);
            const uint32_t C2 = __funnelshift_r(C, Cn, 1), D2 = __funnelshift_r(D, Dn, 1);
            const uint32_t m0 = xin ? (A ^ B) : 0u, m1 = yin ? (A ^ D) : 0u, m2 = (A ^ A2) & zv;  // :29-45
            uint32_t nf = 0;
            if (xin && yin) {  // cells :48-66: the cells with mixed corners are looked up
                for (uint32_t rem = ((A ^ B) | (A2 ^ B2) | (D ^ C) | (D2 ^ C2) | (A ^ D) | (A2 ^ D2) | (B ^ C) | (B2 ^ C2) | (A ^ A2) |
                                     (B ^ B2) | (C ^ C2) | (D ^ D2)) & zv; rem;) {
                    const int i = __ffs(rem) - 1;
                    rem &= rem - 1;
                    nf += (uint32_t)s_ntri[case_index(A, B, C, D, A2, B2, C2, D2, i)];
                }
            }
            const uint32_t nx = __popc(m0), ny = __popc(m1), nz = __popc(m2);
            ws.corner[c0 + lane] = make_uint4(A, B, C, D);
            ws.cnt[c0 + lane] = nx | (ny << 6) | (nz << 12) | (nf << 18) | (nb << 28);
            cta_v += nx + ny + nz;
            cta_f += nf;
        }
    }
    cta_v = warp_sum64(cta_v);
    cta_f = warp_sum64(cta_f);
    if (lane == 0) s_warp[warp][0] = cta_v, s_warp[warp][1] = cta_f;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long v = warp_sum64(s_warp[lane][0]), f = warp_sum64(s_warp[lane][1]);
        if (lane == 0) {
            ws.cta_sums[2 * blockIdx.x] = v;
            ws.cta_sums[2 * blockIdx.x + 1] = f;
        }
    }
    grid_barrier(ws.sync, 1u);

    // ------------------------------------------------------------------ phase 2: prefix (a thread per word)
    {
        unsigned long long v = 0, f = 0;
        if (warp == 0) {
            const ulonglong2 *cs = reinterpret_cast<const ulonglong2 *>(ws.cta_sums);
            for (int k = lane; k < (int)blockIdx.x; k += 32) {
                const ulonglong2 t = __ldcg(cs + k);
                v += t.x, f += t.y;
            }
            v = warp_sum64(v), f = warp_sum64(f);
            if (lane == 0) s_carry[0] = v, s_carry[1] = f;
        }
        if (blockIdx.x == gridDim.x - 1 && warp == 1) {  // batch totals
            const ulonglong2 *cs = reinterpret_cast<const ulonglong2 *>(ws.cta_sums);
            unsigned long long tv = 0, tf = 0;
            for (int k = lane; k < (int)gridDim.x; k += 32) {
                const ulonglong2 t = __ldcg(cs + k);
                tv += t.x, tf += t.y;
            }
            tv = warp_sum64(tv), tf = warp_sum64(tf);
            if (lane == 0) {
                ws.header->total_v = tv, ws.header->total_f = tf;
                if (totals_host) totals_host[0] = tv, totals_host[1] = tf;
            }
        }
        __syncthreads();
    }
    for (int64_t base = w_begin; base < w_end; base += kSmallSlab) {
        const int64_t mine = base + tid;
        const uint32_t packed = mine < w_end ? ws.cnt[mine] : 0u;
        const uint32_t nx = packed & 63u, ny = (packed >> 6) & 63u, nz = (packed >> 12) & 63u, nf = (packed >> 18) & 1023u;
        // CTA-wide exclusive scan of {vertices, faces} of the slab's words, one 64-bit scan (a slab of 1024 words has up
        // to 98 k vertices and 164 k triangles)
        const unsigned long long both = (unsigned long long)(nx + ny + nz) | ((unsigned long long)nf << 32);
        const unsigned long long incl = warp_incl_scan64(both, lane);
        __syncthreads();  // s_warp is reused from the last round
        if (lane == 31) s_warp[warp][0] = incl;
        __syncthreads();
        // second level: every warp scans the 32 warp totals
        const unsigned long long wt = s_warp[lane][0], wincl = warp_incl_scan64(wt, lane);
        const unsigned long long before = __shfl_sync(kFull, wincl - wt, warp), slab_total = __shfl_sync(kFull, wincl, 31);
        const unsigned long long excl = before + incl - both;
        const unsigned long long vfirst = s_carry[0] + (excl & 0xffffffffull), ffirst = s_carry[1] + (excl >> 32);
        __syncthreads();
        if (tid == 0) s_carry[0] += slab_total & 0xffffffffull, s_carry[1] += slab_total >> 32;
        if (mine < w_end) {
            int g;
            if (locate_grid(b, grids, mine, g) == 0u) {  // numbering restarts at every grid
                ws.grid_base[2 * g] = vfirst;
                ws.grid_base[2 * g + 1] = ffirst;
            }
            ws.first[mine] = make_uint2((uint32_t)vfirst, (uint32_t)ffirst);  // batch-wide (< 2^32: the host checks the sizes)
        }
    }
    grid_barrier(ws.sync, 2u);
    grid_barrier_release(ws.sync);

    // ------------------------------------------------------------------ phase 3: vertices and faces
    // They need the grids' bases, i.e. phase 2 of every CTA: they run after the second barrier.  Again a warp takes
    // `group` consecutive words: lane j fetches everything word j needs from the workspace (one round of independent
    // loads for the whole group), then the warp emits the words that have anything to emit one after the other, lane i
    // = sample / cell i.  Data written by other CTAs is read past L1 (__ldcg).
    for (int64_t c0 = w_begin + (int64_t)warp * group; c0 < w_end; c0 += (int64_t)kSmallWarps * group) {
        const int n = (int)(w_end - c0 < group ? w_end - c0 : group);
        const int64_t mine = c0 + lane;
        uint32_t packed = 0;
        if (lane < n) packed = ws.cnt[mine];
        const uint32_t my_nf = (packed >> 18) & 1023u;
        const bool busy = (packed & 0x0fffffffu) != 0u;
        uint32_t my_A = 0, my_B = 0, my_C = 0, my_D = 0, my_vlocal = 0, my_flocal = 0;
        uint32_t my_by = 0, my_bz = 0, my_cz = 0, my_dx = 0, my_dz = 0, my_n4 = 0, my_n5 = 0, my_n6 = 0, my_n7 = 0;
        int my_x = 0, my_y = 0, my_w = 0, my_g = 0;
        bool my_faces = false;
        if (busy) {
            const SmallGrid *gr;
            const WordGeom wg = locate_word(b, grids, mine, gr);
            my_x = wg.x, my_y = wg.y, my_w = wg.w, my_g = wg.g;
            const uint4 cw = ws.corner[mine];
            const uint2 fst = __ldcg(ws.first + mine);
            const unsigned long long gv = __ldcg(ws.grid_base + 2 * wg.g), gf = __ldcg(ws.grid_base + 2 * wg.g + 1);
            my_A = cw.x, my_B = cw.y, my_C = cw.z, my_D = cw.w;
            my_vlocal = fst.x - (uint32_t)gv;  // first vertex id of my word within its grid
            my_flocal = fst.y - (uint32_t)gf;
            my_faces = my_nf != 0u && gr->faces != nullptr && (int64_t)my_flocal + my_nf <= gr->face_capacity;
            if (my_faces) {  // otherwise the caller redoes the grid with an exact buffer
                // the words of the four rows: a = mine, b = + one plane, d = + one row, c = both (they exist: the word has cells)
                const int64_t wb = mine + (int64_t)gr->ry * gr->wpr, wd = mine + gr->wpr, wc = wb + gr->wpr;
                // first vertex ids of a row's word (within the grid): x-edge vertices first, then y, then z
                auto firsts = [&](int64_t word, uint32_t &fx, uint32_t &fy, uint32_t &fz) {
                    const uint32_t pk = __ldcg(ws.cnt + word);
                    fx = __ldcg(ws.first + word).x - (uint32_t)gv, fy = fx + (pk & 63u), fz = fy + ((pk >> 6) & 63u);
                };
                uint32_t t0, t1, t2;
                firsts(wb, t0, my_by, my_bz), firsts(wc, t0, t1, my_cz), firsts(wd, my_dx, t1, my_dz);
                // the cell at bit 31 (it exists if sample z0 + 32 does): its z + 1 edges are bit 0 of the NEXT words of rows a, b, d
                if (gr->rz - 32 * wg.w > 32) {
                    firsts(mine + 1, my_n4, my_n7, t0);   // e4 = x-edge of row a, e7 = y-edge of row a
                    firsts(wb + 1, t0, my_n5, t1);        // e5 = y-edge of row b
                    firsts(wd + 1, my_n6, t1, t2);        // e6 = x-edge of row d
                }
            }
        }
        for (uint32_t todo = __ballot_sync(kFull, busy); todo;) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t pk = __shfl_sync(kFull, packed, j);
            const uint32_t nx = pk & 63u, ny = (pk >> 6) & 63u, nz = (pk >> 12) & 63u, nb = pk >> 28;
            const uint32_t A = __shfl_sync(kFull, my_A, j), B = __shfl_sync(kFull, my_B, j);
            const uint32_t C = __shfl_sync(kFull, my_C, j), D = __shfl_sync(kFull, my_D, j);
            const uint32_t vlocal = __shfl_sync(kFull, my_vlocal, j);
            const int x = __shfl_sync(kFull, my_x, j), y = __shfl_sync(kFull, my_y, j), z0 = 32 * __shfl_sync(kFull, my_w, j);
            const SmallGrid *gr = grids + (b.ngrids > 1 ? __shfl_sync(kFull, my_g, j) : 0);
            const bool do_faces = __shfl_sync(kFull, (int)my_faces, j) != 0;
            const uint32_t An = nb & 1u, Bn = (nb >> 1) & 1u, Cn = (nb >> 2) & 1u, Dn = (nb >> 3) & 1u;
            const uint32_t zv = low_mask_small(gr->rz - 1 - z0);
            const bool xin = x + 1 < gr->rx, yin = y + 1 < gr->ry;
            const uint32_t A2 = __funnelshift_r(A, An, 1), B2 = __funnelshift_r(B, Bn, 1);
            const uint32_t C2 = __funnelshift_r(C, Cn, 1), D2 = __funnelshift_r(D, Dn, 1);
            // ---- vertices of the word's own edges (:100-137), position scaled as :298: lane i owns the edges of sample i ----
            if (gr->vertices && nx + ny + nz) {
                const uint32_t m0 = xin ? (A ^ B) : 0u, m1 = yin ? (A ^ D) : 0u, m2 = (A ^ A2) & zv;
                const float *pa = gr->grid + ((int64_t)x * gr->ry + y) * gr->rz + z0;
                const bool cx = (m0 >> lane) & 1u, cy = (m1 >> lane) & 1u, cz = (m2 >> lane) & 1u;
                float d0 = 0.0f, dx = 0.0f, dy = 0.0f;
                if (z0 + lane < gr->rz) d0 = __ldg(pa + lane);
                if (cx) dx = __ldg(pa + (int64_t)gr->ry * gr->rz + lane);
                if (cy) dy = __ldg(pa + gr->rz + lane);
                float dz = __shfl_down_sync(kFull, d0, 1);
                if (lane == 31 && cz) dz = __ldg(pa + 32);
                const float th = gr->thresh;
                const float fx = (float)x, fy = (float)y, fz = (float)(z0 + lane);
                auto put = [&](uint32_t id, float px, float py, float pz) {
                    if ((int64_t)id >= gr->vertex_capacity) return;
                    float *out = gr->vertices + (int64_t)id * 3;
                    out[0] = __fadd_rn(__fmul_rn(px, gr->scale[0]), gr->offset[0]);
                    out[1] = __fadd_rn(__fmul_rn(py, gr->scale[1]), gr->offset[1]);
                    out[2] = __fadd_rn(__fmul_rn(pz, gr->scale[2]), gr->offset[2]);
                };
                // dt = (thresh - d_self) / (d_next - d_self), IEEE fp32, no contraction (:105)
                if (cx) put(vlocal + __popc(m0 & below), __fadd_rn(fx, __fdiv_rn(__fsub_rn(th, d0), __fsub_rn(dx, d0))), fy, fz);
                if (cy) put(vlocal + nx + __popc(m1 & below), fx, __fadd_rn(fy, __fdiv_rn(__fsub_rn(th, d0), __fsub_rn(dy, d0))), fz);
                if (cz) put(vlocal + nx + ny + __popc(m2 & below), fx, fy, __fadd_rn(fz, __fdiv_rn(__fsub_rn(th, d0), __fsub_rn(dz, d0))));
            }
            // ---- faces of the word's cells (:140-209): lane i owns the cell at sample i ----
            if (!do_faces) continue;
            const uint32_t flocal = __shfl_sync(kFull, my_flocal, j);
            const uint32_t ax_ = vlocal, ay_ = ax_ + nx, az_ = ay_ + ny;
            const uint32_t by_ = __shfl_sync(kFull, my_by, j), bz_ = __shfl_sync(kFull, my_bz, j), cz_ = __shfl_sync(kFull, my_cz, j);
            const uint32_t dx_ = __shfl_sync(kFull, my_dx, j), dz_ = __shfl_sync(kFull, my_dz, j);
            // per cube edge e (:178-192): crossing mask and id of the mask's first crossing, such that
            //   id(e, cell i) = first[e] + popc(mask[e] & ((1 << i) - 1));
            // the edges at sample z + 1 (e4..e7) use the mask shifted down by one bit and the id advanced by its bit 0.
            // Lane e < 12 writes pair e.
            const uint32_t xa = A ^ B, ya = A ^ D, yb = B ^ C, xd = D ^ C;
            const uint32_t za = (A ^ A2) & zv, zb = (B ^ B2) & zv, zc = (C ^ C2) & zv, zd = (D ^ D2) & zv;
            {
                const int e = lane & 3, grp = lane >> 2;               // grp 0: e0..e3, 1: e4..e7 (z + 1), 2: e8..e11 (z edges)
                const uint32_t mxy = e == 0 ? xa : (e == 1 ? yb : (e == 2 ? xd : ya));
                const uint32_t fxy = e == 0 ? ax_ : (e == 1 ? by_ : (e == 2 ? dx_ : ay_));
                const uint32_t mz = e == 0 ? za : (e == 1 ? zb : (e == 2 ? zc : zd));
                const uint32_t fz = e == 0 ? az_ : (e == 1 ? bz_ : (e == 2 ? cz_ : dz_));
                const uint2 pair = grp == 0 ? make_uint2(mxy, fxy) : (grp == 1 ? make_uint2(mxy >> 1, fxy + (mxy & 1u)) : make_uint2(mz, fz));
                if (lane < 12) s_edge[warp][lane] = pair;
            }
            const uint32_t cells = (xa | ya | za | yb | zb | xd | zd | zc) & zv;
            // lane 31 = the cell at bit 31: ids of its z + 1 edges
            const uint32_t nx4 = __shfl_sync(kFull, my_n4, j), nx5 = __shfl_sync(kFull, my_n5, j);
            const uint32_t nx6 = __shfl_sync(kFull, my_n6, j), nx7 = __shfl_sync(kFull, my_n7, j);
            uint64_t row = 0;
            if ((cells >> lane) & 1u) row = s_table[case_index(A, B, C, D, A2, B2, C2, D2, lane)];
            const uint32_t nt = (uint32_t)(row >> 60);  // <= 5
            // triangles of the cells before mine in the word: three ballots of the count's bits
            const uint32_t t0 = __ballot_sync(kFull, nt & 1u), t1 = __ballot_sync(kFull, nt & 2u), t2 = __ballot_sync(kFull, nt & 4u);
            const uint32_t before = __popc(t0 & below) + 2u * __popc(t1 & below) + 4u * __popc(t2 & below);
            __syncwarp();  // s_edge
            int32_t *out = gr->faces + ((int64_t)flocal + before) * 3;
            for (uint32_t t = 0; t < 3 * nt; ++t, row >>= 4) {
                const uint32_t e = (uint32_t)row & 15u;
                const uint2 pr = s_edge[warp][e];
                uint32_t id = pr.y + __popc(pr.x & below);
                if (lane == 31 && (e & 12u) == 4u) id = e == 4u ? nx4 : (e == 5u ? nx5 : (e == 6u ? nx6 : nx7));
                out[t] = (int32_t)id;
            }
            __syncwarp();  // the next word of this warp overwrites s_edge
        }
    }
}

}  // namespace

size_t small_workspace_bytes(int64_t nwords, int ngrids) {
    const size_t a = 256;
    auto up = [&](size_t v) { return (v + a - 1) / a * a; };
    return up(sizeof(SmallHeader)) + up((size_t)kSmallMaxCtas * 16) + up((size_t)ngrids * 16) + up((size_t)ngrids * sizeof(SmallGrid)) +
           up((size_t)(nwords + 1) * 16) + up((size_t)(nwords + 1) * 4) + up((size_t)(nwords + 1) * 8);
}

SmallWorkspace bind_small(void *base, int64_t nwords, int ngrids, SmallGrid **grids_dev) {
    const size_t a = 256;
    auto up = [&](size_t v) { return (v + a - 1) / a * a; };
    char *p = static_cast<char *>(base);
    SmallWorkspace ws;
    ws.header = reinterpret_cast<SmallHeader *>(p);      p += up(sizeof(SmallHeader));
    ws.cta_sums = reinterpret_cast<unsigned long long *>(p);  p += up((size_t)kSmallMaxCtas * 16);
    ws.grid_base = reinterpret_cast<unsigned long long *>(p); p += up((size_t)ngrids * 16);
    *grids_dev = reinterpret_cast<SmallGrid *>(p);       p += up((size_t)ngrids * sizeof(SmallGrid));
    ws.corner = reinterpret_cast<uint4 *>(p);            p += up((size_t)(nwords + 1) * 16);
    ws.cnt = reinterpret_cast<uint32_t *>(p);            p += up((size_t)(nwords + 1) * 4);
    ws.first = reinterpret_cast<uint2 *>(p);
    ws.sync = nullptr;
    return ws;
}

// The device-wide barrier's words: one allocation per host thread and device, zeroed once; every launch leaves them
// zero (grid_barrier_release).  Launches of one host thread never overlap: both callers wait for the stream.
static unsigned int *small_sync_words() {
    thread_local unsigned int *words[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    if (!words[dev]) {
        unsigned int *p = nullptr;
        if (cudaMalloc(reinterpret_cast<void **>(&p), 256) != cudaSuccess) return nullptr;
        if (cudaMemset(p, 0, 256) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
            cudaFree(p);
            return nullptr;
        }
        words[dev] = p;
    }
    return words[dev];
}

// grids_dev: the batch's descriptors in device memory (already queued on `s`)
bool launch_small(const SmallBatch &b, const SmallGrid *grids_dev, const SmallWorkspace &ws_in, cudaStream_t s,
                  unsigned long long *totals_host) {
    static const int group_env = [] {
        const char *e = getenv("P3D_SMALL_GROUP");  // tuning runs: words a warp takes at a time
        return e ? atoi(e) : 0;
    }();
    SmallWorkspace ws = ws_in;
    ws.sync = small_sync_words();
    if (!ws.sync) return false;
    static int cache[kMaxDevices];
    const int per_sm = per_device(cache, [] {
        int n = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_small, kSmallThreads, 0);
        return n > 0 ? n : 1;
    });
    // every CTA must be resident (device-wide barrier); a CTA takes at least kSmallMinWords words
    int64_t ctas = (b.nwords + kSmallMinWords - 1) / kSmallMinWords;
    const int64_t cap = (int64_t)sm_count() * per_sm < kSmallMaxCtas ? (int64_t)sm_count() * per_sm : kSmallMaxCtas;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    // words a warp takes at a time: as few as keeps every resident warp busy (latency on a small grid), 32 on a batch
    int64_t group = (b.nwords + ctas * kSmallWarps - 1) / (ctas * kSmallWarps);
    group = group < 2 ? 2 : (group > 32 ? 32 : group);
    if (group_env >= 1 && group_env <= 32) group = group_env;
    // A cooperative launch: the device-wide barrier needs every CTA resident at once, and only the driver can promise
    // that when another stream (another host thread's k_small, say) competes for the SMs -- two plain launches could
    // each hold half of the SMs and wait for the other half for ever.
    static const bool plain = [] {
        const char *e = getenv("P3D_SMALL_PLAIN_LAUNCH");  // tuning runs only
        return e && atoi(e) != 0;
    }();
    int group_i = (int)group;
    if (plain) {
        k_small<<<(unsigned)ctas, kSmallThreads, 0, s>>>(b, grids_dev, ws, totals_host, group_i);
        return true;
    }
    void *args[] = {const_cast<SmallBatch *>(&b), &grids_dev, &ws, &totals_host, &group_i};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(k_small), dim3((unsigned)ctas), dim3(kSmallThreads), args, 0, s) ==
           cudaSuccess;
}

}  // namespace p3d
