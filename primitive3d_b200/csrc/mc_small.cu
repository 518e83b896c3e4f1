// primitive3d_b200/csrc/mc_small.cu -- marching cubes of SMALL grids (and batches of them) in ONE kernel launch.
// Replaces, for grids of up to a few million samples, the launch chain of the tiled path (memset, k_tile,
// k_round_sums, k_faces, readback): at the reference's own example sizes (sphere 128^3, bunny 66^3;
// /root/reference/examples/sphere.py:8, bunny_sdf.py:10) that chain is bound by launch and dependency latency, not
// by bandwidth.  The reference's TODO for this regime is marching_cubes.cu:255-256.  Citations below are into
// /root/reference/src/prim3d/Utility/marching_cubes.cu.
//
// One persistent kernel, every CTA resident, three phases separated by a device-wide barrier:
//   1  a thread per bit word (32 samples of a row) of the batch: inside bits of the four rows a word's cells touch
//      (value > thresh, :25), read in rounds of 32 independent predicated loads; crossing masks, vertex counts
//      (:29-45) and triangle counts (:48-66) per word, packed into one 32-bit count word; CTA totals
//   2  exclusive prefix over the words (CTA totals -> per-word first vertex id / first face index, numbering restarts
//      at every grid); vertices of the word's own +x / +y / +z edges, interpolated in the reference's fp32 order
//      (:105-109, :298)
//   3  faces of the word's cells, voxel-major, table order inside a cell (:194-208); the id of a cube edge is the
//      first id of its word and axis + popc(mask below the cell)
// Vertex numbering: voxel-major by (row, 32-sample word), x-edge vertices of a word first, then y, then z (a free
// choice: the reference's is atomicAdd-arbitrary).  The grid is read from L2 / HBM 4x in phase 1 (it is small).
#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_case_table.h"
#include "mc_kernels.cuh"
#include "scan_utils.cuh"

namespace p3d {

namespace {

__constant__ uint64_t c_case_table_small[256] = P3D_MC_CASE_TABLE_INIT;

constexpr int kSmallThreads = 256;
constexpr int kSmallSlab = kSmallThreads;  // words a CTA scans at a time

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
        __threadfence();
        atomicAdd(counter, 1u);
        const unsigned int want = phase * gridDim.x;
        while (*reinterpret_cast<volatile unsigned int *>(counter) < want) {}
        __threadfence();
    }
    __syncthreads();
}

struct WordGeom {   // where word i of the batch sits
    int g;          // grid
    int x, y, w;    // row (x, y), word of the row
    int64_t local;  // word index within its grid
};

__device__ __forceinline__ WordGeom locate_word(const SmallBatch &b, const SmallGrid *grids, int64_t i, const SmallGrid *&gr) {
    WordGeom o;
    int g = 0;
    if (b.ngrids > 1) {  // binary search over the grids' first words
        int lo = 0, hi = b.ngrids - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (grids[mid].word0 <= i) lo = mid; else hi = mid - 1;
        }
        g = lo;
    }
    gr = grids + g;
    o.g = g;
    o.local = i - gr->word0;
    const int64_t row = o.local / gr->wpr;
    o.w = (int)(o.local - row * gr->wpr);
    o.x = (int)(row / gr->ry);
    o.y = (int)(row - (int64_t)o.x * gr->ry);
    return o;
}

__global__ void __launch_bounds__(kSmallThreads, 4) k_small(const SmallBatch b, const SmallGrid *grids, SmallWorkspace ws) {
    __shared__ uint64_t s_table[256];
    __shared__ int8_t s_ntri[256];
    __shared__ unsigned long long s_warp[kSmallThreads / 32][2];
    __shared__ unsigned long long s_carry[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (b.ngrids == 1) grids = &b.g0;
    {
        const uint64_t t = c_case_table_small[tid];
        s_table[tid] = t;
        s_ntri[tid] = (int8_t)(t >> 60);
    }
    __syncthreads();
    const int64_t nwords = b.nwords;
    // a CTA owns a contiguous range of words: the prefix of a word is the CTAs before + the words before in the CTA
    const int64_t per_cta = ((nwords + gridDim.x - 1) / gridDim.x + kSmallSlab - 1) / kSmallSlab * kSmallSlab;
    const int64_t w_begin = (int64_t)blockIdx.x * per_cta, w_end = w_begin + per_cta < nwords ? w_begin + per_cta : nwords;

    // ------------------------------------------------------------------ phase 1: bits, masks, counts
    unsigned long long cta_v = 0, cta_f = 0;
    for (int64_t mine = w_begin + tid; mine < w_end; mine += kSmallThreads) {
        uint32_t packed = 0;
        {
            const SmallGrid *gr;
            const WordGeom wg = locate_word(b, grids, mine, gr);
            // inside bits (value > thresh, :25) of my 32 samples in the four rows my cells touch, and of the sample
            // after them (bit 0 of the next word); samples outside the grid count as outside the surface, their masks
            // are cut by the validity tests below.  All loads are independent: the latency is paid once.
            const int z0 = 32 * wg.w, nsamp = gr->rz - z0 < 33 ? gr->rz - z0 : 33;
            const bool xin = wg.x + 1 < gr->rx, yin = wg.y + 1 < gr->ry;
            const float *pa = gr->grid + ((int64_t)wg.x * gr->ry + wg.y) * gr->rz + z0;
            const float *pb = pa + (int64_t)gr->ry * gr->rz, *pd = pa + gr->rz, *pc = pb + gr->rz;
            const float th = gr->thresh;
            uint32_t A = 0, B = 0, C = 0, D = 0, nb = 0;
            // 8 samples of each row per round: 32 independent (predicated) loads in flight
            const bool xy = xin && yin;
#pragma unroll
            for (int i0 = 0; i0 < 32; i0 += 8) {
                float va[8], vb[8], vc[8], vd[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const bool in = i0 + k < nsamp;
                    va[k] = in ? __ldg(pa + i0 + k) : th;
                    vb[k] = in && xin ? __ldg(pb + i0 + k) : th;
                    vd[k] = in && yin ? __ldg(pd + i0 + k) : th;
                    vc[k] = in && xy ? __ldg(pc + i0 + k) : th;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    A |= (va[k] > th ? 1u : 0u) << (i0 + k);
                    B |= (vb[k] > th ? 1u : 0u) << (i0 + k);
                    C |= (vc[k] > th ? 1u : 0u) << (i0 + k);
                    D |= (vd[k] > th ? 1u : 0u) << (i0 + k);
                }
            }
            if (nsamp == 33) {
                nb |= __ldg(pa + 32) > th ? 1u : 0u;
                if (xin) nb |= __ldg(pb + 32) > th ? 2u : 0u;
                if (xin && yin) nb |= __ldg(pc + 32) > th ? 4u : 0u;
                if (yin) nb |= __ldg(pd + 32) > th ? 8u : 0u;
            }
            const uint32_t An = nb & 1u, Bn = (nb >> 1) & 1u, Cn = (nb >> 2) & 1u, Dn = (nb >> 3) & 1u;
            const uint32_t zv = low_mask_small(gr->rz - 1 - 32 * wg.w);  // samples with z + 1 < rz
            const uint32_t A2 = __funnelshift_r(A, An, 1), B2 = __funnelshift_r(B, Bn, 1);
            const uint32_t C2 = __funnelshift_r(C, Cn, 1), D2 = __funnelshift_r(D, Dn, 1);
            const uint32_t m0 = xin ? (A ^ B) : 0u, m1 = yin ? (A ^ D) : 0u, m2 = (A ^ A2) & zv;  // :29-45
            uint32_t nf = 0;
            if (xin && yin) {  // cells :48-66: E - 2 per loop, table lookup for the cells whose corners fall apart
                const uint32_t xa0 = (A ^ B) & zv, xa1 = (A2 ^ B2) & zv, xd0 = (D ^ C) & zv, xd1 = (D2 ^ C2) & zv;
                const uint32_t ya0 = (A ^ D) & zv, ya1 = (A2 ^ D2) & zv, yb0 = (B ^ C) & zv, yb1 = (B2 ^ C2) & zv;
                const uint32_t za = (A ^ A2) & zv, zb = (B ^ B2) & zv, zc = (C ^ C2) & zv, zd = (D ^ D2) & zv;
                for (uint32_t rem = xa0 | xa1 | xd0 | xd1 | ya0 | ya1 | yb0 | yb1 | za | zb | zc | zd; rem;) {
                    const int i = __ffs(rem) - 1;
                    rem &= rem - 1;
                    // case index: corner k in bit k, corners (0,0,0),(1,0,0),(1,1,0),(0,1,0),(0,0,1),... (:50-57)
                    const uint32_t cs = ((A >> i) & 1u) | (((B >> i) & 1u) << 1) | (((C >> i) & 1u) << 2) | (((D >> i) & 1u) << 3) |
                                        (((A2 >> i) & 1u) << 4) | (((B2 >> i) & 1u) << 5) | (((C2 >> i) & 1u) << 6) | (((D2 >> i) & 1u) << 7);
                    nf += (uint32_t)s_ntri[cs];
                }
            }
            const uint32_t nx = __popc(m0), ny = __popc(m1), nz = __popc(m2);
            packed = nx | (ny << 6) | (nz << 12) | (nf << 18);
            ws.corner[mine] = make_uint4(A, B, C, D);
            ws.cnt[mine] = packed | (nb << 28);
            cta_v += nx + ny + nz;
            cta_f += nf;
        }
    }
    cta_v = warp_sum64(cta_v);
    cta_f = warp_sum64(cta_f);
    if (lane == 0) s_warp[warp][0] = cta_v, s_warp[warp][1] = cta_f;
    __syncthreads();
    if (tid == 0) {
        unsigned long long v = 0, f = 0;
        for (int k = 0; k < kSmallThreads / 32; ++k) v += s_warp[k][0], f += s_warp[k][1];
        ws.cta_sums[2 * blockIdx.x] = v;
        ws.cta_sums[2 * blockIdx.x + 1] = f;
    }
    grid_barrier(&ws.header->barrier, 1u);

    // ------------------------------------------------------------------ phase 2: prefix, vertices
    {
        unsigned long long v = 0, f = 0;
        if (warp == 0) {
            for (int k = lane; k < (int)blockIdx.x; k += 32) v += ws.cta_sums[2 * k], f += ws.cta_sums[2 * k + 1];
            v = warp_sum64(v), f = warp_sum64(f);
            if (lane == 0) s_carry[0] = v, s_carry[1] = f;
        }
        if (blockIdx.x == gridDim.x - 1 && warp == 1) {  // batch totals
            unsigned long long tv = 0, tf = 0;
            for (int k = lane; k < (int)gridDim.x; k += 32) tv += ws.cta_sums[2 * k], tf += ws.cta_sums[2 * k + 1];
            tv = warp_sum64(tv), tf = warp_sum64(tf);
            if (lane == 0) ws.header->total_v = tv, ws.header->total_f = tf;
        }
        __syncthreads();
    }
    for (int64_t base = w_begin; base < w_end; base += kSmallSlab) {
        const int64_t mine = base + tid;
        const uint32_t packed = mine < w_end ? ws.cnt[mine] : 0u;
        const uint32_t nx = packed & 63u, ny = (packed >> 6) & 63u, nz = (packed >> 12) & 63u, nf = (packed >> 18) & 1023u;
        // CTA-wide exclusive scan of {vertices, faces} of the slab's words (both < 2^16 per slab: one 32-bit scan)
        const uint32_t both = (nx + ny + nz) | (nf << 16);
        const uint32_t incl = warp_incl_scan(both, lane);
        __syncthreads();  // s_warp is reused from the last round
        if (lane == 31) s_warp[warp][0] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (int k = 0; k < warp; ++k) before += (uint32_t)s_warp[k][0];
        uint32_t slab_total = 0;
        for (int k = 0; k < kSmallThreads / 32; ++k) slab_total += (uint32_t)s_warp[k][0];
        const uint32_t excl = before + incl - both;
        const unsigned long long vfirst = s_carry[0] + (excl & 0xffffu), ffirst = s_carry[1] + (excl >> 16);
        __syncthreads();
        if (tid == 0) s_carry[0] += slab_total & 0xffffu, s_carry[1] += slab_total >> 16;
        if (mine < w_end) {
            const SmallGrid *gr;
            const WordGeom wg = locate_word(b, grids, mine, gr);
            if (wg.local == 0) {  // numbering restarts at every grid
                ws.grid_base[2 * wg.g] = vfirst;
                ws.grid_base[2 * wg.g + 1] = ffirst;
            }
            ws.first[mine] = make_uint2((uint32_t)vfirst, (uint32_t)ffirst);  // batch-wide (< 2^32: the host checks the sizes)
        }
    }
    grid_barrier(&ws.header->barrier, 2u);

    // vertices and faces need the grids' bases, i.e. phase 2 of every CTA: they run after the second barrier
    for (int64_t mine = w_begin + tid; mine < w_end; mine += kSmallThreads) {
        const uint32_t packed = ws.cnt[mine];
        const uint32_t nx = packed & 63u, ny = (packed >> 6) & 63u, nz = (packed >> 12) & 63u, nf = (packed >> 18) & 1023u;
        if (nx + ny + nz + nf == 0) continue;
        const SmallGrid *gr;
        const WordGeom wg = locate_word(b, grids, mine, gr);
        const uint4 cw = ws.corner[mine];
        const uint32_t A = cw.x, B = cw.y, C = cw.z, D = cw.w, nb = packed >> 28;
        const uint32_t An = nb & 1u, Bn = (nb >> 1) & 1u, Cn = (nb >> 2) & 1u, Dn = (nb >> 3) & 1u;
        const uint32_t zv = low_mask_small(gr->rz - 1 - 32 * wg.w);
        const bool xin = wg.x + 1 < gr->rx, yin = wg.y + 1 < gr->ry;
        const uint32_t A2 = __funnelshift_r(A, An, 1), B2 = __funnelshift_r(B, Bn, 1);
        const uint32_t C2 = __funnelshift_r(C, Cn, 1), D2 = __funnelshift_r(D, Dn, 1);
        const uint2 fst = ws.first[mine];
        const unsigned long long gv = ws.grid_base[2 * wg.g], gf = ws.grid_base[2 * wg.g + 1];
        const uint32_t vlocal = fst.x - (uint32_t)gv;  // first vertex id of my word within its grid
        // ---- vertices of my word's own edges (:100-137), position scaled as :298 ----
        if (gr->vertices) {
            const uint32_t masks[3] = {xin ? (A ^ B) : 0u, yin ? (A ^ D) : 0u, (A ^ A2) & zv};
            const int64_t stride[3] = {(int64_t)gr->ry * gr->rz, gr->rz, 1};
            const float *p0 = gr->grid + ((int64_t)wg.x * gr->ry + wg.y) * gr->rz + 32 * wg.w;
            uint32_t id = vlocal;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax)
                for (uint32_t rem = masks[ax]; rem; ++id) {
                    const int i = __ffs(rem) - 1;
                    rem &= rem - 1;
                    if ((int64_t)id >= gr->vertex_capacity) continue;
                    const float d0 = __ldg(p0 + i), d1 = __ldg(p0 + i + stride[ax]);
                    const float dt = __fdiv_rn(__fsub_rn(gr->thresh, d0), __fsub_rn(d1, d0));
                    float px = (float)(wg.x), py = (float)wg.y, pz = (float)(32 * wg.w + i);
                    if (ax == 0) px = __fadd_rn(px, dt);
                    if (ax == 1) py = __fadd_rn(py, dt);
                    if (ax == 2) pz = __fadd_rn(pz, dt);
                    float *out = gr->vertices + (int64_t)id * 3;
                    out[0] = __fadd_rn(__fmul_rn(px, gr->scale[0]), gr->offset[0]);
                    out[1] = __fadd_rn(__fmul_rn(py, gr->scale[1]), gr->offset[1]);
                    out[2] = __fadd_rn(__fmul_rn(pz, gr->scale[2]), gr->offset[2]);
                }
        }
        // ---- faces of my word's cells (:140-209) ----
        if (nf == 0 || !gr->faces) continue;
        const uint32_t flocal = fst.y - (uint32_t)gf;
        if ((int64_t)flocal + nf > gr->face_capacity) continue;  // the caller redoes the grid with an exact buffer
        // the words of the four rows: a = mine, b = + one plane, d = + one row, c = both
        const int64_t wb = mine + (int64_t)gr->ry * gr->wpr, wd = mine + gr->wpr, wc = wb + gr->wpr;
        // first vertex ids of a row's word (within the grid): x-edge vertices first, then y, then z
        auto firsts = [&](int64_t word, uint32_t &fx, uint32_t &fy, uint32_t &fz) {
            const uint32_t pk = ws.cnt[word];
            fx = ws.first[word].x - (uint32_t)gv, fy = fx + (pk & 63u), fz = fy + ((pk >> 6) & 63u);
        };
        uint32_t ax_, ay_, az_, bx_, by_, bz_, cx_, cy_, cz_, dx_, dy_, dz_;
        firsts(mine, ax_, ay_, az_), firsts(wb, bx_, by_, bz_), firsts(wc, cx_, cy_, cz_), firsts(wd, dx_, dy_, dz_);
        // per cube edge e (:178-192): crossing mask and id of the mask's first crossing, such that
        //   id(e, cell i) = ef[e] + popc(em[e] & ((1 << i) - 1));
        // the edges at sample z + 1 (e4..e7) use the mask shifted down by one bit and the id advanced by its bit 0
        const uint32_t xa = A ^ B, ya = A ^ D, yb = B ^ C, xd = D ^ C;
        const uint32_t em[12] = {xa, yb, xd, ya, xa >> 1, yb >> 1, xd >> 1, ya >> 1,
                                 (A ^ A2) & zv, (B ^ B2) & zv, (C ^ C2) & zv, (D ^ D2) & zv};
        const uint32_t ef[12] = {ax_, by_, dx_, ay_, ax_ + (xa & 1u), by_ + (yb & 1u), dx_ + (xd & 1u), ay_ + (ya & 1u),
                                 az_, bz_, cz_, dz_};
        (void)bx_, (void)cx_, (void)cy_, (void)dy_;
        const uint32_t cells = (xa | ya | em[8] | yb | em[9] | xd | em[11] | em[10]) & zv;
        // the cell at bit 31: its z + 1 edges are bit 0 of the NEXT words of rows a, b, d
        uint32_t nx4 = 0, nx5 = 0, nx6 = 0, nx7 = 0;
        if (cells >> 31) {
            uint32_t t0, t1, t2;
            firsts(mine + 1, nx4, nx7, t0);   // e4 = x-edge of row a, e7 = y-edge of row a
            firsts(wb + 1, t0, nx5, t1);      // e5 = y-edge of row b
            firsts(wd + 1, nx6, t1, t2);      // e6 = x-edge of row d
        }
        int32_t *out = gr->faces + (int64_t)flocal * 3;
        for (uint32_t rem = cells; rem;) {
            const int i = __ffs(rem) - 1;
            rem &= rem - 1;
            const uint32_t cs = ((A >> i) & 1u) | (((B >> i) & 1u) << 1) | (((C >> i) & 1u) << 2) | (((D >> i) & 1u) << 3) |
                                (((A2 >> i) & 1u) << 4) | (((B2 >> i) & 1u) << 5) | (((C2 >> i) & 1u) << 6) | (((D2 >> i) & 1u) << 7);
            uint64_t row = s_table[cs];
            const uint32_t nt = (uint32_t)(row >> 60), below = (1u << i) - 1u;
            for (uint32_t t = 0; t < 3 * nt; ++t, row >>= 4) {
                const uint32_t e = (uint32_t)row & 15u;
                uint32_t id = ef[e] + __popc(em[e] & below);
                if (i == 31 && (e & 12u) == 4u) id = e == 4u ? nx4 : (e == 5u ? nx5 : (e == 6u ? nx6 : nx7));
                out[t] = (int32_t)id;
            }
            out += 3 * nt;
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
    return ws;
}

// grids_dev: the batch's descriptors in device memory (already queued on `s`)
void launch_small(const SmallBatch &b, const SmallGrid *grids_dev, const SmallWorkspace &ws, cudaStream_t s) {
    static int cache[kMaxDevices];
    const int per_sm = per_device(cache, [] {
        int n = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_small, kSmallThreads, 0);
        return n > 0 ? n : 1;
    });
    // every CTA must be resident (device-wide barrier); a CTA takes at least one slab of 256 words
    int64_t ctas = (b.nwords + kSmallSlab - 1) / kSmallSlab;
    const int64_t cap = (int64_t)sm_count() * per_sm < kSmallMaxCtas ? (int64_t)sm_count() * per_sm : kSmallMaxCtas;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    k_small<<<(unsigned)ctas, kSmallThreads, 0, s>>>(b, grids_dev, ws);
}

}  // namespace p3d
