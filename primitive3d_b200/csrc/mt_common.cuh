// primitive3d_b200/csrc/mt_common.cuh -- tables and per-element arithmetic shared by the staged
// marching-tetrahedra kernels (mt_kernels.cu) and the one-call path (mt_extract.cu), so that both produce
// the same bits.  Reference: prim3d/utility/marching_tetrahedras.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p3d {
namespace {

// marching_tetrahedras.py:29-32 num_triangles_table, 2 bits per code
constexpr uint32_t pack_num_tri() {
    const int nt[16] = {0, 1, 1, 2, 1, 2, 2, 1, 1, 2, 2, 1, 2, 1, 1, 0};
    uint32_t w = 0;
    for (int i = 0; i < 16; ++i) w |= (uint32_t)nt[i] << (2 * i);
    return w;
}
constexpr uint32_t kNumTri = pack_num_tri();
__host__ __device__ __forceinline__ uint32_t num_tri(uint32_t code) { return (kNumTri >> (2 * code)) & 3u; }

// marching_tetrahedras.py:7-27 triangle_table: six local edge ids per code, one nibble each.
__constant__ uint32_t c_tri_rows[16] = {
    0xffffff, 0xfff201, 0xfff304, 0x431241, 0xfff513, 0x352032, 0x451041, 0xfff524,
    0xfff254, 0x154014, 0x253023, 0xfff531, 0x134214, 0xfff403, 0xfff102, 0xffffff};
// marching_tetrahedras.py:33-43 base_tet_edges: local edge -> (corner a, corner b), 2 bits each
//   e: 0:(0,1) 1:(0,2) 2:(0,3) 3:(1,2) 4:(1,3) 5:(2,3)
constexpr uint32_t kEdgeA = (0u << 0) | (0u << 2) | (0u << 4) | (1u << 6) | (1u << 8) | (2u << 10);
constexpr uint32_t kEdgeB = (1u << 0) | (2u << 2) | (3u << 4) | (2u << 6) | (3u << 8) | (3u << 10);

__device__ __forceinline__ uint64_t make_key(int64_t a, int64_t b) {
    const uint64_t lo = (uint64_t)(a < b ? a : b), hi = (uint64_t)(a < b ? b : a);
    return (lo << 32) | hi;
}

// det([1|p0; 1|p1; 1|p2; 1|p3]) = a . (b x c) with a, b, c = p1-p0, p2-p0, p3-p0 (marching_tetrahedras.py:50-65),
// in float64 with every operation spelled out so that no two kernels contract it differently.  A deliberate
// deviation from the reference's float32 torch.det, whose LU sign is backend dependent on numerically degenerate
// tets: the two agree wherever |det| is above rounding noise (oracle/mt.py::orientation_margin).
__device__ __forceinline__ double orient_det_f64(const float *__restrict__ pts, uint64_t i0, uint64_t i1, uint64_t i2, uint64_t i3) {
    const double p0x = __ldg(pts + 3 * i0), p0y = __ldg(pts + 3 * i0 + 1), p0z = __ldg(pts + 3 * i0 + 2);
    const double ax = __dsub_rn(__ldg(pts + 3 * i1), p0x), ay = __dsub_rn(__ldg(pts + 3 * i1 + 1), p0y), az = __dsub_rn(__ldg(pts + 3 * i1 + 2), p0z);
    const double bx = __dsub_rn(__ldg(pts + 3 * i2), p0x), by = __dsub_rn(__ldg(pts + 3 * i2 + 1), p0y), bz = __dsub_rn(__ldg(pts + 3 * i2 + 2), p0z);
    const double cx = __dsub_rn(__ldg(pts + 3 * i3), p0x), cy = __dsub_rn(__ldg(pts + 3 * i3 + 1), p0y), cz = __dsub_rn(__ldg(pts + 3 * i3 + 2), p0z);
    const double t1 = __fma_rn(by, cz, -__dmul_rn(bz, cy));
    const double t2 = __fma_rn(bz, cx, -__dmul_rn(bx, cz));
    const double t3 = __fma_rn(bx, cy, -__dmul_rn(by, cx));
    return __fma_rn(ax, t1, __fma_rn(ay, t2, __dmul_rn(az, t3)));
}

// The rare slow path of the float32-filtered orientation test: out of line, so that its registers are not the kernel's.
static __device__ __noinline__ bool orient_negative_f64(const float *__restrict__ pts, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t i3) {
    return orient_det_f64(pts, i0, i1, i2, i3) < 0.0;
}

// marching_tetrahedras.py:177-189, every op separately rounded: vertex v of the unique edge `key`
__device__ __forceinline__ void emit_vertex(const float *__restrict__ pts, const float *__restrict__ sdf, uint64_t key, int64_t v,
                                            float *__restrict__ verts, int64_t *__restrict__ edges) {
    const int64_t a = (int64_t)(key >> 32), b = (int64_t)(key & 0xffffffffull);
    const float s0 = __ldg(sdf + a);
    const float s1n = __fmul_rn(__ldg(sdf + b), -1.0f);  // edges_to_interp_sdf[:, -1] *= -1
    const float den = __fadd_rn(s0, s1n);                 // .sum(1)
    const float w0 = __fdiv_rn(s1n, den);                 // flip(...) / denominator
    const float w1 = __fdiv_rn(s0, den);
#pragma unroll
    for (int c = 0; c < 3; ++c)
        verts[3 * v + c] = __fadd_rn(__fmul_rn(__ldg(pts + 3 * a + c), w0), __fmul_rn(__ldg(pts + 3 * b + c), w1));
    if (edges) {
        edges[2 * v] = a;
        edges[2 * v + 1] = b;
    }
}

inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

inline int id_bits(int64_t num_points) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < num_points) ++b;
    return b;
}

}  // namespace
}  // namespace p3d
