// primitive3d_b200/csrc/ply_kernels.cu -- device-side assembly of the binary PLY body.
//
// The reference's save_mesh_as_ply (/root/reference/src/prim3d/Utility/marching_cubes.cu:307-352) copies
// the mesh to the host and then calls ofstream.write once per vertex (three floats + three bytes) and
// once per face (the count 3 + three ints).  Here the two record sections are built on the device --
// 15-byte vertex records {x, y, z, r, g, b} and 16-byte face records {3, a, b, c}, the same bytes --
// so the host does one copy per section and one write.
#include "../../include/prim3d_b200.h"

#include <cuda_runtime.h>

#include <string>

#include "p3d_error.h"

namespace {

// a thread per 32-bit word of the packed vertex section: coalesced stores, byte gathers from L1
__global__ void __launch_bounds__(256) k_ply_vertices(const unsigned char *__restrict__ verts, const unsigned char *__restrict__ colors,
                                                     long long nv, unsigned int *__restrict__ out, long long nwords) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    unsigned int word = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long b = 4 * w + k, rec = b / 15;
        const int off = (int)(b - rec * 15);
        unsigned int byte = 0;
        if (rec < nv) byte = off < 12 ? verts[rec * 12 + off] : colors[rec * 3 + (off - 12)];
        word |= byte << (8 * k);
    }
    out[w] = word;
}

__global__ void __launch_bounds__(256) k_ply_faces(const int *__restrict__ faces, long long nf, int4 *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nf) out[i] = make_int4(3, faces[3 * i], faces[3 * i + 1], faces[3 * i + 2]);
}

}  // namespace

extern "C" p3d_status p3d_ply_pack(const float *vertices, const uint8_t *colors, int64_t num_vertices, const int32_t *faces,
                                   int64_t num_faces, void *vertex_records, void *face_records, void *stream) {
    if (num_vertices < 0 || num_faces < 0) return p3d::set_error(P3D_ERR_INVALID, "p3d_ply_pack: negative count");
    if ((num_vertices && (!vertices || !colors || !vertex_records)) || (num_faces && (!faces || !face_records)))
        return p3d::set_error(P3D_ERR_INVALID, "p3d_ply_pack: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (num_vertices) {
        const long long nwords = (15 * (long long)num_vertices + 3) / 4;
        k_ply_vertices<<<(unsigned)((nwords + 255) / 256), 256, 0, s>>>(reinterpret_cast<const unsigned char *>(vertices), colors,
                                                                       num_vertices, static_cast<unsigned int *>(vertex_records), nwords);
    }
    if (num_faces)
        k_ply_faces<<<(unsigned)((num_faces + 255) / 256), 256, 0, s>>>(faces, num_faces, static_cast<int4 *>(face_records));
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return p3d::set_error(P3D_ERR_CUDA, std::string("p3d_ply_pack: ") + cudaGetErrorString(e));
    return P3D_OK;
}
