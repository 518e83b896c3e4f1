// primitive3d_b200/csrc/mc_peer.cu -- the shard-boundary exchange of the multi-GPU path over PEER MEMORY (NVLink /
// NVSwitch stores into the other ranks' device memory), in place of the NCCL all-gather of p3d_mc_sharded_extract.
//
// What crosses a shard boundary (primitive3d_b200/sharded.py; the reference has no multi-GPU
// path): rank r needs the first-plane table of rank r + 1 (the numbering of the vertices its last-plane cells refer to)
// and the vertex counts of the ranks below it (its vertex id base); the host wants every rank's {V, F}.  The
// all-gather moved every rank's table to every rank (7/8 of the payload unused at 8 ranks) behind a collective launch.
// Here every rank owns a mailbox in device memory that the other ranks map through CUDA IPC:
//   control   flags[2][32]   flags[p][t] = epoch of the last call of parity p whose payload rank t has delivered
//             done           arrival counter of k_export_p2p's CTAs (back to zero when the kernel ends)
//             timeout        set by k_wait_p2p if a peer never arrived (the host turns it into an error)
//   recv      [2][world][n + 1] uint4: the all-gather layout of p3d_mc_faces_exchanged, one copy per call parity
// and the exchange of one call is two small kernels on the caller's stream:
//   k_export_p2p  stores this rank's first-plane table into rank r - 1's recv[p][r] and its {V, F} into EVERY rank's
//                 recv[p][r][n]; system-scope fence; the last CTA to finish raises flags[p][r] = epoch on every rank
//   k_wait_p2p    one warp waits until flags[p][t] == epoch for every t (bounded spin)
// after which k_apply_exchange / the face pass read recv[p] exactly as they read the all-gather's output.  No rank
// waits for anything but the export kernels of the other ranks, which wait for nobody: no cycle.  Two parities make
// the mailbox safe against a fast neighbour: to deliver call k + 2 a rank must have passed its wait of call k + 1,
// hence seen MY export of call k + 1, which my stream runs after everything that reads call k.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_kernels.cuh"

namespace p3d {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ uint4 *recv_slot(char *base, int parity, int world, int src, int64_t n) {
    return reinterpret_cast<uint4 *>(base + kPeerDataOffset) + ((int64_t)parity * world + src) * (n + 1);
}

__global__ void __launch_bounds__(256)
    k_export_p2p(PeerParams pp, const uint4 *__restrict__ first_plane, const McHeader *header, unsigned int epoch) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int parity = (int)(epoch & 1u);
    if (i < pp.n) {
        if (pp.rank > 0) {  // my first-plane table -> the rank below me
            const uint4 t = first_plane[i];
            recv_slot(pp.peer[pp.rank - 1], parity, pp.world, pp.rank, pp.n)[i] = make_uint4(t.x, t.y, t.z, 0u);
        }
    } else if (i < pp.n + pp.world) {  // my {V, F} -> every rank (mine included)
        const int t = (int)(i - pp.n);
        const unsigned long long v = header->total_v, f = header->total_f;
        recv_slot(pp.peer[t], parity, pp.world, pp.rank, pp.n)[pp.n] =
            make_uint4((uint32_t)v, (uint32_t)(v >> 32), (uint32_t)f, (uint32_t)(f >> 32));
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int *ctl = reinterpret_cast<unsigned int *>(pp.peer[pp.rank]);
        const unsigned int arrived = atomicAdd(ctl + kPeerDoneWord, 1u);
        if (arrived == gridDim.x - 1) {  // every CTA's stores are fenced: raise my flag everywhere
            ctl[kPeerDoneWord] = 0u;
            __threadfence_system();
            for (int t = 0; t < pp.world; ++t)
                st_release_sys(reinterpret_cast<unsigned int *>(pp.peer[t]) + parity * kMaxPeers + pp.rank, epoch);
        }
    }
}

__global__ void k_wait_p2p(unsigned int *ctl, int world, unsigned int epoch, unsigned int max_spins) {
    const int t = threadIdx.x;
    if (t >= world) return;
    const unsigned int *flag = ctl + (epoch & 1u) * kMaxPeers + t;
    for (unsigned int spins = 0; ld_acquire_sys(flag) != epoch; ++spins) {
        if (spins >= max_spins) {  // a peer that never delivers must not hang the device
            ctl[kPeerTimeoutWord] = epoch;
            return;
        }
        __nanosleep(200);
    }
}

}  // namespace

void launch_export_p2p(const PeerParams &pp, const McWorkspace &ws, uint32_t epoch, cudaStream_t s) {
    const int64_t items = pp.n + pp.world;
    k_export_p2p<<<(unsigned)((items + 255) / 256), 256, 0, s>>>(pp, ws.ptab, ws.header, epoch);
}

void launch_wait_p2p(char *own_base, int world, uint32_t epoch, cudaStream_t s) {
    // ~200 ns per spin: a few seconds before giving up
    k_wait_p2p<<<1, kMaxPeers, 0, s>>>(reinterpret_cast<unsigned int *>(own_base), world, epoch, 5u * 1000u * 1000u);
}

}  // namespace p3d
