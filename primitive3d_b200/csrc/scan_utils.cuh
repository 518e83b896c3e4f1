// primitive3d_b200/csrc/scan_utils.cuh -- warp primitives and the single-pass decoupled look-back
// scan shared by the marching-cubes and marching-tetrahedra kernels (no CUB / thrust).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p3d {

constexpr unsigned kFull = 0xffffffffu;

// Look-back status word: [63:62] flag, [61:0] value.  Flag and value travel in one 64-bit
// store/load, so no fence is needed between them.
constexpr uint64_t kFlagShift = 62;
constexpr uint64_t kFlagAggregate = 1ull << kFlagShift;
constexpr uint64_t kFlagPrefix = 2ull << kFlagShift;
constexpr uint64_t kValueMask = (1ull << kFlagShift) - 1;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ unsigned long long warp_incl_scan64(unsigned long long v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    return v;
}

__device__ __forceinline__ uint32_t warp_sum32(uint32_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    return v;
}

// Decoupled look-back (Merrill & Garland): called by ONE full warp of the CTA that owns `tile`.
// Publishes the tile's aggregate, walks back over predecessor tiles 32 at a time until one with
// an inclusive prefix is found, publishes this tile's inclusive prefix and returns its exclusive
// prefix.  Tiles must be handed out in increasing order to running CTAs (ticket counter), which
// guarantees every predecessor is resident or finished.  `status` must be zeroed beforehand.
__device__ __forceinline__ unsigned long long lookback(volatile unsigned long long *status, int64_t tile,
                                                       unsigned long long aggregate, int lane) {
    if (tile == 0) {
        if (lane == 0) status[0] = kFlagPrefix | aggregate;
        return 0ull;
    }
    if (lane == 0) status[tile] = kFlagAggregate | aggregate;
    unsigned long long excl = 0ull;
    for (int64_t j = tile - 1;; j -= 32) {
        const int64_t idx = j - lane;
        unsigned long long s;
        do {
            s = idx >= 0 ? status[idx] : kFlagPrefix;  // "tile -1" is an inclusive prefix of 0
        } while (__any_sync(kFull, (s >> kFlagShift) == 0ull));
        const unsigned have_prefix = __ballot_sync(kFull, (s >> kFlagShift) == 2ull);
        if (have_prefix) {
            const int first = __ffs(have_prefix) - 1;  // nearest predecessor with an inclusive prefix
            excl += warp_sum64(lane <= first ? (s & kValueMask) : 0ull);
            break;
        }
        excl += warp_sum64(s & kValueMask);
    }
    if (lane == 0) status[tile] = kFlagPrefix | (excl + aggregate);
    return excl;
}

// Function attributes (the opt-in shared-memory size) and occupancy are per DEVICE: a process that uses two GPUs must
// set them on both.  `setup` runs once per device and returns a positive value (e.g. resident CTAs per SM) that is
// cached; a host thread that races another one merely repeats the setup.
constexpr int kMaxDevices = 64;
template <typename Setup>
inline int per_device(int (&cache)[kMaxDevices], Setup setup) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) return setup();
    if (cache[dev] == 0) cache[dev] = setup();
    return cache[dev];
}

inline int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

}  // namespace p3d
