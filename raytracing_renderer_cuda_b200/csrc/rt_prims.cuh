// rt_prims.cuh — the two device-wide primitives the scene build and the output stage need, hand-written for sm_100a
// (round 1 called cub::DeviceScan / cub::DeviceRadixSort here):
//
//   exclusive_sum<T>      exclusive prefix sum of n values (T = uint32_t / unsigned long long), in place allowed
//   sort_pairs_u64_u32    stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits per pass
//
// Both are off the frame loop (LBVH build: Morton order and 4-wide compaction; JPEG writer: bit offsets of the blocks and
// of the stuffed bytes), both are bandwidth-trivial at their sizes (<= a few million elements), so the forms are the plain
// ones: a scan by tiles of 2048 with the tile totals scanned recursively, a radix sort by per-tile digit histograms,
// one scan of the digit-major histogram table and a stable scatter that ranks a tile's keys round by round.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtd {
namespace prims {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems; // 2048

// ------------------------------------------------------------------------------------------------ scan ----
// exclusive prefix of one value per thread over the CTA (kThreads threads); returns the CTA total in `total`
template <typename T>
__device__ __forceinline__ T cta_exclusive(T v, T* s_warp /* [kThreads / 32] */, T& total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T up = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= unsigned(off)) incl += up;
    }
    if (lane == 31u) s_warp[warp] = incl;
    __syncthreads();
    T before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const T c = s_warp[w];
        before += unsigned(w) < warp ? c : T(0);
        all += c;
    }
    total = all;
    __syncthreads(); // (s_warp may be reused by the caller's next call)
    return before + incl - v;
}

// one tile per CTA: out[i] = exclusive prefix WITHIN the tile, tile_sums[tile] = the tile's total (if tile_sums)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_scan_tiles(const T* in, T* out, size_t n, T* tile_sums) { // (in == out allowed: no __restrict__)
    __shared__ T s_warp[kThreads / 32];
    const size_t base = size_t(blockIdx.x) * kTile + size_t(threadIdx.x) * kItems;
    T v[kItems];
    T sum = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        v[k] = base + k < n ? in[base + k] : T(0);
        sum += v[k];
    }
    T total;
    T run = cta_exclusive<T>(sum, s_warp, total);
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (tile_sums && threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
template <typename T>
__global__ void __launch_bounds__(kThreads) k_scan_add(T* __restrict__ out, size_t n, const T* __restrict__ tile_offsets) {
    const T add = tile_offsets[blockIdx.x];
    const size_t base = size_t(blockIdx.x) * kTile;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const size_t i = base + size_t(k) * kThreads + threadIdx.x;
        if (i < n) out[i] += add;
    }
}

// elements of scratch exclusive_sum needs for n values
inline size_t scan_scratch_elems(size_t n) {
    size_t total = 0;
    while (n > size_t(kTile)) {
        n = (n + kTile - 1) / kTile;
        total += n;
    }
    return total + 1;
}
// out[i] = in[0] + ... + in[i - 1]; in == out allowed; scratch: scan_scratch_elems(n) elements of T
template <typename T>
inline cudaError_t exclusive_sum(const T* in, T* out, size_t n, T* scratch, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const size_t tiles = (n + kTile - 1) / kTile;
    if (tiles == 1) {
        k_scan_tiles<T><<<1, kThreads, 0, st>>>(in, out, n, nullptr);
        return cudaGetLastError();
    }
    k_scan_tiles<T><<<unsigned(tiles), kThreads, 0, st>>>(in, out, n, scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = exclusive_sum<T>(scratch, scratch, tiles, scratch + tiles, st); // the tile totals -> tile offsets, in place
    if (e != cudaSuccess) return e;
    k_scan_add<T><<<unsigned(tiles), kThreads, 0, st>>>(out, n, scratch);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------ radix sort ----
// digit histogram of every tile, digit-major: hist[digit * n_tiles + tile]
static __global__ void __launch_bounds__(kThreads) k_rs_hist(const unsigned long long* __restrict__ keys, uint32_t n, int shift,
                                                      uint32_t* __restrict__ hist, uint32_t n_tiles) {
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t base = blockIdx.x * uint32_t(kTile);
#pragma unroll
    for (int r = 0; r < kItems; ++r) {
        const uint32_t i = base + uint32_t(r) * kThreads + threadIdx.x;
        if (i < n) atomicAdd(&s_h[uint32_t(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[size_t(threadIdx.x) * n_tiles + blockIdx.x] = s_h[threadIdx.x];
}
// stable scatter of one tile: element order inside the tile = (round, warp, lane), i.e. index order
static __global__ void __launch_bounds__(kThreads) k_rs_scatter(const unsigned long long* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                         unsigned long long* __restrict__ kout, uint32_t* __restrict__ vout, uint32_t n,
                                                         int shift, const uint32_t* __restrict__ offs, uint32_t n_tiles) {
    __shared__ uint32_t s_base[256];                    // next output position of every digit for this tile
    __shared__ uint32_t s_wcount[kThreads / 32][256];   // keys of the current round, per warp and digit
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    s_base[threadIdx.x] = offs[size_t(threadIdx.x) * n_tiles + blockIdx.x];
    const uint32_t base = blockIdx.x * uint32_t(kTile);
#pragma unroll 1
    for (int r = 0; r < kItems; ++r) {
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s_wcount[w][threadIdx.x] = 0u;
        __syncthreads();
        const uint32_t i = base + uint32_t(r) * kThreads + threadIdx.x;
        const bool valid = i < n;
        unsigned long long key = 0ull;
        uint32_t val = 0u;
        if (valid) {
            key = kin[i];
            val = vin[i];
        }
        const uint32_t d = valid ? uint32_t(key >> shift) & 255u : 256u; // (lanes past the end only match each other)
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = uint32_t(__popc(peers & lt_mask));
        if (valid && rank == 0u) s_wcount[warp][d] = uint32_t(__popc(peers));
        __syncthreads();
        if (valid) {
            uint32_t pos = s_base[d] + rank;
            for (unsigned w = 0; w < warp; ++w) pos += s_wcount[w][d];
            kout[pos] = key;
            vout[pos] = val;
        }
        __syncthreads();
        uint32_t add = 0u;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) add += s_wcount[w][threadIdx.x];
        s_base[threadIdx.x] += add;
        __syncthreads();
    }
}

// scratch (uint32_t elements) sort_pairs_u64_u32 needs for n pairs
inline size_t sort_scratch_elems(size_t n) {
    const size_t tiles = (n + kTile - 1) / kTile;
    return 256 * tiles + scan_scratch_elems(256 * tiles);
}
// Sorts n pairs by the key bits [0, 8 * passes); ping-pongs between (keys, vals) and (keys_alt, vals_alt).  Returns in
// *in_alt whether the sorted pairs ended up in the alt buffers (odd number of passes).
inline cudaError_t sort_pairs_u64_u32(unsigned long long* keys, uint32_t* vals, unsigned long long* keys_alt, uint32_t* vals_alt, uint32_t n,
                                      int passes, uint32_t* scratch, bool* in_alt, cudaStream_t st) {
    *in_alt = false;
    if (n == 0) return cudaSuccess;
    const uint32_t tiles = (n + kTile - 1) / kTile;
    uint32_t* hist = scratch;
    uint32_t* scan_tmp = scratch + size_t(256) * tiles;
    for (int p = 0; p < passes; ++p) {
        const bool fwd = (p & 1) == 0;
        const unsigned long long* kin = fwd ? keys : keys_alt;
        const uint32_t* vin = fwd ? vals : vals_alt;
        unsigned long long* kout = fwd ? keys_alt : keys;
        uint32_t* vout = fwd ? vals_alt : vals;
        k_rs_hist<<<tiles, kThreads, 0, st>>>(kin, n, 8 * p, hist, tiles);
        cudaError_t e = exclusive_sum<uint32_t>(hist, hist, size_t(256) * tiles, scan_tmp, st);
        if (e != cudaSuccess) return e;
        k_rs_scatter<<<tiles, kThreads, 0, st>>>(kin, vin, kout, vout, n, 8 * p, hist, tiles);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        *in_alt = fwd;
    }
    return cudaSuccess;
}

} // namespace prims
} // namespace rtd
