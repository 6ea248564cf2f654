// radix_sort.cuh -- hand-written onesweep LSD radix sort for (uint32 key, uint32 value) pairs.  No CUB/Thrust.
//
// Onesweep (Adinets & Merrill 2022): one histogram kernel reads the keys once and counts all digit places, one tiny
// kernel turns the counts into exclusive bin offsets, then ONE kernel per 8-bit digit does tile-local stable
// ranking + a decoupled look-back over earlier tiles' digit counts + the scatter, so each pass reads and writes the
// data exactly once (2 x 8 B per element per pass).
//
//   tile    = 256 threads x 16 keys = 4096 keys; tiles are handed out by an atomic ticket so that every tile a
//             block waits on belongs to a block that is already resident (forward progress of the look-back).
//   ranking = per warp, per round of 32 consecutive keys: __match_any_sync groups equal digits, the lowest lane of
//             each group bumps the warp-private digit counter in shared memory; a 256-thread pass then turns the
//             8 warp histograms into warp offsets.  Order inside a tile is (warp, round, lane) = input order, so the
//             sort is stable, which LSD needs.
//   look-back = status word per (tile, digit): [31:30] flag (0 empty, 1 tile aggregate, 2 inclusive prefix),
//             [29:0] count.  Flag and payload share one 32-bit word, so plain volatile accesses suffice.
//   scatter = keys/values are first written to shared memory in tile-sorted order, then streamed out so that
//             consecutive threads write consecutive addresses within each digit run.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vn {
namespace rs {

constexpr int kRadix = 256;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kItems = 16;
constexpr int kTile = kThreads * kItems;
constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagInclusive = 2u << 30;
constexpr uint32_t kValueMask = (1u << 30) - 1u;

// counts of every digit place in one read of the keys: ghist[pass * 256 + digit]
__global__ void __launch_bounds__(kThreads) k_histogram(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ ghist, int passes) {
    __shared__ uint32_t sh[4][kRadix];
    for (int i = threadIdx.x; i < 4 * kRadix; i += kThreads) (&sh[0][0])[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const uint32_t k = keys[i];
        for (int p = 0; p < passes; p++) atomicAdd(&sh[p][(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int p = 0; p < passes; p++) {
        const uint32_t c = sh[p][threadIdx.x];
        if (c) atomicAdd(&ghist[p * kRadix + threadIdx.x], c);
    }
}

__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* s_warp /*[8]*/, uint32_t* total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        const uint32_t t = s_warp[w];
        if (w < (int)warp) base += t;
        tot += t;
    }
    __syncthreads();
    if (total) *total = tot;
    return base + inc - v;
}

// exclusive scan of each pass's 256 counts, in place.  One block of 256 threads.
__global__ void __launch_bounds__(kThreads) k_scan_histogram(uint32_t* __restrict__ ghist, int passes) {
    __shared__ uint32_t s_warp[kWarps];
    for (int p = 0; p < passes; p++) {
        const uint32_t v = ghist[p * kRadix + threadIdx.x];
        const uint32_t e = block_exclusive_scan_256(v, s_warp, nullptr);
        ghist[p * kRadix + threadIdx.x] = e;
    }
}

// one digit pass.  bin_base = exclusive offsets of this pass's 256 digits; status = [num_tiles][256] zeroed words.
__global__ void __launch_bounds__(kThreads) k_onesweep_pass(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                           uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                           uint32_t n, int shift, const uint32_t* __restrict__ bin_base,
                                                           volatile uint32_t* status, uint32_t* tile_ticket) {
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warp_hist[kWarps][kRadix];
    __shared__ uint32_t s_digit_base[kRadix];
    __shared__ uint32_t s_global_base[kRadix];
    __shared__ uint32_t s_scan[kWarps];
    __shared__ uint32_t s_keys[kTile];
    __shared__ uint32_t s_vals[kTile];

    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_ticket, 1u);
    for (int i = tid; i < kWarps * kRadix; i += kThreads) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * (uint32_t)kTile;
    const uint32_t tile_n = min((uint32_t)kTile, n - tile_base);

    uint32_t key[kItems], val[kItems], rank[kItems];
    const uint32_t lane_lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kItems; r++) {
        const uint32_t local = warp * (32u * kItems) + (uint32_t)r * 32u + lane;
        const bool valid = local < tile_n;
        key[r] = valid ? keys_in[tile_base + local] : 0xFFFFFFFFu;
        val[r] = valid ? vals_in[tile_base + local] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kItems; r++) {
        const uint32_t local = warp * (32u * kItems) + (uint32_t)r * 32u + lane;
        const bool valid = local < tile_n;
        const uint32_t digit = (key[r] >> shift) & 255u;
        // out-of-range lanes vote with a value no real digit can have, so they never join a peer group
        const uint32_t vote = valid ? digit : (256u + lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, vote);
        const uint32_t before = __popc(peers & lane_lt);
        uint32_t prev = 0;
        if (valid) prev = s_warp_hist[warp][digit];
        __syncwarp();
        if (valid && before == 0) s_warp_hist[warp][digit] = prev + __popc(peers);
        __syncwarp();
        rank[r] = prev + before;
    }
    __syncthreads();

    // thread d owns digit d: warp histograms -> warp-exclusive offsets, tile count
    uint32_t tile_count = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
        const uint32_t c = s_warp_hist[w][tid];
        s_warp_hist[w][tid] = tile_count;
        tile_count += c;
    }
    // decoupled look-back over earlier tiles for this digit
    uint32_t exclusive = 0;
    if (tile == 0) {
        status[tid] = kFlagInclusive | tile_count;
    } else {
        status[(size_t)tile * kRadix + tid] = kFlagAggregate | tile_count;
        for (int t = (int)tile - 1; t >= 0; t--) {
            uint32_t s;
            do { s = status[(size_t)t * kRadix + tid]; } while ((s >> 30) == 0u);
            exclusive += s & kValueMask;
            if ((s >> 30) == 2u) break;
        }
        status[(size_t)tile * kRadix + tid] = kFlagInclusive | (exclusive + tile_count);
    }
    const uint32_t digit_base = block_exclusive_scan_256(tile_count, s_scan, nullptr);   // syncs inside
    s_digit_base[tid] = digit_base;
    s_global_base[tid] = bin_base[tid] + exclusive - digit_base;
    __syncthreads();

#pragma unroll
    for (int r = 0; r < kItems; r++) {
        const uint32_t local = warp * (32u * kItems) + (uint32_t)r * 32u + lane;
        if (local < tile_n) {
            const uint32_t digit = (key[r] >> shift) & 255u;
            const uint32_t pos = s_digit_base[digit] + s_warp_hist[warp][digit] + rank[r];
            s_keys[pos] = key[r];
            s_vals[pos] = val[r];
        }
    }
    __syncthreads();
    for (uint32_t j = tid; j < tile_n; j += kThreads) {
        const uint32_t k = s_keys[j];
        const uint32_t out = s_global_base[(k >> shift) & 255u] + j;
        keys_out[out] = k;
        vals_out[out] = s_vals[j];
    }
}

// Workspace: ghist[4*256] | ticket[4] | status[passes][tiles][256]
inline size_t workspace_bytes(uint64_t n) {
    const uint64_t tiles = (n + kTile - 1) / kTile;
    return (size_t)(4 * kRadix + 4 + 4ull * tiles * kRadix) * sizeof(uint32_t);
}

// Sorts n pairs by the low key_bits of the key.  (k0,v0) holds the input; (k1,v1) is the alternate buffer.  Returns
// 0 when the result is in (k0,v0) and 1 when it is in (k1,v1); *launches is incremented per kernel launched.
inline int sort_pairs(uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, uint32_t n, int key_bits, void* workspace,
                      cudaStream_t stream, int num_sms, uint32_t* launches) {
    if (n == 0) return 0;
    const int passes = (key_bits + 7) / 8;
    const uint32_t tiles = (n + kTile - 1) / kTile;
    uint32_t* ghist = (uint32_t*)workspace;
    uint32_t* ticket = ghist + 4 * kRadix;
    uint32_t* status = ticket + 4;
    cudaMemsetAsync(workspace, 0, (size_t)(4 * kRadix + 4 + (size_t)passes * tiles * kRadix) * sizeof(uint32_t), stream);
    const uint64_t hist_want = ((uint64_t)n + kThreads - 1) / kThreads, hist_cap = (uint64_t)num_sms * 8ull;
    const uint32_t hist_blocks = (uint32_t)(hist_want < hist_cap ? hist_want : hist_cap);
    k_histogram<<<hist_blocks, kThreads, 0, stream>>>(k0, n, ghist, passes);
    k_scan_histogram<<<1, kThreads, 0, stream>>>(ghist, passes);
    if (launches) *launches += 2;
    uint32_t *ki = k0, *vi = v0, *ko = k1, *vo = v1;
    for (int p = 0; p < passes; p++) {
        k_onesweep_pass<<<tiles, kThreads, 0, stream>>>(ki, vi, ko, vo, n, 8 * p, ghist + p * kRadix,
                                                        status + (size_t)p * tiles * kRadix, ticket + p);
        if (launches) *launches += 1;
        uint32_t* t;
        t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    return passes & 1;
}

}  // namespace rs
}  // namespace vn
