// Block-level building blocks shared by the selection kernels (select.cu) and the batched refine kernel (batch.cu).
#pragma once
#include "kernels.cuh"

namespace svsb {

constexpr int SEL_THREADS = 1024;
constexpr int HIST_BITS = 11;
constexpr int HIST_BINS = 1 << HIST_BITS;           // 2048
constexpr int SORT_CAP = K_FAST_MAX;                // 2048 keys sorted in shared memory

struct SelectSmem {
    u64 sortbuf[SORT_CAP];
    int64_t payload[SORT_CAP];                       // merge kernel only
    uint32_t hist[HIST_BINS];
    u64 red_a[32];
    u64 red_b[32];
    u64 bcast64[2];
    uint32_t counter;
    int32_t bcast32[4];
};

__device__ __forceinline__ int bitlen64(u64 v) { return v ? 64 - __clzll((long long)v) : 0; }

// Bitonic sort, descending, of buf[0..npow2) (npow2 a power of two <= SORT_CAP); optional payload.
// Thread t owns elements t, t + blockDim, ...: for j < 32 both partners of a compare-exchange live in the
// same 32-aligned block, i.e. in the same warp, so those steps need only __syncwarp(); block-wide barriers
// are paid only around the j >= 32 steps (6 instead of 28 for 128 keys, 27 instead of 66 for 2048).
template <bool PAYLOAD>
__device__ inline void block_bitonic_desc(u64* buf, int64_t* pay, int npow2) {
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const u64 a = buf[i], b = buf[ixj];
                    const bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) {
                        buf[i] = b; buf[ixj] = a;
                        if (PAYLOAD) { int64_t t = pay[i]; pay[i] = pay[ixj]; pay[ixj] = t; }
                    }
                }
            }
            if (j > 32 || (j == 32) || (j == 1 && (k << 1) > 32)) __syncthreads();   // next step crosses warps
            else __syncwarp();
        }
    }
    __syncthreads();
}

// Sort c <= RANK_SORT_MAX distinct keys descending by counting: rank(i) = #{j : key_j > key_i}.
// One barrier, c broadcast shared-memory reads per participating thread.  src and dst must not alias.
constexpr int RANK_SORT_MAX = 256;
__device__ inline void block_rank_sort_desc(const u64* src, u64* dst, int c) {
    if ((int)threadIdx.x < c) {
        const u64 mine = src[threadIdx.x];
        int rank = 0;
        for (int j = 0; j < c; ++j) rank += (src[j] > mine) ? 1 : 0;
        dst[rank] = mine;
    }
    __syncthreads();
}

// The rank-th largest (1-based) of ONE key per thread, keys pairwise distinct; blockDim.x a multiple of 32, <= 1024.
// Every warp sorts its 32 keys with a shuffle bitonic network; a key's global rank is its position in its own warp's
// list plus, per other warp, the number of greater keys there (binary search in shared memory) -- ~35 shared-memory
// reads per thread instead of blockDim.x.  sorted: blockDim.x keys, bcast: one key (shared).  All threads get the result.
template <class K>
__device__ inline K block_select_unique(K key, int rank, K* sorted, K* bcast) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const K o = __shfl_xor_sync(0xffffffffu, key, j);
            const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
            key = keep_max ? (key > o ? key : o) : (key < o ? key : o);
        }
    }
    sorted[tid] = key;                                         // warp w's list, descending: sorted[32 w .. 32 w + 32)
    __syncthreads();
    int r = lane;
    for (int w = 0; w < nw; ++w) {
        if (w == warp) continue;
        const K* L = sorted + 32 * w;
        int lo = 0, hi = 32;                                   // entries of L greater than key
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (L[mid] > key) lo = mid + 1; else hi = mid; }
        r += lo;
    }
    if (r == rank - 1) *bcast = key;
    __syncthreads();
    return *bcast;
}

constexpr int KTH_BINS = 2048;
constexpr int KTH_SMALL = 256;
constexpr int KTH_UNROLL = 1;

// kk-th largest (1-based, duplicates counted) of the 32-bit ordered values load(i), i < count (count >= kk >= 1).
// Range-adaptive radix select: the 2048 bins always span [min, max] of the values still in play, so scores packed
// into a narrow band (the README recipe: 0.75 +- 0.007) spread over all bins instead of hammering two of them with
// shared-memory atomics; typically min/max pass + one histogram pass + one collect pass.
// All threads call it; all get the result.  hist: KTH_BINS words, scratch: 72 words, small: KTH_SMALL words (shared).
template <class Load>
__device__ inline uint32_t block_kth_largest_o32(Load load, int64_t count, int kk, uint32_t* hist, uint32_t* scratch, uint32_t* small) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    uint32_t mn = 0xffffffffu, mx = 0u;
    // every scan issues KTH_UNROLL independent loads before consuming them: the passes are latency-, not bandwidth-bound
    for (int64_t i0 = tid; i0 < count; i0 += (int64_t)blockDim.x * KTH_UNROLL) {
        uint32_t o[KTH_UNROLL];
#pragma unroll
        for (int u = 0; u < KTH_UNROLL; ++u) { const int64_t i = i0 + (int64_t)u * blockDim.x; o[u] = i < count ? load(i) : 0u; }
#pragma unroll
        for (int u = 0; u < KTH_UNROLL; ++u) if (i0 + (int64_t)u * blockDim.x < count) { mn = min(mn, o[u]); mx = max(mx, o[u]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if (lane == 0) { scratch[8 + warp] = mn; scratch[40 + warp] = mx; }
    __syncthreads();
    if (warp == 0) {
        mn = lane < nwarps ? scratch[8 + lane] : 0xffffffffu;
        mx = lane < nwarps ? scratch[40 + lane] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
        if (lane == 0) { scratch[4] = mn; scratch[5] = mx; }
    }
    __syncthreads();
    uint32_t lo = scratch[4], hi = scratch[5];
    uint32_t remaining = (uint32_t)kk;
    __syncthreads();
    while (true) {
        const uint32_t span = hi - lo;
        const int shift = max(0, (span ? 32 - __clz(span) : 0) - 11);
        for (int i = tid; i < KTH_BINS; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int64_t i0 = tid; i0 < count; i0 += (int64_t)blockDim.x * KTH_UNROLL) {
            uint32_t o[KTH_UNROLL];
#pragma unroll
            for (int u = 0; u < KTH_UNROLL; ++u) { const int64_t i = i0 + (int64_t)u * blockDim.x; o[u] = i < count ? load(i) : 0u; }
#pragma unroll
            for (int u = 0; u < KTH_UNROLL; ++u)
                if (i0 + (int64_t)u * blockDim.x < count && o[u] >= lo && o[u] <= hi) atomicAdd(&hist[(o[u] - lo) >> shift], 1u);
        }
        __syncthreads();
        if (warp == 0) {
            // lane l owns bins [64 l, 64 l + 64); scan from the top bin down
            uint32_t mine = 0;
            for (int j = 0; j < 64; ++j) mine += hist[lane * 64 + j];
            uint32_t incl = mine;                                      // suffix sum over lanes >= l
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_down_sync(0xffffffffu, incl, o); if (lane + o < 32) incl += t; }
            const uint32_t above = incl - mine;
            if (above < remaining && incl >= remaining) {              // exactly one lane
                uint32_t acc = above;
                int bin = lane * 64 + 63;
                for (; bin > lane * 64; --bin) { if (acc + hist[bin] >= remaining) break; acc += hist[bin]; }
                scratch[0] = (uint32_t)bin; scratch[1] = acc; scratch[2] = hist[bin];
            }
        }
        __syncthreads();
        const uint32_t bin = scratch[0], above = scratch[1], inbin = scratch[2];
        remaining -= above;
        const uint32_t nlo = lo + (bin << shift);
        uint32_t nhi = nlo + ((1u << shift) - 1u);
        if (nhi > hi || nhi < nlo) nhi = hi;
        lo = nlo; hi = nhi;
        __syncthreads();
        if (shift == 0) return lo;                                     // the bin is one value
        if (inbin <= (uint32_t)KTH_SMALL) break;
    }
    // finish: the <= KTH_SMALL values of the final bin, rank-selected (duplicates allowed)
    if (tid == 0) scratch[3] = 0;
    __syncthreads();
    for (int64_t i0 = tid; i0 < count; i0 += (int64_t)blockDim.x * KTH_UNROLL) {
        uint32_t o[KTH_UNROLL];
#pragma unroll
        for (int u = 0; u < KTH_UNROLL; ++u) { const int64_t i = i0 + (int64_t)u * blockDim.x; o[u] = i < count ? load(i) : 0u; }
#pragma unroll
        for (int u = 0; u < KTH_UNROLL; ++u)
            if (i0 + (int64_t)u * blockDim.x < count && o[u] >= lo && o[u] <= hi) {
                const uint32_t p = atomicAdd(&scratch[3], 1u); if (p < (uint32_t)KTH_SMALL) small[p] = o[u];
            }
    }
    __syncthreads();
    const int c = (int)min(scratch[3], (uint32_t)KTH_SMALL);
    if (tid < c) {
        const uint32_t mine = small[tid];
        uint32_t gt = 0, ge = 0;
        for (int j = 0; j < c; ++j) { gt += small[j] > mine ? 1u : 0u; ge += small[j] >= mine ? 1u : 0u; }
        if (gt < remaining && ge >= remaining) scratch[6] = mine;      // equal values write the same word
    }
    __syncthreads();
    const uint32_t ans = scratch[6];
    __syncthreads();
    return ans;
}

}  // namespace svsb
