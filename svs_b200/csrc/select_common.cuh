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

}  // namespace svsb
