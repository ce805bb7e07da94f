// K1 -- the similarity kernel: scores = M . q for a row-major fp32 matrix, streamed once from HBM.
//
// Replaces `x = np.dot(embeddings_matrix, query_vec)` (reference src/svs/kb.py:1185, 1623), which is
// ~97 % of the reference's warm query time and purely DRAM-bound.  Algorithmic traffic: n*d*4 bytes read
// per query (+ 4n bytes of scores written, + one 64-bit atomic per warp-iteration for the group maxima
// that seed the exact top-k, see select.cu).  No tensor cores: one query is a GEMV, 0.5 FLOP/byte.
//
// Two implementations of the same contract (kernels.cuh: launch_gemv):
//   variant 1  "ldg" : persistent grid; each warp owns R consecutive rows per iteration and issues
//                      R*U independent 128-bit streaming loads (ld.global.nc.L1::no_allocate) per lane
//                      before consuming them; query vector in shared memory; warp-shuffle reduction.
//   variant 2  "tma" : persistent grid, one CTA per SM; a producer thread streams contiguous tiles of
//                      rows into a shared-memory ring with cp.async.bulk (TMA bulk copy, UBLKCP) completing
//                      on mbarriers; consumer warps read the tile and the query from shared memory.
// Which one runs by default is decided by measurement (profiles/), not by taste.
#include "kernels.cuh"
#include <cstdlib>

namespace svsb {

// =============================================================================================
// variant 1: LDG.128 streaming
// =============================================================================================
template <int R, int U, int THREADS>
__global__ void __launch_bounds__(THREADS)
gemv_ldg_kernel(const float4* __restrict__ M, int64_t n, int d4, const float4* __restrict__ q,
                float* __restrict__ scores, u64* __restrict__ gmax, int group_shift, const uint8_t* __restrict__ live)
{
    extern __shared__ float4 sq[];
    pdl_wait();                                   // the query (and the scores buffer) belong to the previous kernels
    pdl_trigger();
    for (int c = threadIdx.x; c < d4; c += THREADS) sq[c] = __ldcg(q + c);   // L2 only: see gemv_tma_kernel
    __syncthreads();

    constexpr int WARPS = THREADS / 32;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    const int64_t ngroups = (n + R - 1) / R;

    for (int64_t g = warp_global; g < ngroups; g += nwarps) {
        const int64_t r0 = g * R;
        const float4* rowp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int64_t rr = r0 + r; if (rr > n - 1) rr = n - 1;     // tail rows alias the last row (discarded)
            rowp[r] = M + rr * (int64_t)d4;
        }
        float4 acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);

        int c = lane;
        for (; c + 32 * (U - 1) < d4; c += 32 * U) {
            float4 m[U][R];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R; ++r) m[u][r] = ldg_stream(rowp[r] + c + 32 * u);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float4 qv = sq[c + 32 * u];
#pragma unroll
                for (int r = 0; r < R; ++r) fma4(acc[r], m[u][r], qv);
            }
        }
        for (; c < d4; c += 32) {
            float4 m[R];
#pragma unroll
            for (int r = 0; r < R; ++r) m[r] = ldg_stream(rowp[r] + c);
            const float4 qv = sq[c];
#pragma unroll
            for (int r = 0; r < R; ++r) fma4(acc[r], m[r], qv);
        }

        u64 kmax = 0;
        float mine = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float s = warp_sum((acc[r].x + acc[r].y) + (acc[r].z + acc[r].w));
            if (live && r0 + r < n && !live[r0 + r]) s = dead_score();      // tombstoned row: sorts below every real score
            if (r0 + r < n) {
                u64 key = make_key(s, (uint32_t)(r0 + r));
                kmax = key > kmax ? key : kmax;
            }
            if (lane == r) mine = s;
        }
        if (lane < R && r0 + lane < n) scores[r0 + lane] = mine;
        if (lane == 0) atomicMax(&gmax[r0 >> group_shift], kmax);
    }
}

// =============================================================================================
// variant 2: TMA bulk copy ring
// =============================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, u64 policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// NQ > 0: each lane keeps its NQ float4 of the query in registers (d4 <= 32 * NQ), so the only shared
// memory traffic is the row itself (one LDS.128 per 16 bytes streamed).  NQ == 0: query read from smem.
template <int CW, int NQ>
__global__ void __launch_bounds__((CW + 1) * 32, 1)
gemv_tma_kernel(const float* __restrict__ M, int64_t n, int d4, int tile_rows, int stages,
                const float* __restrict__ q, float* __restrict__ scores, u64* __restrict__ gmax, int group_shift,
                const uint8_t* __restrict__ live)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t row_bytes = (uint32_t)d4 * 16u;
    const uint32_t stage_bytes = (uint32_t)tile_rows * row_bytes;
    float4* sq = reinterpret_cast<float4*>(smem_raw + (size_t)stages * stage_bytes);   // used only when NQ == 0
    uint64_t* full = reinterpret_cast<uint64_t*>(sq + (NQ == 0 ? d4 : 0));
    uint64_t* empty = full + stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // Programmatic dependent launch (one-query-in-flight chains): the matrix does not depend on the previous kernel, so
    // the producer starts streaming tiles at once; whoever reads the query or writes scores / group maxima waits first.
    if (NQ == 0) {
        pdl_wait();
        for (int c = threadIdx.x; c < d4; c += blockDim.x) sq[c] = __ldcg(reinterpret_cast<const float4*>(q) + c);
    }
    if (threadIdx.x == 0) {
        pdl_trigger();
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t ntiles = (n + tile_rows - 1) / tile_rows;
    if (warp == CW) {
        // ---- producer: one thread streams tiles of consecutive rows (contiguous bytes) into the ring
        if (lane == 0) {
            u64 policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            int s = 0; uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                mbar_wait(&empty[s], phase ^ 1u);
                const int64_t r0 = t * tile_rows;
                const int64_t left = n - r0;
                const uint32_t rows = left < tile_rows ? (uint32_t)left : (uint32_t)tile_rows;
                const uint32_t bytes = rows * row_bytes;
                mbar_arrive_expect_tx(&full[s], bytes);
                bulk_g2s(smem_raw + (size_t)s * stage_bytes, M + r0 * (int64_t)d4 * 4, bytes, &full[s], policy);
                if (++s == stages) { s = 0; phase ^= 1u; }
            }
        }
    } else {
        // ---- consumers: warp w takes rows w, w+CW, ... of each tile
        float4 qr[NQ > 0 ? NQ : 1];
        if (NQ > 0) {
            pdl_wait();
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const int c = lane + 32 * j;
                // ld.global.cg, not the read-only (.nc) path: under programmatic dependent launch this grid is resident
                // while the previous kernel WRITES q, which the read-only path is not allowed to observe (measured:
                // 1 query in ~1000 came out with a mixture of old and new q, scripts/pdl_stress.py)
                qr[j] = c < d4 ? __ldcg(reinterpret_cast<const float4*>(q) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        int s = 0; uint32_t phase = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int64_t r0 = t * tile_rows;
            const int64_t left = n - r0;
            const int rows = left < tile_rows ? (int)left : tile_rows;
            mbar_wait(&full[s], phase);
            const float4* tile = reinterpret_cast<const float4*>(smem_raw + (size_t)s * stage_bytes);
            u64 kmax = 0;
            for (int r = warp; r < rows; r += CW) {
                const float4* p = tile + (size_t)r * d4;
                // tombstone byte of this row (svsb_apply_mutations); issued ahead of the row's arithmetic
                const uint8_t alive = live ? __ldg(live + r0 + r) : (uint8_t)1;
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                if (NQ > 0) {
                    float4 m[NQ > 0 ? NQ : 1];
#pragma unroll
                    for (int j = 0; j < NQ; ++j) {
                        const int c = lane + 32 * j;
                        m[j] = c < d4 ? p[c] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < NQ; ++j) { if (j & 1) fma4(a1, m[j], qr[j]); else fma4(a0, m[j], qr[j]); }
                } else {
                    int c = lane;
                    for (; c + 32 < d4; c += 64) { fma4(a0, p[c], sq[c]); fma4(a1, p[c + 32], sq[c + 32]); }
                    if (c < d4) fma4(a0, p[c], sq[c]);
                }
                float sc = warp_sum(((a0.x + a1.x) + (a0.y + a1.y)) + ((a0.z + a1.z) + (a0.w + a1.w)));
                if (!alive) sc = dead_score();
                if (lane == 0) {
                    scores[r0 + r] = sc;
                    const u64 k0 = make_key(sc, (uint32_t)(r0 + r));
                    kmax = k0 > kmax ? k0 : kmax;
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&empty[s]);
                if (kmax) atomicMax(&gmax[r0 >> group_shift], kmax);
            }
            if (++s == stages) { s = 0; phase ^= 1u; }
        }
    }
}

// =============================================================================================
// host side
// =============================================================================================
int sm_count(int device) {
    static int cache[64] = {0};
    if (device < 0 || device >= 64) device = 0;
    if (cache[device] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        cache[device] = v;
    }
    return cache[device];
}

template <int R, int U, int THREADS>
static cudaError_t run_ldg(cudaStream_t st, int device, const float* M, int64_t n, int d4, const float* q,
                           float* scores, u64* gmax, int group_shift, const uint8_t* live, int blocks_per_sm)
{
    auto kern = gemv_ldg_kernel<R, U, THREADS>;
    const size_t smem = (size_t)d4 * 16;
    static thread_local int configured_dev = -1;   // opt-in to >48 KB dynamic smem (d > 3072)
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    (void)configured_dev;
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) occ = 1;
    if (blocks_per_sm > 0 && blocks_per_sm < occ) occ = blocks_per_sm;
    constexpr int WARPS = THREADS / 32;
    const int64_t ngroups = (n + R - 1) / R;
    int64_t grid = (int64_t)sm_count(device) * occ;
    const int64_t need = (ngroups + WARPS - 1) / WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    e = launch_kernel(kern, dim3((unsigned)grid), dim3(THREADS), smem, st, reinterpret_cast<const float4*>(M), n, d4,
                      reinterpret_cast<const float4*>(q), scores, gmax, group_shift, live);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <int CW, int NQ>
static cudaError_t run_tma_inst(cudaStream_t st, int64_t grid, size_t smem, const float* M, int64_t n, int d4,
                                int tile_rows, int stages, const float* q, float* scores, u64* gmax, int group_shift,
                                const uint8_t* live)
{
    auto kern = gemv_tma_kernel<CW, NQ>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = launch_kernel(kern, dim3((unsigned)grid), dim3((CW + 1) * 32), smem, st, M, n, d4, tile_rows, stages, q, scores, gmax, group_shift, live);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

// tile_rows / stages: 0 = default.  cw: consumer warps (8 or 16; 0 = default).
static cudaError_t run_tma(cudaStream_t st, int device, const float* M, int64_t n, int d4, const float* q,
                           float* scores, u64* gmax, int group_shift, const uint8_t* live, int tile_rows, int stages, int cw,
                           int reserve_sms)
{
    const size_t row_bytes = (size_t)d4 * 16;
    const int nq = d4 <= 64 ? 2 : d4 <= 192 ? 6 : d4 <= 384 ? 12 : d4 <= 768 ? 24 : 0;
    const size_t q_bytes = nq == 0 ? row_bytes : 0;
    const size_t budget = 224 * 1024 - q_bytes - 256;       // ring bytes available next to q + barriers
    // Measured on B200 (profiles/r01_tune_gemv.md): ~48 KB tiles x 3 stages (~144 KB in flight per SM) is
    // the optimum; deeper rings are SLOWER (more concurrent DRAM streams), shallower ones starve.
    if (tile_rows <= 0) {
        tile_rows = 64;                                       // power of two <= 64
        while (tile_rows > 1 && (size_t)tile_rows * row_bytes > 48 * 1024) tile_rows >>= 1;
    }
    if (stages <= 0) {
        stages = (int)(budget / ((size_t)tile_rows * row_bytes));
        if (stages > 3) stages = 3;
    }
    if (tile_rows > 64 || (tile_rows & (tile_rows - 1))) return cudaErrorInvalidConfiguration;
    if (stages < 2 || stages > 16 || (size_t)stages * tile_rows * row_bytes > budget) return cudaErrorInvalidConfiguration;
    const size_t smem = (size_t)stages * tile_rows * row_bytes + q_bytes + (size_t)stages * 16 + 64;
    const int64_t ntiles = (n + tile_rows - 1) / tile_rows;
    int64_t grid = sm_count(device) - reserve_sms;          // reserve_sms: SMs left free for a concurrent selection kernel
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
#define SVSB_TMA_CASE(CWV, NQV) \
    return run_tma_inst<CWV, NQV>(st, grid, smem, M, n, d4, tile_rows, stages, q, scores, gmax, group_shift, live)
    if (cw == 16) {
        switch (nq) { case 2: SVSB_TMA_CASE(16, 2); case 6: SVSB_TMA_CASE(16, 6); case 12: SVSB_TMA_CASE(16, 12);
                      case 24: SVSB_TMA_CASE(16, 24); default: SVSB_TMA_CASE(16, 0); }
    }
    switch (nq) { case 2: SVSB_TMA_CASE(8, 2); case 6: SVSB_TMA_CASE(8, 6); case 12: SVSB_TMA_CASE(8, 12);
                  case 24: SVSB_TMA_CASE(8, 24); default: SVSB_TMA_CASE(8, 0); }
#undef SVSB_TMA_CASE
}

// Force the (lazily loaded) similarity kernels onto the device now.  With CUDA's lazy module loading the FIRST launch
// of a kernel synchronises the context; issued while another engine of this process has a merge kernel spinning on a
// peer flag, that first launch would wait for the spin to end -- which may be waiting for this very launch.
cudaError_t preload_gemv_kernels()
{
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
#define SVSB_PRE(K) do { if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, K); } while (0)
    SVSB_PRE((gemv_tma_kernel<8, 2>)); SVSB_PRE((gemv_tma_kernel<8, 6>)); SVSB_PRE((gemv_tma_kernel<8, 12>));
    SVSB_PRE((gemv_tma_kernel<8, 24>)); SVSB_PRE((gemv_tma_kernel<8, 0>));
    SVSB_PRE((gemv_tma_kernel<16, 2>)); SVSB_PRE((gemv_tma_kernel<16, 6>)); SVSB_PRE((gemv_tma_kernel<16, 12>));
    SVSB_PRE((gemv_tma_kernel<16, 24>)); SVSB_PRE((gemv_tma_kernel<16, 0>));
    SVSB_PRE((gemv_ldg_kernel<4, 3, 256>));
#undef SVSB_PRE
    return e;
}

cudaError_t launch_gemv(cudaStream_t st, int device, const float* M, int64_t n, int d, int ld,
                        const float* q, float* scores, u64* gmax, int group_shift,
                        int variant, int tune_a, int tune_b, int reserve_sms, const uint8_t* live)
{
    (void)d;
    if (n <= 0) return cudaSuccess;
    const int d4 = ld / 4;
    if (variant == 0) {
        // rows longer than a TMA stage can hold fall back to the LDG kernel
        variant = ((size_t)d4 * 16 * 4 <= 200 * 1024) ? 2 : 1;
        // tuning knobs for the measurement harness (profiles/): variant and its two parameters
        if (const char* v = getenv("SVSB_GEMV_VARIANT")) { int x = atoi(v); if (x == 1 || x == 2) variant = x; }
        if (const char* v = getenv("SVSB_GEMV_TUNE_A")) tune_a = atoi(v);
        if (const char* v = getenv("SVSB_GEMV_TUNE_B")) tune_b = atoi(v);
    }
    if (variant == 2) {
        // tune_a = tile rows, tune_b = stages + 100 * consumer warps (e.g. 1604 = 16 warps, 4 stages)
        cudaError_t e = run_tma(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_a, tune_b % 100, tune_b / 100, reserve_sms);
        if (e != cudaErrorInvalidConfiguration) return e;
        (void)cudaGetLastError();
        tune_a = 0; tune_b = 0;                               // TMA knobs mean nothing to the LDG kernel
    }
    // LDG variant; tune_a selects the (R, U, THREADS) instantiation, tune_b caps blocks per SM
    switch (tune_a) {
        case 1:  return run_ldg<4, 2, 256>(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_b);
        case 2:  return run_ldg<4, 3, 256>(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_b);
        case 3:  return run_ldg<8, 1, 256>(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_b);
        case 4:  return run_ldg<8, 2, 256>(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_b);
        case 5:  return run_ldg<2, 4, 256>(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_b);
        case 6:  return run_ldg<2, 6, 256>(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_b);
        case 7:  return run_ldg<4, 3, 512>(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_b);
        case 8:  return run_ldg<4, 4, 128>(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_b);
        case 9:  return run_ldg<8, 3, 128>(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_b);
        default: return run_ldg<4, 3, 256>(st, device, M, n, d4, q, scores, gmax, group_shift, live, tune_b);
    }
}

}  // namespace svsb
