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
// variant 2b: the same ring, several queries per pass (small exact batches)
// =============================================================================================
// b queries against the matrix in ONE pass over HBM, bit-identical to b single-query passes: every (row, query) dot
// product is computed by one warp with exactly gemv_tma_kernel's arithmetic (lane l owns the float4 chunks l, l+32, ...;
// even chunks accumulate into a0, odd ones into a1; same combine, same xor tree).  The consumer warps are split into
// `groups` groups; a warp keeps BQ queries (its slice of each) in registers -- 4 * NQ * BQ of them, which is why the
// consumers run at 240 registers (setmaxnreg: the producer warpgroup gives its registers up) -- and the warps of a group
// share the rows of a tile.  Shared-memory traffic per streamed byte: 1 write (TMA) + `groups` reads, against ~5x
// headroom of shared-memory over HBM bandwidth per SM; the FMA pipe sees b FMAs per streamed float (b = 8: 40 %).
// Warps 0-3: producer warpgroup (one lane issues the bulk copies); warps 4-11: consumers.
// Two fp32 FMAs per instruction (Blackwell FFMA2, PTX fma.rn.f32x2): each half is an ordinary IEEE round-to-nearest FMA,
// so the bits equal two fmaf calls; what halves is the number of issue slots the b-fold arithmetic of a row needs.
struct f4x2 { u64 xy, zw; };                                    // a float4 held as two packed pairs
__device__ __forceinline__ void fma4_x2(f4x2& acc, const f4x2& a, const f4x2& b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.xy) : "l"(a.xy), "l"(b.xy));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.zw) : "l"(a.zw), "l"(b.zw));
}
__device__ __forceinline__ f4x2 as_f4x2(const ulonglong2& v) { f4x2 r; r.xy = v.x; r.zw = v.y; return r; }
__device__ __forceinline__ float4 as_float4(const f4x2& v) {
    float4 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v.xy));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.z), "=f"(r.w) : "l"(v.zw));
    return r;
}

// FULL: d4 == 32 * NQ exactly (d = 1536, 3072, 768, 256 ...): no per-chunk bounds predicate / zero select.
template <int NQ, int BQ, bool FULL>
__global__ void __launch_bounds__(384, 1)
gemv_tma_mq_kernel(const float* __restrict__ M, int64_t n, int d4, int tile_rows, int stages,
                   const float* __restrict__ Q, int b, int groups, float* __restrict__ scores, int64_t n_stride,
                   u64* __restrict__ gmax, int64_t g_stride, int group_shift, const uint8_t* __restrict__ live)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t row_bytes = (uint32_t)d4 * 16u;
    const uint32_t stage_bytes = (uint32_t)tile_rows * row_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * stage_bytes);
    uint64_t* empty = full + stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int CW = 8;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ntiles = (n + tile_rows - 1) / tile_rows;
    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (warp == 0 && lane == 0) {
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
        asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
        const int cw = warp - 4, grp = cw % groups, sub = cw / groups, nsub = CW / groups;
        f4x2 qr[BQ][NQ];
#pragma unroll
        for (int u = 0; u < BQ; ++u) {
            int qi = grp * BQ + u;
            if (qi > b - 1) qi = b - 1;                          // padding slots recompute the last query; never stored
            const ulonglong2* qp = reinterpret_cast<const ulonglong2*>(Q) + (int64_t)qi * d4;
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                const int c = lane + 32 * j;
                qr[u][j] = as_f4x2((FULL || c < d4) ? __ldcg(qp + c) : make_ulonglong2(0ull, 0ull));
            }
        }
        int s = 0; uint32_t phase = 0;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int64_t r0 = t * tile_rows;
            const int64_t left = n - r0;
            const int rows = left < tile_rows ? (int)left : tile_rows;
            mbar_wait(&full[s], phase);
            const ulonglong2* tile = reinterpret_cast<const ulonglong2*>(smem_raw + (size_t)s * stage_bytes);
            for (int r = sub; r < rows; r += nsub) {
                const ulonglong2* p = tile + (size_t)r * d4;
                const uint8_t alive = live ? __ldg(live + r0 + r) : (uint8_t)1;
                f4x2 a0[BQ], a1[BQ];
#pragma unroll
                for (int u = 0; u < BQ; ++u) { a0[u].xy = 0ull; a0[u].zw = 0ull; a1[u] = a0[u]; }
                // With 192 registers of queries there is no room to load the whole row slice ahead of the arithmetic (as
                // the single-query kernel does): keep PF loads in flight instead, issued a chunk ahead of their FMAs --
                // without it every chunk pays a shared-memory round trip with only two warps per scheduler to hide it
                // (measured: 2.1 ms instead of 1.0 ms for 8 queries at 1M x 1536).
                constexpr int PF = (4 * NQ * BQ >= 192) ? 2 : NQ;
                ulonglong2 mb[PF];
#pragma unroll
                for (int j = 0; j < PF && j < NQ; ++j) { const int c = lane + 32 * j; mb[j] = (FULL || c < d4) ? p[c] : make_ulonglong2(0ull, 0ull); }
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const f4x2 m = as_f4x2(mb[j % PF]);
                    if (j + PF < NQ) { const int c = lane + 32 * (j + PF); mb[j % PF] = (FULL || c < d4) ? p[c] : make_ulonglong2(0ull, 0ull); }
#pragma unroll
                    for (int u = 0; u < BQ; ++u) { if (j & 1) fma4_x2(a1[u], m, qr[u][j]); else fma4_x2(a0[u], m, qr[u][j]); }
                }
                // the BQ xor trees advance in lockstep (BQ independent shuffles per round instead of BQ serial trees), then
                // lane u publishes query u: one predicated block of stores for the whole row
                float v[BQ];
#pragma unroll
                for (int u = 0; u < BQ; ++u) {
                    const float4 x0 = as_float4(a0[u]), x1 = as_float4(a1[u]);
                    v[u] = ((x0.x + x1.x) + (x0.y + x1.y)) + ((x0.z + x1.z) + (x0.w + x1.w));
                }
                // BQ xor trees in one: at distance 16 a lane keeps the upper or the lower half of the queries and trades the
                // other half with its partner, at distance 8 a quarter, ... -- every addition is still x[l] + x[l ^ o] of ONE
                // query at the distances 16, 8, 4, 2, 1 in that order, i.e. exactly warp_sum's tree (fp32 addition is
                // commutative bit for bit), with 2*BQ - 2 + log2(32 / BQ) shuffles instead of 5 * BQ.
                int o = 16;
#pragma unroll
                for (int cnt = BQ; cnt > 1; cnt >>= 1, o >>= 1) {
                    const int half = cnt >> 1;
                    const bool up = (lane & o) != 0;
#pragma unroll
                    for (int i = 0; i < half; ++i) {
                        const float send = up ? v[i] : v[i + half];
                        const float keep = up ? v[i + half] : v[i];
                        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                }
#pragma unroll
                for (; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
                float sc = v[0];                                 // the sum of query u_of(lane), complete in every lane
                if (!alive) sc = dead_score();
                // which query this lane holds: bit 4 picked a half, bit 3 a quarter, ...; lanes with the low bits clear publish
                constexpr int LANES_PER_Q = 32 / BQ;
                int u_mine = 0;
#pragma unroll
                for (int step = 0, bit = 16, w = BQ >> 1; w >= 1; ++step, bit >>= 1, w >>= 1) u_mine += (lane & bit) ? w : 0;
                const int qi = grp * BQ + u_mine;
                if ((lane & (LANES_PER_Q - 1)) == 0 && qi < b) {
                    scores[(int64_t)qi * n_stride + r0 + r] = sc;
                    atomicMax(&gmax[(int64_t)qi * g_stride + ((r0 + r) >> group_shift)], make_key(sc, (uint32_t)(r0 + r)));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
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

// ---- multi-query launcher ---------------------------------------------------------------------
template <int NQ, int BQ, bool FULL>
static cudaError_t run_mq_inst2(cudaStream_t st, int64_t grid, size_t smem, const float* M, int64_t n, int d4, int tile_rows, int stages,
                               const float* Q, int b, int groups, float* scores, int64_t n_stride, u64* gmax, int64_t g_stride,
                               int group_shift, const uint8_t* live)
{
    auto kern = gemv_tma_mq_kernel<NQ, BQ, FULL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)grid, 384, smem, st>>>(M, n, d4, tile_rows, stages, Q, b, groups, scores, n_stride, gmax, g_stride, group_shift, live);
    count_launch();
    return cudaGetLastError();
}

template <int NQ, int BQ>
static cudaError_t run_mq_inst(cudaStream_t st, int64_t grid, size_t smem, const float* M, int64_t n, int d4, int tile_rows, int stages,
                               const float* Q, int b, int groups, float* scores, int64_t n_stride, u64* gmax, int64_t g_stride,
                               int group_shift, const uint8_t* live)
{
    if (d4 == 32 * NQ)
        return run_mq_inst2<NQ, BQ, true>(st, grid, smem, M, n, d4, tile_rows, stages, Q, b, groups, scores, n_stride, gmax, g_stride, group_shift, live);
    return run_mq_inst2<NQ, BQ, false>(st, grid, smem, M, n, d4, tile_rows, stages, Q, b, groups, scores, n_stride, gmax, g_stride, group_shift, live);
}

int gemv_mq_queries_per_warp(int ld) {
    const int d4 = ld / 4;
    return d4 <= 64 ? 8 : d4 <= 192 ? 4 : d4 <= 384 ? 4 : d4 <= 768 ? 2 : 0;   // 16 * NQ * BQ query + 8 * BQ accumulator registers <= ~224
}

cudaError_t launch_gemv_mq(cudaStream_t st, int device, const float* M, int64_t n, int d, int ld, const float* Q, int b,
                           float* scores, int64_t n_stride, u64* gmax, int64_t g_stride, int group_shift, const uint8_t* live,
                           int reserve_sms)
{
    (void)d;
    if (n <= 0 || b <= 0) return cudaSuccess;
    const int d4 = ld / 4;
    const int bq_hi = gemv_mq_queries_per_warp(ld);
    if (bq_hi == 0) return cudaErrorInvalidConfiguration;       // rows too long for register-resident queries
    int bq = bq_hi;
    if (bq_hi > 2 && b <= bq_hi / 2) bq = bq_hi / 2;             // fewer wasted FMAs for a small batch
    int groups = 1;
    while (groups * bq < b) groups <<= 1;
    if (groups > 8) return cudaErrorInvalidConfiguration;       // the caller chunks larger batches
    const size_t row_bytes = (size_t)d4 * 16;
    int tile_rows = 64;
    while (tile_rows > 1 && (size_t)tile_rows * row_bytes > 48 * 1024) tile_rows >>= 1;
    if (tile_rows < 8 / groups) { /* every warp of a group still gets whole rows: fine, some idle */ }
    const int stages = 3;
    const size_t smem = (size_t)stages * tile_rows * row_bytes + (size_t)stages * 16 + 64;
    const int64_t ntiles = (n + tile_rows - 1) / tile_rows;
    int64_t grid = sm_count(device) - reserve_sms;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    const int nq = d4 <= 64 ? 2 : d4 <= 192 ? 6 : d4 <= 384 ? 12 : 24;
#define SVSB_MQ_CASE(NQV, BQV) \
    return run_mq_inst<NQV, BQV>(st, grid, smem, M, n, d4, tile_rows, stages, Q, b, groups, scores, n_stride, gmax, g_stride, group_shift, live)
    switch (nq * 100 + bq) {
        case 204: SVSB_MQ_CASE(2, 4);
        case 208: SVSB_MQ_CASE(2, 8);
        case 602: SVSB_MQ_CASE(6, 2);
        case 604: SVSB_MQ_CASE(6, 4);
        case 1202: SVSB_MQ_CASE(12, 2);
        case 1204: SVSB_MQ_CASE(12, 4);
        case 2402: SVSB_MQ_CASE(24, 2);
        default: return cudaErrorInvalidConfiguration;
    }
#undef SVSB_MQ_CASE
}

}  // namespace svsb
