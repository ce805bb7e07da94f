// K2 -- the batched path's coarse contraction: S~ = M16 . Q16^T on the 5th-generation tensor cores.
//
// The reference has no batched API (a batch is a Python loop over `retrieve`, src/svs/kb.py:1608-1640); for b
// queries the work np.dot(M, q) x b (src/svs/kb.py:1623) is a dense contraction of 2*n*d*b FLOP, so it belongs on
// tcgen05.  The contract, however, is the fp32 ranking (<= 1e-5 relative), which no single-pass 16/19-bit product
// gives.  This file computes a COARSE score for every (row, query) pair with a proven error bound and keeps only
// the pairs that can still matter; batch.cu re-scores those exactly in fp32.
//
//   operands   M16 = fp16(M * 2^12) [n][ld16], Q16 = fp16(Q * 2^12) [b_pad][ld16]   (K-major, zero padded)
//   kernel     persistent, warp-specialised, one CTA per SM, cta_group::1:
//                warp 16  TMA producer: cp.async.bulk.tensor 2-D boxes (64 halfs x 128 rows of M16, 64 x 256 rows of
//                         Q16), 128-byte swizzle, 4-stage shared-memory ring on mbarriers
//                warp 17  MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128 (rows) x N=256 (queries) x K=16, fp32
//                         accumulators in TMEM, two accumulator buffers (2 x 256 columns) so that the epilogue of tile
//                         i overlaps the MMAs of tile i+1; tcgen05.commit releases ring slots / publishes accumulators
//                warps 0-15 epilogue (4 per TMEM lane quarter, 2 column chunks each): tcgen05.ld 32 lanes x 32 columns
//                         at a time; S~ is NEVER written out:
//                         MODE_FILTER  compare with the query's threshold (shared memory), append the rare survivors
//                                      as 64-bit keys (ordered coarse score << 32 | ~row) to the query's candidate list
//                         MODE_SAMPLE  dump raw coarse scores of the sampled row tiles (bootstrap of the thresholds)
//   roofline   tensor pipe: 2*n*d*b FLOP per batch; HBM floor n*ld16*2 bytes per pass.
#include "kernels.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

namespace svsb {

// ---------------------------------------------------------------------------------------------
// small PTX wrappers (sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cvta_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarrier_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarrier_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbarrier_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// Bounded wait (2 s of %globaltimer): a protocol bug must surface as a CUDA error, never as a hung GPU.
__device__ __forceinline__ void mbarrier_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    u64 t0 = 0;
    for (uint32_t spins = 0; ; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if ((spins & 1023u) == 1023u) {
            u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, u64 adesc, u64 bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tile geometry
// ---------------------------------------------------------------------------------------------
constexpr int CG_BM = 128;                 // rows (documents) per tile  = UMMA M
constexpr int CG_BN = 256;                 // queries per tile           = UMMA N
constexpr int CG_BK = 64;                  // halfs per k-block: 128 bytes = one 128B-swizzle row
constexpr int CG_UK = 16;                  // UMMA K for 16-bit operands
constexpr int CG_STAGES = 4;
constexpr uint32_t CG_A_BYTES = CG_BM * CG_BK * 2;      // 16 KB
constexpr uint32_t CG_B_BYTES = CG_BN * CG_BK * 2;      // 32 KB
constexpr uint32_t CG_STAGE_BYTES = CG_A_BYTES + CG_B_BYTES;
constexpr int CG_EPI_WARPS = 16;           // 4 per TMEM lane quarter (= per SM sub-partition): latency hiding for the read-out
constexpr int CG_THREADS = (CG_EPI_WARPS + 2) * 32;   // epilogue warps + producer warp + MMA warp
constexpr int CG_CHUNKS_PER_WARP = (256 / 32) / (CG_EPI_WARPS / 4);   // 32-column chunks of a tile per epilogue warp
constexpr int CG_TMEM_COLS = 512;          // two 256-column fp32 accumulators
constexpr int CG_LCAP = 128;               // staged survivors per epilogue warp before a flush to the global lists
constexpr size_t CG_SMEM = 1024 /*align slack*/ + (size_t)CG_STAGES * CG_STAGE_BYTES + COARSE_MAX_BATCH * 4 + 256
                         + CG_EPI_WARPS * (size_t)CG_LCAP * (8 + 2);

// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=B=f16 (0), K-major both, N>>3 at [17,23),
// M>>4 at [24,29) -- see CgCfg::IDESC (M = 128 for one CTA, 256 for a CTA pair)

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor) of a K-major, 128B-swizzled operand tile whose rows
// are 128 bytes: start address >> 4, LBO (unused for swizzled K-major) = 1, SBO = 8 rows * 128 B = 1024 B >> 4,
// version = 1 (Blackwell) at [46,48), layout SWIZZLE_128B = 2 at [61,64)
__device__ __forceinline__ u64 smem_desc_sw128(uint32_t smem_addr) {
    return (u64)((smem_addr >> 4) & 0x3FFFu) | ((u64)1 << 16) | ((u64)(1024 >> 4) << 32) | ((u64)1 << 46) | ((u64)2 << 61);
}

// Pairwise mode: a row range [row_first, row_first + rows) whose every row index is <= every query's row index lies on
// or below the diagonal.
__device__ __forceinline__ bool rows_below_diagonal(int64_t tri_q0, int64_t row_first, int rows, int qb) {
    return tri_q0 >= 0 && row_first + (rows - 1) <= tri_q0 + (int64_t)qb * CG_BN;
}

// cluster helpers (NCTA == 2)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {        // the same smem offset in the cluster's CTA 0
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr)); return r;
}
__device__ __forceinline__ void mbarrier_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {          // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_d, u64 adesc, u64 bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// NCTA == 1: one CTA per SM, tile 128 rows x 256 queries, operands A (16 KB) + B (32 KB) per k-block, 4 stages.
// NCTA == 2: CTA pairs (cluster of 2, cta_group::2): the pair's tile is 256 rows x 256 queries; each CTA stages its own
//            128 rows of A and HALF of the queries (16 KB + 16 KB per k-block, 6 stages), the leader's elected lane issues
//            M=256 MMAs that read both CTAs' shared memory, and each CTA reads its own 128 accumulator lanes out of its
//            own TMEM.  Per SM that halves the B-operand traffic into and out of shared memory -- the resource the
//            single-CTA kernel saturates (TMA writes 96 B/clk + MMA reads 96 B/clk against 128 B/clk).
template <int NCTA> struct CgCfg {
    static constexpr int STAGES = NCTA == 2 ? 6 : 4;
    static constexpr uint32_t B_BYTES = CG_B_BYTES / NCTA;
    static constexpr uint32_t STAGE_BYTES = CG_A_BYTES + B_BYTES;
    static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(CG_BN >> 3) << 17) | ((uint32_t)((CG_BM * NCTA) >> 4) << 24);
};
static_assert(CgCfg<1>::STAGES * CgCfg<1>::STAGE_BYTES == CgCfg<2>::STAGES * CgCfg<2>::STAGE_BYTES, "both variants use the same ring bytes");

template <int MODE, int NCTA>   // MODE 0 = filter (main pass), 1 = sample dump
__global__ void __launch_bounds__(CG_THREADS, 1)
coarse_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   int64_t n, int n_tiles, int tile_stride, int n_qblocks, int n_kblocks,
                   const float* __restrict__ thr,          // [b_pad] thresholds (unscaled coarse scores), MODE 0
                   u64* __restrict__ cand, int32_t* __restrict__ cand_cnt, int cand_cap,   // MODE 0
                   __half* __restrict__ sample, int64_t sample_rows,                       // MODE 1: [b_pad][sample_rows]
                   int64_t tri_q0)   // >= 0: pairwise mode, query column q is matrix row tri_q0 + q; keep only row > that
{
    using Cfg = CgCfg<NCTA>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ unsigned char cg_smem_raw[];
    const uint32_t raw = cvta_smem(cg_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                       // 128B swizzle atoms need 1024-byte alignment
    unsigned char* gen_base = cg_smem_raw + (base - raw);
    float* thr_s = reinterpret_cast<float*>(gen_base + (size_t)STAGES * Cfg::STAGE_BYTES);
    u64* bars = reinterpret_cast<u64*>(thr_s + COARSE_MAX_BATCH);
    const uint32_t bar0 = cvta_smem(bars);
    auto full_bar  = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (uint32_t)(STAGES + s); };
    auto tfull_bar = [&](int b) { return bar0 + 8u * (uint32_t)(2 * STAGES + b); };
    auto tempty_bar = [&](int b) { return bar0 + 8u * (uint32_t)(2 * STAGES + 2 + b); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    uint32_t* lcnt_all = tmem_slot + 4;                                 // per-warp staging counters (16 x 4 B)
    u64* lkey_all = reinterpret_cast<u64*>(reinterpret_cast<unsigned char*>(bars) + 256);
    uint16_t* lq_all = reinterpret_cast<uint16_t*>(lkey_all + CG_EPI_WARPS * CG_LCAP);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float kScale = 16777216.0f;                                   // 2^24 = (2^12)^2, the operands' scaling
    const int rank = NCTA == 2 ? (int)cluster_ctarank() : 0;            // CTA within the pair
    const bool leader = rank == 0;

    if (MODE == 0) {
        const int nq = n_qblocks * CG_BN;
        for (int i = threadIdx.x; i < nq; i += CG_THREADS) thr_s[i] = thr[i] * kScale;
    }
    if (threadIdx.x < CG_EPI_WARPS) lcnt_all[threadIdx.x] = 0;
    if (warp == CG_EPI_WARPS && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbarrier_init(full_bar(s), 1); mbarrier_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbarrier_init(tfull_bar(b), 1); mbarrier_init(tempty_bar(b), CG_EPI_WARPS * NCTA); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == CG_EPI_WARPS + 1) {
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(cvta_smem(tmem_slot)), "n"(CG_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(cvta_smem(tmem_slot)), "n"(CG_TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work units: (group of NCTA consecutive row tiles, query block); adjacent units share rows, so M16 is read from
    // HBM once and from L2 for the other query blocks
    const int n_groups = (n_tiles + NCTA - 1) / NCTA;
    const int64_t total = (int64_t)n_groups * n_qblocks;
    const int64_t unit0 = blockIdx.x / NCTA, unit_step = gridDim.x / NCTA;
    auto my_row0 = [&](int grp) { return (int64_t)(grp * NCTA + rank) * tile_stride * CG_BM; };
    auto skip_unit = [&](int grp, int qb) {       // pairwise mode: the unit's rows all lie on or below the diagonal
        return tile_stride == 1 && rows_below_diagonal(tri_q0, (int64_t)grp * NCTA * CG_BM, CG_BM * NCTA, qb);
    };

    if (warp == CG_EPI_WARPS) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0; uint32_t phase = 0;
            const uint32_t full0 = NCTA == 2 ? mapa_rank0(full_bar(0)) : full_bar(0);     // the LEADER's full barriers
            for (int64_t w = unit0; w < total; w += unit_step) {
                const int grp = (int)(w / n_qblocks), qb = (int)(w - (int64_t)grp * n_qblocks);
                if (skip_unit(grp, qb)) continue;
                const int row0 = (int)my_row0(grp);                   // < 2^31 rows per engine (checked on the host)
                const int q0 = qb * CG_BN + rank * (CG_BN / NCTA);    // this CTA's share of the query block
                for (int kb = 0; kb < n_kblocks; ++kb) {
                    mbarrier_wait(empty_bar(s), phase ^ 1u);
                    const uint32_t sa = base + (uint32_t)s * Cfg::STAGE_BYTES;
                    if (NCTA == 2) {
                        // both CTAs' bytes complete on the leader's barrier; only the leader arrives (with the pair's total)
                        if (leader) mbarrier_arrive_expect_tx(full_bar(s), 2 * Cfg::STAGE_BYTES);
                        tma_load_2d_2sm(sa, &tmA, kb * CG_BK, row0, full0 + 8u * (uint32_t)s);
                        tma_load_2d_2sm(sa + CG_A_BYTES, &tmB, kb * CG_BK, q0, full0 + 8u * (uint32_t)s);
                    } else {
                        mbarrier_arrive_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
                        tma_load_2d(sa, &tmA, kb * CG_BK, row0, full_bar(s));
                        tma_load_2d(sa + CG_A_BYTES, &tmB, kb * CG_BK, q0, full_bar(s));
                    }
                    if (++s == STAGES) { s = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == CG_EPI_WARPS + 1) {
        // ------------------------------------------------------------------ MMA issuer (one elected lane of the leader CTA)
        if (lane == 0 && leader) {
            int s = 0; uint32_t phase = 0;
            int64_t it = 0;                                              // units actually processed
            for (int64_t w = unit0; w < total; w += unit_step) {
                const int grp = (int)(w / n_qblocks), qb = (int)(w - (int64_t)grp * n_qblocks);
                if (skip_unit(grp, qb)) continue;
                const int buf = (int)(it & 1);
                const uint32_t use = (uint32_t)(it >> 1);
                ++it;
                mbarrier_wait(tempty_bar(buf), (use & 1u) ^ 1u);        // every epilogue warp (of both CTAs) has drained it
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * CG_BN;
                for (int kb = 0; kb < n_kblocks; ++kb) {
                    mbarrier_wait(full_bar(s), phase);
                    tc_fence_after();
                    const uint32_t sa = base + (uint32_t)s * Cfg::STAGE_BYTES;
                    const u64 adesc = smem_desc_sw128(sa), bdesc = smem_desc_sw128(sa + CG_A_BYTES);
#pragma unroll
                    for (int k = 0; k < CG_BK / CG_UK; ++k) {        // +32 bytes (>>4 = 2) per K=16 step inside the swizzle row
                        if (NCTA == 2) tc_mma_f16_2sm(tmem_d, adesc + (u64)(2 * k), bdesc + (u64)(2 * k), Cfg::IDESC, (uint32_t)((kb | k) != 0));
                        else           tc_mma_f16(tmem_d, adesc + (u64)(2 * k), bdesc + (u64)(2 * k), Cfg::IDESC, (uint32_t)((kb | k) != 0));
                    }
                    if (NCTA == 2) tc_commit_2sm(empty_bar(s)); else tc_commit(empty_bar(s));   // ring slot free once read
                    if (++s == STAGES) { s = 0; phase ^= 1u; }
                }
                if (NCTA == 2) tc_commit_2sm(tfull_bar(buf)); else tc_commit(tfull_bar(buf));   // accumulator complete
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps: warp w reads TMEM lanes
        // 32*(w%4).. (the hardware ties a warp to the lane quarter of its sub-partition) and the column chunks
        // (w/4)*CG_CHUNKS_PER_WARP ...  Survivors are staged in a per-warp shared-memory list (shared-memory atomics: tens of cycles) and flushed
        // to the per-query global lists by all 32 lanes at once, so the ~700-cycle global atomic round trip is
        // paid once per ~128 survivors instead of once per survivor in the middle of the TMEM read-out.
        u64* lkey = lkey_all + warp * CG_LCAP;
        uint16_t* lq = lq_all + warp * CG_LCAP;
        uint32_t* lcnt = lcnt_all + warp;
        const uint32_t tempty0 = NCTA == 2 ? mapa_rank0(tempty_bar(0)) : tempty_bar(0);     // the LEADER's barriers
        auto append_global = [&](int q, u64 key) {
            const int slot = atomicAdd(&cand_cnt[q], 1);
            if (slot < cand_cap) cand[(size_t)q * cand_cap + slot] = key;
        };
        // staged entry: raw accumulator bits << 32 | row; the key (ordered score) is made at flush time, off the hot path
        auto flush = [&]() {
            __syncwarp();
            const int cnt = min((int)*reinterpret_cast<volatile uint32_t*>(lcnt), CG_LCAP);
            for (int i0 = 0; i0 < cnt; i0 += 128) {                    // 4 independent atomics in flight per lane
                int qs[4], slots[4]; u64 keys[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 32 + lane;
                    qs[u] = -1;
                    if (i < cnt) {
                        qs[u] = lq[i];
                        const u64 e = lkey[i];
                        keys[u] = make_key(__uint_as_float((uint32_t)(e >> 32)) * (1.0f / kScale), (uint32_t)e);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) slots[u] = qs[u] >= 0 ? atomicAdd(&cand_cnt[qs[u]], 1) : cand_cap;
#pragma unroll
                for (int u = 0; u < 4; ++u) if (slots[u] < cand_cap) cand[(size_t)qs[u] * cand_cap + slots[u]] = keys[u];
            }
            __syncwarp();
            if (lane == 0) *lcnt = 0;
            __syncwarp();
        };
        int64_t it = 0;
        for (int64_t w = unit0; w < total; w += unit_step) {
            const int grp = (int)(w / n_qblocks), qb = (int)(w - (int64_t)grp * n_qblocks);
            if (skip_unit(grp, qb)) continue;
            const int buf = (int)(it & 1);
            const uint32_t use = (uint32_t)(it >> 1);
            ++it;
            mbarrier_wait(tfull_bar(buf), use & 1u);
            tc_fence_after();
            const int lq4 = warp & 3;                                  // TMEM lane quarter
            const int64_t row = my_row0(grp) + lq4 * 32 + lane;
            const bool row_ok = row < n;
            const uint32_t taddr0 = tmem_base + ((uint32_t)(lq4 * 32) << 16) + (uint32_t)buf * CG_BN;
            const int c_begin = (warp >> 2) * CG_CHUNKS_PER_WARP;
#pragma unroll 1
            for (int c = c_begin; c < c_begin + CG_CHUNKS_PER_WARP; ++c) {
                uint32_t v[32];
                tc_ld_32x32(taddr0 + (uint32_t)c * 32, v);
                tc_wait_ld();
                if (c == c_begin + CG_CHUNKS_PER_WARP - 1) {           // this warp's share is read: one of the arrivals
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (NCTA == 2) mbarrier_arrive_cluster(tempty0 + 8u * (uint32_t)buf);
                        else           mbarrier_arrive(tempty_bar(buf));
                    }
                }
                const int q0 = qb * CG_BN + c * 32;
                if (MODE == 0) {
                    const float4* th4 = reinterpret_cast<const float4*>(thr_s + q0);
                    uint32_t mask = 0;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 th = th4[j4];
                        mask |= (__uint_as_float(v[4 * j4 + 0]) >= th.x ? 1u : 0u) << (4 * j4 + 0);
                        mask |= (__uint_as_float(v[4 * j4 + 1]) >= th.y ? 1u : 0u) << (4 * j4 + 1);
                        mask |= (__uint_as_float(v[4 * j4 + 2]) >= th.z ? 1u : 0u) << (4 * j4 + 2);
                        mask |= (__uint_as_float(v[4 * j4 + 3]) >= th.w ? 1u : 0u) << (4 * j4 + 3);
                    }
                    if (!row_ok) mask = 0;
                    if (tri_q0 >= 0) {                                 // upper triangle: column j allowed iff row > tri_q0 + q0 + j
                        const int64_t lim = row - tri_q0 - q0;
                        mask &= lim >= 32 ? 0xffffffffu : (lim <= 0 ? 0u : ((1u << (int)lim) - 1u));
                    }
                    // Survivors are rare (well under 1 % of the scores): walk the columns in which ANY lane has one
                    // (warp-uniform loop, one uniform indexed branch per column to pick the register) instead of
                    // testing all 32 columns in every lane.
                    const uint32_t wor = __reduce_or_sync(0xffffffffu, mask);
                    if (wor) {
                        const uint32_t tot = __reduce_add_sync(0xffffffffu, (uint32_t)__popc(mask));
                        if (*reinterpret_cast<volatile uint32_t*>(lcnt) + tot > (uint32_t)CG_LCAP) flush();   // uniform: make room
                        if (tot <= (uint32_t)CG_LCAP) {
                            uint32_t pos = mask ? atomicAdd(lcnt, (uint32_t)__popc(mask)) : 0u;       // shared-memory atomic
                            uint32_t todo = wor;
                            while (todo) {
                                const int j = __ffs(todo) - 1;
                                todo &= todo - 1;
                                uint32_t val;
                                switch (j) {
#define SVSB_PICK(J) case J: val = v[J]; break;
                                    SVSB_PICK(0) SVSB_PICK(1) SVSB_PICK(2) SVSB_PICK(3) SVSB_PICK(4) SVSB_PICK(5) SVSB_PICK(6) SVSB_PICK(7)
                                    SVSB_PICK(8) SVSB_PICK(9) SVSB_PICK(10) SVSB_PICK(11) SVSB_PICK(12) SVSB_PICK(13) SVSB_PICK(14) SVSB_PICK(15)
                                    SVSB_PICK(16) SVSB_PICK(17) SVSB_PICK(18) SVSB_PICK(19) SVSB_PICK(20) SVSB_PICK(21) SVSB_PICK(22) SVSB_PICK(23)
                                    SVSB_PICK(24) SVSB_PICK(25) SVSB_PICK(26) SVSB_PICK(27) SVSB_PICK(28) SVSB_PICK(29) SVSB_PICK(30)
                                    default: val = v[31]; break;
#undef SVSB_PICK
                                }
                                if ((mask >> j) & 1u) {
                                    lkey[pos] = ((u64)val << 32) | (u64)(uint32_t)row;
                                    lq[pos] = (uint16_t)(q0 + j);
                                    ++pos;
                                }
                            }
                        } else {
                            // more survivors in one chunk than the staging list holds (thresholds at -inf, massive ties):
                            // straight to the global lists -- slow, rare, correct
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (mask & (1u << j))
                                    append_global(q0 + j, make_key(__uint_as_float(v[j]) * (1.0f / kScale), (uint32_t)row));
                        }
                        __syncwarp();
                    }
                } else {
                    const int64_t srow = (int64_t)(grp * NCTA + rank) * CG_BM + lq4 * 32 + lane;   // position inside the sample
                    if (srow < sample_rows) {
                        const int64_t lim = tri_q0 >= 0 ? row - tri_q0 - q0 : 32;   // pairwise: only columns j < lim count
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            sample[(size_t)(q0 + j) * sample_rows + srow] =      // rounded down: stays a lower bound
                                __float2half_rd((row_ok && j < lim) ? __uint_as_float(v[j]) * (1.0f / kScale) : __int_as_float(0xff800000));
                    }
                }
            }
        }
        if (MODE == 0) flush();
    }

    tc_fence_before();
    if (NCTA == 2) cluster_sync_all(); else __syncthreads();        // the peer must not leave while its smem / barriers are in use
    if (warp == CG_EPI_WARPS + 1) {
        tc_fence_after();
        if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(CG_TMEM_COLS) : "memory");
        else           asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(CG_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
// M16[r][c] = fp16(M[r][c] * scale) for c < ld, 0 for ld <= c < ld16.   ld % 4 == 0, ld16 % 8 == 0.
__global__ void __launch_bounds__(256)
rows_to_f16_kernel(const float* __restrict__ M, int64_t n, int ld, __half* __restrict__ M16, int ld16, float scale)
{
    const int q16 = ld16 / 4;                                   // 4-element groups per output row
    const int64_t total = n * (int64_t)q16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / q16; const int c = (int)(i - r * q16) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < ld) v = ldg_stream(reinterpret_cast<const float4*>(M + r * (int64_t)ld + c));
        const __half2 lo = __floats2half2_rn(v.x * scale, v.y * scale);
        const __half2 hi = __floats2half2_rn(v.z * scale, v.w * scale);
        uint2 out;
        out.x = *reinterpret_cast<const uint32_t*>(&lo);
        out.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(M16 + r * (int64_t)ld16 + c) = out;
    }
}

cudaError_t launch_rows_to_f16(cudaStream_t st, int device, const float* M, int64_t n, int ld, void* M16, int ld16)
{
    if (n <= 0) return cudaSuccess;
    const int64_t total = n * (int64_t)(ld16 / 4);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count(device) * 16;
    if (blocks > cap) blocks = cap;
    rows_to_f16_kernel<<<(unsigned)blocks, 256, 0, st>>>(M, n, ld, reinterpret_cast<__half*>(M16), ld16, COARSE_OPERAND_SCALE);
    count_launch();
    return cudaGetLastError();
}

// One warp per query row of Q16 (b_pad rows; rows >= b are zero).  Also: eps[q] = the bound on |coarse - exact| for
// this query (see batch.cu), thr[q] = +inf for padding queries, flags[q] = 1 when the coarse path cannot be trusted
// for this query (non-finite or huge norm) and it must take the exact single-query path.
__global__ void __launch_bounds__(256)
queries_to_f16_kernel(const float* __restrict__ Q, int b, int b_pad, int d, int ldq, __half* __restrict__ Q16, int ld16,
                      float scale, float eps_coef, float max_row_norm, float* __restrict__ eps, float* __restrict__ thr,
                      int32_t* __restrict__ flags, int32_t* __restrict__ cand_cnt)
{
    const int lane = threadIdx.x & 31;
    const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= b_pad) return;
    float ss = 0.f;
    for (int c = lane; c < ld16; c += 32) {
        float v = 0.f;
        if (q < b && c < d) v = Q[(int64_t)q * ldq + c];
        ss = fmaf(v, v, ss);
        Q16[(int64_t)q * ld16 + c] = __float2half_rn(v * scale);
    }
    ss = warp_sum(ss);
    if (lane == 0) {
        const float nrm = sqrtf(ss) * 1.000001f;
        const bool bad = !(nrm <= 8.0f);                         // also catches NaN / inf
        eps[q] = eps_coef * nrm * max_row_norm + 1e-8f;
        thr[q] = __int_as_float(0x7f800000);                     // +inf until the sample pass sets it (padding keeps it)
        flags[q] = (q < b && bad) ? 1 : 0;
        if (cand_cnt) cand_cnt[q] = 0;                           // the filter pass counts from zero (saves a memset node per batch)
    }
}

cudaError_t launch_queries_to_f16(cudaStream_t st, const float* Q, int b, int b_pad, int d, int ldq, void* Q16, int ld16,
                                  float eps_coef, float max_row_norm, float* eps, float* thr, int32_t* flags, int32_t* cand_cnt)
{
    if (b_pad <= 0) return cudaSuccess;
    const int blocks = (b_pad * 32 + 255) / 256;
    queries_to_f16_kernel<<<blocks, 256, 0, st>>>(Q, b, b_pad, d, ldq, reinterpret_cast<__half*>(Q16), ld16,
                                                  COARSE_OPERAND_SCALE, eps_coef, max_row_norm, eps, thr, flags, cand_cnt);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    // initialised once, thread-safely (C++11 magic static): several host threads reach this concurrently on a multi-device
    // engine, one worker per device
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiledFn>(p);
        (void)cudaGetLastError();
        return (EncodeTiledFn) nullptr;
    }();
    return fn;
}

// 2-D fp16 tensor [rows][ld16] (K contiguous), box = 64 halfs x box_rows, 128-byte swizzle, out-of-bounds = 0.
static cudaError_t make_map(CUtensorMap* map, const void* ptr, int64_t rows, int ld16, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)ld16, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld16 * 2};
    const cuuint32_t box[2] = {(cuuint32_t)CG_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int MODE, int NCTA>
static cudaError_t launch_coarse_inst(cudaStream_t st, int device, unsigned grid, const CUtensorMap& tmA, const CUtensorMap& tmB,
                                      int64_t n, int n_tiles, int tile_stride, int n_qblocks, int n_kblocks, const float* thr, u64* cand,
                                      int32_t* cand_cnt, int cand_cap, __half* sample, int64_t sample_rows, int64_t tri_q0)
{
    auto kern = coarse_gemm_kernel<MODE, NCTA>;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CG_SMEM);
        if (e != cudaSuccess) return e;
        attr_set[device] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(CG_THREADS); cfg.dynamicSmemBytes = CG_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = NCTA == 2 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, n, n_tiles, tile_stride, n_qblocks, n_kblocks, thr, cand, cand_cnt, cand_cap,
                              sample, sample_rows, tri_q0);
}

// SVSB_COARSE_CTAS = 1 | 2 selects the single-CTA or the CTA-pair kernel (default decided by measurement, profiles/).
static int coarse_ctas() {
    static int v = 0;
    if (v == 0) { const char* s = getenv("SVSB_COARSE_CTAS"); v = (s && atoi(s) == 1) ? 1 : (s && atoi(s) == 2) ? 2 : COARSE_DEFAULT_CTAS; }
    return v;
}

cudaError_t launch_coarse_gemm(cudaStream_t st, int device, int mode, const void* M16, int64_t n, const void* Q16, int b_pad,
                               int ld16, int n_tiles, int tile_stride, const float* thr, u64* cand, int32_t* cand_cnt,
                               int cand_cap, void* sample, int64_t sample_rows, int64_t q_rows, int64_t tri_q0)
{
    if (n <= 0 || b_pad <= 0 || b_pad % CG_BN || b_pad > COARSE_MAX_BATCH || ld16 % 8 || n_tiles <= 0) return cudaErrorInvalidValue;
    if (n > 0x7fffff00ll) return cudaErrorInvalidValue;
    const int ncta = coarse_ctas();
    CUtensorMap tmA, tmB;
    cudaError_t e = make_map(&tmA, M16, n, ld16, CG_BM);
    if (e != cudaSuccess) return e;
    e = make_map(&tmB, Q16, q_rows > 0 ? q_rows : b_pad, ld16, CG_BN / ncta);   // rows beyond the extent read as zero
    if (e != cudaSuccess) return e;
    const int n_qblocks = b_pad / CG_BN;
    const int n_kblocks = (ld16 + CG_BK - 1) / CG_BK;
    const int64_t total = (int64_t)((n_tiles + ncta - 1) / ncta) * n_qblocks;
    int64_t units = sm_count(device) / ncta;
    if (units > total) units = total;
    const unsigned grid = (unsigned)(units * ncta);
    __half* smp = reinterpret_cast<__half*>(sample);
#define SVSB_CG(MODEV, NCTAV) launch_coarse_inst<MODEV, NCTAV>(st, device, grid, tmA, tmB, n, n_tiles, tile_stride, n_qblocks, n_kblocks, \
                                                               thr, cand, cand_cnt, cand_cap, smp, sample_rows, tri_q0)
    if (ncta == 2) e = mode ? SVSB_CG(1, 2) : SVSB_CG(0, 2);
    else           e = mode ? SVSB_CG(1, 1) : SVSB_CG(0, 1);
#undef SVSB_CG
    if (e != cudaSuccess) return e;
    count_launch();
    return cudaGetLastError();
}

// Load the coarse kernels onto the current device now (see preload_peer_kernels).
cudaError_t preload_coarse_kernels()
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, coarse_gemm_kernel<0, 1>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, coarse_gemm_kernel<1, 1>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, coarse_gemm_kernel<0, 2>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, coarse_gemm_kernel<1, 2>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, queries_to_f16_kernel);
    return e;
}

}  // namespace svsb
