// K2 (second half) -- thresholds for the coarse pass and the exact refine of its candidates.
//
// Why the result is exact.  For one query let e(r) = coarse(r) - exact(r) and |e(r)| <= eps
// for every row r (the bound computed on the host, see engine.cu: two fp16 roundings per product, a generous
// bound on the tensor core's fp32 accumulation and on the exact kernel's own fp32 rounding).  Let tau~ be the
// kk-th largest COARSE score.  kk rows have coarse >= tau~, hence exact >= tau~ - eps, so the kk-th largest EXACT
// score s_k >= tau~ - eps.  Any row in the exact top-kk has exact >= s_k, hence coarse >= tau~ - 2 eps.  The refine
// kernel therefore re-scores every candidate with coarse >= tau~ - 2 eps in fp32 -- with the very summation
// order of the similarity kernel (gemv.cu, TMA variant), so scores are bit-identical to svsb_query's -- and the
// exact top-kk of those is the exact top-kk of all rows, under the same total order (score desc, row asc).
// The candidate list holds every row with coarse >= tau~ - 2 eps whenever the filter threshold T <= tau~ - 2 eps.
// Guaranteed mode: T = (the kk-th largest coarse score of a SAMPLE of the rows) - 2 eps; the sample's kk-th largest
// can only be lower than tau~.  Statistical mode (default): T = (the m-th largest of the sample) - 2 eps with
// m < kk chosen on the host so that "the sample holds m of the overall top kk" has probability ~1e-9 for a random
// sample -- several times fewer candidates -- and the refine kernel CHECKS T <= tau~ - 2 eps per query: tau~ is
// exact as soon as the list holds kk entries (every row >= T is in it), so the check is exact; a query that fails
// it (flag 16) is redone by the caller, never answered from an incomplete list.
#include "select_common.cuh"

#include <cuda_fp16.h>

namespace svsb {

constexpr int RF_THREADS = 512;
constexpr int RF_BINS = KTH_BINS;
constexpr int RF_SMALL = KTH_SMALL;

// ---------------------------------------------------------------------------------------------
// thresholds from the sample: thr[q] = kk-th largest coarse score among the sampled rows
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RF_THREADS)
sample_threshold_kernel(const __half* __restrict__ sample, int64_t sample_rows, int rank, const float* __restrict__ eps,
                        float* __restrict__ thr)
{
    __shared__ uint32_t hist[RF_BINS];
    __shared__ uint32_t scratch[72];
    __shared__ uint32_t small[RF_SMALL];
    const int q = blockIdx.x;
    const __half* s = sample + (size_t)q * sample_rows;          // coarse scores rounded DOWN to fp16: still a lower bound
    const uint32_t o = block_kth_largest_o32([&](int64_t i) { return f32_to_ordered(__half2float(s[i])); }, sample_rows, rank,
                                             hist, scratch, small);
    // The filter must let through every row with coarse >= tau~ - 2 eps; the sample's kk-th largest is <= tau~
    // (rank < kk: with overwhelming probability, verified by refine_kernel).
    if (threadIdx.x == 0) thr[q] = ordered_to_f32(o) - 2.0f * eps[q];
}

cudaError_t launch_sample_threshold(cudaStream_t st, const void* sample, int64_t sample_rows, int b, int rank, const float* eps,
                                    float* thr)
{
    if (b <= 0) return cudaSuccess;
    if (rank < 1 || rank > sample_rows) return cudaErrorInvalidValue;
    sample_threshold_kernel<<<b, RF_THREADS, 0, st>>>(reinterpret_cast<const __half*>(sample), sample_rows, rank, eps, thr);
    count_launch();
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// refine
// ---------------------------------------------------------------------------------------------
constexpr int RF_CACHE = 8192;                  // candidates whose coarse scores are kept in shared memory across passes
struct RefineSmem {
    u64 keys[REFINE_SURVIVOR_CAP];
    uint32_t rows[REFINE_SURVIVOR_CAP];
    uint32_t cached[RF_CACHE];
    uint32_t hist[RF_BINS];
    uint32_t scratch[72];
    uint32_t small[RF_SMALL];
    uint32_t counter;
    // float4 q[ld / 4] follows
};

__global__ void __launch_bounds__(RF_THREADS, 2)
refine_kernel(const float* __restrict__ M, int64_t n, int d4, const int64_t* __restrict__ ids, int64_t row0,
              const float* __restrict__ Q, int ldq, int k, const u64* __restrict__ cand, const int32_t* __restrict__ cand_cnt,
              int cand_cap, const float* __restrict__ eps, const float* __restrict__ thr, int32_t* __restrict__ flags,
              RefineOut out, int32_t* __restrict__ stats)
{
    extern __shared__ __align__(16) unsigned char rf_smem_raw[];
    RefineSmem& sm = *reinterpret_cast<RefineSmem*>(rf_smem_raw);
    float4* sq = reinterpret_cast<float4*>(rf_smem_raw + ((sizeof(RefineSmem) + 15) & ~(size_t)15));
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = RF_THREADS / 32;
    const int kk = (int)min((int64_t)k, n);
    int32_t* out_count = out.counts + (int64_t)q * out.count_stride;
    if (tid == 0) { *out_count = 0; if (stats) stats[q] = 0; }
    if (flags[q] != 0) return;                                       // already routed to the exact path
    const int total = cand_cnt[q];
    if (total > cand_cap || total < kk) {
        if (tid == 0) flags[q] = total > cand_cap ? 2 : 4;
        return;
    }
    const u64* cq = cand + (size_t)q * cand_cap;
    for (int c = tid; c < d4; c += RF_THREADS) sq[c] = reinterpret_cast<const float4*>(Q + (size_t)q * ldq)[c];
    if (tid == 0) sm.counter = 0;
    // one trip to global memory for the coarse scores (the key's high word IS the ordered score); the select passes
    // below then run out of shared memory
    {
        const int lim = min(total, RF_CACHE);
        for (int i0 = tid; i0 < lim; i0 += RF_THREADS * 4) {
            u64 kv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * RF_THREADS; kv[u] = i < lim ? cq[i] : 0ull; }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * RF_THREADS; if (i < lim) sm.cached[i] = (uint32_t)(kv[u] >> 32); }
        }
    }
    __syncthreads();
    auto score_o = [&](int64_t i) { return i < RF_CACHE ? sm.cached[i] : (uint32_t)(cq[i] >> 32); };

    // tau~: the kk-th largest coarse score among the candidates
    const uint32_t tau_o = block_kth_largest_o32(score_o, total, kk, sm.hist, sm.scratch, sm.small);
    const float cutoff = ordered_to_f32(tau_o) - 2.0f * eps[q];
    // The list holds exactly the rows with coarse >= thr[q] and at least kk of them, so tau_o IS the kk-th largest
    // coarse score of all rows; rows with coarse in [cutoff, thr[q]) would be missing from the list.
    if (thr && !(cutoff >= thr[q])) { if (tid == 0) flags[q] = REFINE_FLAG_THRESHOLD_HIGH; return; }

    for (int i = tid; i < total; i += RF_THREADS) {
        if (ordered_to_f32(score_o(i)) >= cutoff) {
            const uint32_t p = atomicAdd(&sm.counter, 1u);
            if (p < (uint32_t)REFINE_SURVIVOR_CAP) sm.rows[p] = key_row(cq[i]);
        }
    }
    __syncthreads();
    const int C = (int)sm.counter;
    if (C > REFINE_SURVIVOR_CAP) { if (tid == 0) flags[q] = 8; return; }
    if (tid == 0 && stats) stats[q] = C;

    // exact fp32 re-score, one warp per survivor.  Summation order == gemv_tma_kernel: lane l takes the float4
    // chunks l, l+32, ...; even chunks accumulate into a0, odd ones into a1; then the same combine + xor tree.
    // Two survivors per warp iteration: twice the loads in flight, each row's own summation order untouched.
    const float4* M4 = reinterpret_cast<const float4*>(M);
    for (int i = warp; i < C; i += 2 * nwarps) {
        const int i2 = i + nwarps;
        const bool two = i2 < C;
        const uint32_t rowa = sm.rows[i], rowb = sm.rows[two ? i2 : i];
        const float4* pa = M4 + (int64_t)rowa * d4;                  // candidate keys carry LOCAL rows
        const float4* pb = M4 + (int64_t)rowb * d4;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
        int c = lane;
        for (; c + 96 < d4; c += 128) {
            const float4 m0 = ldg_stream(pa + c), m1 = ldg_stream(pa + c + 32), m2 = ldg_stream(pa + c + 64), m3 = ldg_stream(pa + c + 96);
            const float4 n0 = ldg_stream(pb + c), n1 = ldg_stream(pb + c + 32), n2 = ldg_stream(pb + c + 64), n3 = ldg_stream(pb + c + 96);
            const float4 q0 = sq[c], q1 = sq[c + 32], q2 = sq[c + 64], q3 = sq[c + 96];
            fma4(a0, m0, q0); fma4(a1, m1, q1); fma4(a0, m2, q2); fma4(a1, m3, q3);
            fma4(b0, n0, q0); fma4(b1, n1, q1); fma4(b0, n2, q2); fma4(b1, n3, q3);
        }
        for (; c + 32 < d4; c += 64) {
            const float4 m0 = ldg_stream(pa + c), m1 = ldg_stream(pa + c + 32);
            const float4 n0 = ldg_stream(pb + c), n1 = ldg_stream(pb + c + 32);
            const float4 q0 = sq[c], q1 = sq[c + 32];
            fma4(a0, m0, q0); fma4(a1, m1, q1);
            fma4(b0, n0, q0); fma4(b1, n1, q1);
        }
        if (c < d4) {
            const float4 m0 = ldg_stream(pa + c), n0 = ldg_stream(pb + c);
            const float4 q0 = sq[c];
            fma4(a0, m0, q0); fma4(b0, n0, q0);
        }
        const float sa = warp_sum(((a0.x + a1.x) + (a0.y + a1.y)) + ((a0.z + a1.z) + (a0.w + a1.w)));
        const float sb = warp_sum(((b0.x + b1.x) + (b0.y + b1.y)) + ((b0.z + b1.z) + (b0.w + b1.w)));
        if (lane == 0) { sm.keys[i] = make_key(sa, rowa); if (two) sm.keys[i2] = make_key(sb, rowb); }
    }
    __syncthreads();

    int np2 = 1; while (np2 < C) np2 <<= 1;
    for (int i = C + tid; i < np2; i += RF_THREADS) sm.keys[i] = 0ull;
    __syncthreads();
    block_bitonic_desc<false>(sm.keys, nullptr, np2);

    for (int i = tid; i < kk; i += RF_THREADS) {
        const u64 key = sm.keys[i];
        const uint32_t row = key_row(key);
        const int64_t grow = row0 + (int64_t)row;                    // global row (row0 = first row of this shard)
        if (out.scores) out.scores[(int64_t)q * out.stride + i] = key_score(key);
        if (out.keys) out.keys[(int64_t)q * out.stride + i] = (key & 0xffffffff00000000ull) | (u64)(uint32_t)(~(uint32_t)grow);
        out.ids[(int64_t)q * out.stride + i] = ids ? ids[row] : grow;
    }
    if (tid == 0) *out_count = kk;
}

cudaError_t launch_refine(cudaStream_t st, const float* M, int64_t n, int ld, const int64_t* ids, int64_t row0,
                          const float* Q, int b, int ldq, int k, const u64* cand, const int32_t* cand_cnt, int cand_cap,
                          const float* eps, const float* thr, int32_t* flags, RefineOut out, int32_t* stats)
{
    if (b <= 0) return cudaSuccess;
    if (!out.ids || !out.counts) return cudaErrorInvalidValue;
    if (k < 1 || (ld & 3) || ldq < ld) return cudaErrorInvalidValue;
    const int64_t kk = k < n ? k : n;
    if (kk > REFINE_SURVIVOR_CAP) return cudaErrorInvalidValue;
    const size_t smem = ((sizeof(RefineSmem) + 15) & ~(size_t)15) + (size_t)ld * 4;
    static bool attr_set[64] = {false};
    int dev = 0; cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    refine_kernel<<<b, RF_THREADS, smem, st>>>(M, n, ld / 4, ids, row0, Q, ldq, k, cand, cand_cnt, cand_cap, eps, thr, flags, out, stats);
    count_launch();
    return cudaGetLastError();
}

}  // namespace svsb
