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

// The same order statistic, the fast way (rank <= 256 and a sample of >= 2048 rows: every configuration the engine
// plans).  "Thread maxima + exact refilter" -- the trick of the selection kernel (select.cu) in miniature: with tau0 =
// the rank-th largest of the 256 per-thread maxima, at least `rank` sample values are >= tau0, so the rank-th largest
// overall is among the values >= tau0 -- typically rank plus a handful, because the top of a sample is spread over
// the threads.  Those are collected into shared memory and rank-counted (duplicates counted, as np.partition would).
// Two passes over the sample (the second one out of L1/L2), four barriers, no histogram.  Threads without sample values
// (fewer than 256 vectors) rank last with key 0; the launcher requires rank <= number of vectors.  Works on the raw fp16 bits
// (order-preserving 16-bit keys); the result converts to the very float the generic kernel returns.
constexpr int ST_THREADS = 256;
constexpr int ST_LIST = 1024;

__device__ __forceinline__ uint32_t h16_ordered(uint32_t h) { return (h & 0x8000u) ? (~h & 0xffffu) : (h | 0x8000u); }
__device__ __forceinline__ float ordered16_to_f32(uint32_t o) {
    const uint32_t h = (o & 0x8000u) ? (o & 0x7fffu) : (~o & 0xffffu);
    return __half2float(__ushort_as_half((unsigned short)h));
}

// Publication of a multi-CTA kernel's peer stores (BatchPush): every CTA, after its own stores: barrier, system-scope
// fence (cumulative over the barrier), device-scope ticket; the CTA that draws the last ticket resets the counter, fences
// again and release-stores the sequence number into every rank's flag.  All threads of the CTA call it.
__device__ __forceinline__ void batch_publish(const BatchPush& push) {
    if (push.world == 0) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int total = gridDim.x * gridDim.y;
        if (atomicAdd(push.done, 1u) == total - 1u) {
            *push.done = 0u;
            __threadfence_system();
            for (int p = 0; p < push.world; ++p) st_release_sys(push.flag[p], push.seq);
        }
    }
}

// TOP == false: thr[q] = (rank-th largest of the sample) - 2 eps[q].
// TOP == true : top[q][0..SAMPLE_TOPX) = the SAMPLE_TOPX largest sample values, descending (the sharded path exchanges
//               them and takes the order statistic of the UNION of all ranks' samples, union_threshold_kernel).
template <bool TOP>
__global__ void __launch_bounds__(ST_THREADS)
sample_order_kernel(const __half* __restrict__ sample, int64_t sample_rows, int rank, const float* __restrict__ eps,
                    float* __restrict__ thr, float* __restrict__ top, const __grid_constant__ BatchPush push)
{
    __shared__ uint32_t mx[ST_THREADS];
    __shared__ uint32_t list[ST_LIST];
    __shared__ uint32_t hist[RF_BINS];
    __shared__ uint32_t scratch[72];
    __shared__ uint32_t small[RF_SMALL];
    __shared__ uint32_t s_tau0, s_count, s_ans;
    const int q = blockIdx.x, tid = threadIdx.x;
    const __half* s = sample + (size_t)q * sample_rows;
    const uint4* s4 = reinterpret_cast<const uint4*>(s);          // sample_rows is a multiple of 128: rows are 256-byte aligned
    const int nvec = (int)(sample_rows >> 3);
    const int want = TOP ? SAMPLE_TOPX : rank;
    uint32_t m = 0;
    for (int v = tid; v < nvec; v += ST_THREADS) {
        const uint4 x = s4[v];
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) { m = max(m, h16_ordered(w[u] & 0xffffu)); m = max(m, h16_ordered(w[u] >> 16)); }
    }
    if (tid == 0) s_count = 0;
    // keys made distinct by the thread number: the want-th largest of the thread maxima under (value desc, thread asc)
    const uint32_t tau0 = block_select_unique<uint32_t>((m << 8) | (uint32_t)(ST_THREADS - 1 - tid), want, mx, &s_tau0) >> 8;
    for (int v = tid; v < nvec; v += ST_THREADS) {
        const uint4 x = s4[v];
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t o = h16_ordered((u & 1) ? (w[u >> 1] >> 16) : (w[u >> 1] & 0xffffu));
            if (o >= tau0) { const uint32_t p = atomicAdd(&s_count, 1u); if (p < (uint32_t)ST_LIST) list[p] = o; }
        }
    }
    __syncthreads();
    const int c = (int)s_count;
    if (c > ST_LIST) {                                            // a thousand ties at the top of the sample: generic select
        if (TOP) {                                                // +inf thresholds keep nothing: the query goes to the exact path
            if (tid < SAMPLE_TOPX) {
                if (push.world == 0) top[(size_t)q * SAMPLE_TOPX + tid] = __int_as_float(0x7f800000);
                for (int p = 0; p < push.world; ++p) static_cast<float*>(push.dst[p])[(size_t)q * SAMPLE_TOPX + tid] = __int_as_float(0x7f800000);
            }
            batch_publish(push);
            return;
        }
        const uint32_t o = block_kth_largest_o32([&](int64_t i) { return f32_to_ordered(__half2float(s[i])); }, sample_rows, rank,
                                                 hist, scratch, small);
        if (tid == 0) thr[q] = ordered_to_f32(o) - 2.0f * eps[q];
        return;
    }
    for (int i = tid; i < c; i += ST_THREADS) {
        const uint32_t mine = list[i];
        if (TOP) {
            int rk = 0;
            for (int j = 0; j < c; ++j) { const uint32_t o = list[j]; rk += (o > mine || (o == mine && j < i)) ? 1 : 0; }
            if (rk < SAMPLE_TOPX) {
                const float v = ordered16_to_f32(mine);
                if (push.world == 0) top[(size_t)q * SAMPLE_TOPX + rk] = v;
                for (int p = 0; p < push.world; ++p) static_cast<float*>(push.dst[p])[(size_t)q * SAMPLE_TOPX + rk] = v;
            }
        } else {
            uint32_t gt = 0, ge = 0;
            for (int j = 0; j < c; ++j) { const uint32_t o = list[j]; gt += o > mine ? 1u : 0u; ge += o >= mine ? 1u : 0u; }
            if (gt < (uint32_t)rank && ge >= (uint32_t)rank) s_ans = mine;     // equal values write the same word
        }
    }
    if (TOP) { batch_publish(push); return; }
    __syncthreads();
    if (tid == 0) thr[q] = ordered16_to_f32(s_ans) - 2.0f * eps[q];
}

static bool sample_fast_ok(int64_t sample_rows, int rank) {
    static const bool off = [] { const char* v = getenv("SVSB_SAMPLE_GENERIC"); return v && atoi(v) != 0; }();
    // every thread whose maximum takes part in the ranking must own at least one vector of 8 sample values
    return !off && rank >= 1 && rank <= ST_THREADS && (sample_rows & 127) == 0 && rank <= (sample_rows >> 3);
}

cudaError_t launch_sample_threshold(cudaStream_t st, const void* sample, int64_t sample_rows, int b, int rank, const float* eps,
                                    float* thr)
{
    if (b <= 0) return cudaSuccess;
    if (rank < 1 || rank > sample_rows) return cudaErrorInvalidValue;
    if (sample_fast_ok(sample_rows, rank))
        sample_order_kernel<false><<<b, ST_THREADS, 0, st>>>(reinterpret_cast<const __half*>(sample), sample_rows, rank, eps, thr, nullptr, BatchPush());
    else
        sample_threshold_kernel<<<b, RF_THREADS, 0, st>>>(reinterpret_cast<const __half*>(sample), sample_rows, rank, eps, thr);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_sample_top(cudaStream_t st, const void* sample, int64_t sample_rows, int b, float* top, const BatchPush* push)
{
    if (b <= 0) return cudaSuccess;
    if (!sample_fast_ok(sample_rows, SAMPLE_TOPX)) return cudaErrorInvalidValue;
    if (push ? (push->world < 1 || push->world > XCHG_MAX_RANKS || !push->done) : !top) return cudaErrorInvalidValue;
    sample_order_kernel<true><<<b, ST_THREADS, 0, st>>>(reinterpret_cast<const __half*>(sample), sample_rows, SAMPLE_TOPX, nullptr, nullptr, top,
                                                        push ? *push : BatchPush());
    count_launch();
    return cudaGetLastError();
}

// thr[q] = (rank-th largest of the union of the ranks' top lists tops[r][q][0..SAMPLE_TOPX)) - 2 eps[q]; rank <=
// SAMPLE_TOPX, so the union's rank-th largest is the rank-th largest of ALL ranks' sample values.  One warp per query.
__global__ void __launch_bounds__(256)
union_threshold_kernel(const float* __restrict__ tops, int64_t list_stride, int world, int b, int rank, const float* __restrict__ eps,
                       float* __restrict__ thr)
{
    // One warp per query; lane p holds entry p of every rank's list (ordered 32-bit keys).  The rank-th largest of the
    // union is built bit by bit from the top: x grows by a bit whenever at least `rank` values are still >= x -- one
    // ballot per list and bit, no shared memory, no dependent loads.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * 8 + warp;
    if (q >= b) return;
    uint32_t v[XCHG_MAX_RANKS];
#pragma unroll
    for (int l = 0; l < XCHG_MAX_RANKS; ++l)
        v[l] = l < world ? f32_to_ordered(tops[(size_t)l * list_stride + (size_t)q * SAMPLE_TOPX + lane]) : 0u;
    uint32_t x = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t t = x | (1u << bit);
        int c = 0;
#pragma unroll
        for (int l = 0; l < XCHG_MAX_RANKS; ++l) c += __popc(__ballot_sync(0xffffffffu, l < world && v[l] >= t));
        if (c >= rank) x = t;
    }
    if (lane == 0) thr[q] = ordered_to_f32(x) - 2.0f * eps[q];
}

cudaError_t launch_union_threshold(cudaStream_t st, const float* tops, int64_t list_stride, int world, int b, int rank, const float* eps,
                                   float* thr)
{
    if (b <= 0) return cudaSuccess;
    if (world < 1 || world > XCHG_MAX_RANKS || rank < 1 || rank > SAMPLE_TOPX || list_stride < (int64_t)b * SAMPLE_TOPX) return cudaErrorInvalidValue;
    union_threshold_kernel<<<(b + 7) / 8, 256, 0, st>>>(tops, list_stride, world, b, rank, eps, thr);
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

// ---------------------------------------------------------------------------------------------
// refine, split by what bounds each phase (the default; the one-CTA-per-query kernel above is kept for A/B measurement,
// SVSB_REFINE_FUSED=1).  The fused kernel holds 94 KB of shared memory and 512 threads through its select, filter and
// sort phases, during which its SM's share of the HBM idles (ncu: 43 % DRAM throughput).  Split:
//   refine_select_kernel   one small CTA per query: tau~, cutoff, verification, survivors' rows -> global list
//   rescore_kernel         grid-wide, nothing but the exact fp32 dot products: (query, slice of its survivors) per CTA,
//                          3-12 KB of shared memory (the query), four CTAs per SM, two rows in flight per warp
//   refine_sort_kernel     one small CTA per query: sort the survivors' exact keys, emit the top kk
// mode bit 0 (REFINE_PARTIAL, sharded path with a GLOBAL threshold): fewer than kk local candidates is the normal case
//   (the shard holds only its part of the global top kk): every candidate is re-scored, the record's count is what
//   there is, and the count word's high half carries ver = #{candidates with coarse >= thr + 2 eps}; the merge adds
//   ver over the ranks: sum >= kk proves thr <= (global kk-th largest coarse score) - 2 eps, i.e. no shard's list
//   misses a row of the global top kk.
// mode bit 1 (REFINE_DEFER): a query the coarse path cannot answer (flags[q] != 0) gets count -1 in its record
//   instead of waiting for the host to read the flags; the merge propagates it and the caller redoes such queries.
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_CACHE = 2560;                  // candidates whose coarse scores stay in shared memory (more: re-read through L2)
constexpr int RS_LIST = 1024;                   // candidates at or above tau0 the fast select ranks in shared memory (<= RF_BINS)

__global__ void __launch_bounds__(RS_THREADS)
refine_select_kernel(int64_t n, int k, const u64* __restrict__ cand, const int32_t* __restrict__ cand_cnt, int cand_cap,
                     const float* __restrict__ eps, const float* __restrict__ thr, int32_t* __restrict__ flags, int mode,
                     RefineScratch sc, int32_t* __restrict__ stats)
{
    __shared__ uint32_t cached[RS_CACHE];
    __shared__ uint32_t hist[RF_BINS];
    __shared__ uint32_t scratch[72];
    __shared__ uint32_t small[RF_SMALL];
    __shared__ u64 sel_sorted[RS_THREADS];
    __shared__ u64 sel_bcast;
    __shared__ uint32_t counter, vcount, lcount;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int kk = (int)min((int64_t)k, n);
    const bool partial = (mode & REFINE_PARTIAL) != 0;
    // independent loads first: one round trip instead of a chain of them
    const int flag_in = flags[q];
    const int total = cand_cnt[q];
    const float eps_q = eps[q];
    const float thr_q = thr ? thr[q] : 0.f;
    if (tid == 0) { sc.cnt[q] = 0; sc.ver[q] = 0; if (stats) stats[q] = 0; counter = 0; vcount = 0; }
    if (flag_in != 0) return;                                        // already routed to the exact path
    if (total > cand_cap || (!partial && total < kk)) {
        if (tid == 0) flags[q] = total > cand_cap ? 2 : 4;
        return;
    }
    const u64* cq = cand + (size_t)q * cand_cap;
    {
        const int lim = min(total, RS_CACHE);
        for (int i0 = tid; i0 < lim; i0 += RS_THREADS * 4) {
            u64 kv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * RS_THREADS; kv[u] = i < lim ? cq[i] : 0ull; }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * RS_THREADS; if (i < lim) cached[i] = (uint32_t)(kv[u] >> 32); }
        }
    }
    __syncthreads();
    auto score_o = [&](int64_t i) { return i < RS_CACHE ? cached[i] : (uint32_t)(cq[i] >> 32); };
    float cutoff = __int_as_float(0xff800000);                       // -inf: keep every candidate (partial, total < kk)
    if (total >= kk) {
        uint32_t tau_o = 0;
        bool have = false;
        if (kk <= RS_THREADS / 2) {
            // "thread maxima + exact refilter" (see sample_order_kernel): tau0 = the kk-th largest of the per-thread maxima is a
            // lower bound of tau~ with about kk .. 1.3 kk candidates at or above it; those are rank-counted in shared memory
            uint32_t m = 0;
            for (int i = tid; i < total; i += RS_THREADS) m = max(m, score_o(i));
            if (tid == 0) lcount = 0;
            const uint32_t tau0 = (uint32_t)(block_select_unique<u64>(((u64)m << 32) | (u64)(RS_THREADS - 1 - tid), kk, sel_sorted, &sel_bcast) >> 32);
            for (int i = tid; i < total; i += RS_THREADS) {
                const uint32_t o = score_o(i);
                if (o >= tau0) { const uint32_t p = atomicAdd(&lcount, 1u); if (p < (uint32_t)RS_LIST) hist[p] = o; }
            }
            __syncthreads();
            const int c = (int)lcount;
            if (c <= RS_LIST) {
                for (int i = tid; i < c; i += RS_THREADS) {
                    const uint32_t mine = hist[i];
                    uint32_t gt = 0, ge = 0;
                    for (int j = 0; j < c; ++j) { const uint32_t o = hist[j]; gt += o > mine ? 1u : 0u; ge += o >= mine ? 1u : 0u; }
                    if (gt < (uint32_t)kk && ge >= (uint32_t)kk) scratch[7] = mine;   // equal values write the same word
                }
                __syncthreads();
                tau_o = scratch[7];
                have = true;
            }
            __syncthreads();
        }
        if (!have) tau_o = block_kth_largest_o32(score_o, total, kk, hist, scratch, small);
        cutoff = ordered_to_f32(tau_o) - 2.0f * eps_q;
        // The list holds exactly the rows with coarse >= thr[q] and at least kk of them, so tau_o IS the kk-th largest
        // coarse score of all (local) rows; rows with coarse in [cutoff, thr[q]) would be missing from the list.
        if (!partial && thr && !(cutoff >= thr_q)) { if (tid == 0) flags[q] = REFINE_FLAG_THRESHOLD_HIGH; return; }
    }
    const float vthr = partial ? __fadd_ru(thr_q, __fmul_ru(2.0f, eps_q)) : 0.f;
    uint32_t* rows = sc.rows + (size_t)q * REFINE_SURVIVOR_CAP;
    for (int i0 = 0; i0 < total; i0 += RS_THREADS) {
        const int i = i0 + tid;
        const float s = i < total ? ordered_to_f32(score_o(i)) : 0.f;
        if (i < total && s >= cutoff) {
            const uint32_t p = atomicAdd(&counter, 1u);
            if (p < (uint32_t)REFINE_SURVIVOR_CAP) rows[p] = key_row(cq[i]);
        }
        if (partial) {
            const uint32_t v = __ballot_sync(0xffffffffu, i < total && s >= vthr);
            if (lane == 0 && v) atomicAdd(&vcount, (uint32_t)__popc(v));
        }
    }
    __syncthreads();
    const int C = (int)counter;
    if (C > REFINE_SURVIVOR_CAP) { if (tid == 0) flags[q] = 8; return; }
    if (tid == 0) { sc.cnt[q] = C; sc.ver[q] = (int32_t)vcount; if (stats) stats[q] = C; }
}

// Exact fp32 re-score, one warp per survivor.  Summation order == gemv_tma_kernel (gemv.cu): lane l takes the float4
// chunks l, l+32, ...; even chunks accumulate into a0, odd ones into a1; then the same combine + xor tree.  Two
// survivors per warp iteration: twice the loads in flight, each row's own summation order untouched.
__global__ void __launch_bounds__(RS_THREADS)
rescore_kernel(const float* __restrict__ M, int d4, const float* __restrict__ Q, int ldq, RefineScratch sc)
{
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    float4* sq = reinterpret_cast<float4*>(rs_smem_raw);
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = RS_THREADS / 32;
    const int C = sc.cnt[q];                                          // 0 for flagged queries
    const int first = blockIdx.y * nwarps, stride = gridDim.y * nwarps;
    if (first >= C) return;
    for (int c = tid; c < d4; c += RS_THREADS) sq[c] = reinterpret_cast<const float4*>(Q + (size_t)q * ldq)[c];
    __syncthreads();
    const uint32_t* rows = sc.rows + (size_t)q * REFINE_SURVIVOR_CAP;
    u64* keys = sc.keys + (size_t)q * REFINE_SURVIVOR_CAP;
    const float4* M4 = reinterpret_cast<const float4*>(M);
    for (int i = first + warp; i < C; i += 2 * stride) {
        const int i2 = i + stride;
        const bool two = i2 < C;
        const uint32_t rowa = rows[i], rowb = rows[two ? i2 : i];
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
        if (lane == 0) { keys[i] = make_key(sa, rowa); if (two) keys[i2] = make_key(sb, rowb); }
    }
}

__global__ void __launch_bounds__(RS_THREADS)
refine_sort_kernel(int64_t n, int k, const int64_t* __restrict__ ids, int64_t row0, const int32_t* __restrict__ flags, int mode,
                   RefineScratch sc, RefineOut out, const __grid_constant__ BatchPush push)
{
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    u64* sk = reinterpret_cast<u64*>(rs_smem_raw);                   // np2 (<= REFINE_SURVIVOR_CAP) keys, or 2 x 256 for the rank sort
    const int q = blockIdx.x, tid = threadIdx.x;
    const int kk = (int)min((int64_t)k, n);
    int32_t* out_count = push.world ? nullptr : out.counts + (int64_t)q * out.count_stride;
    if (flags[q] != 0) {
        if (tid == 0) {
            const int32_t c0 = (mode & REFINE_DEFER) ? -1 : 0;
            if (out_count) { out_count[0] = c0; if (mode & REFINE_PARTIAL) out_count[1] = 0; }
            for (int p = 0; p < push.world; ++p) {
                int32_t* cw = reinterpret_cast<int32_t*>(static_cast<u64*>(push.dst[p]) + (int64_t)q * out.stride + 2 * out.cap);
                cw[0] = c0; cw[1] = 0;
            }
        }
        batch_publish(push);
        return;
    }
    const int C = sc.cnt[q];
    const u64* keys = sc.keys + (size_t)q * REFINE_SURVIVOR_CAP;
    const u64* sorted;
    if (C <= RANK_SORT_MAX) {
        if (tid < C) sk[tid] = keys[tid];
        __syncthreads();
        block_rank_sort_desc(sk, sk + RANK_SORT_MAX, C);
        sorted = sk + RANK_SORT_MAX;
    } else {
        int np2 = 1; while (np2 < C) np2 <<= 1;
        for (int i = tid; i < np2; i += RS_THREADS) sk[i] = i < C ? keys[i] : 0ull;
        __syncthreads();
        block_bitonic_desc<false>(sk, nullptr, np2);
        sorted = sk;
    }
    const int full = min(kk, C);
    const int cnt = (out.cap > 0 && (mode & REFINE_PARTIAL)) ? min(full, out.cap) : full;
    for (int i = tid; i < cnt; i += RS_THREADS) {
        const u64 key = sorted[i];
        const uint32_t row = key_row(key);
        const int64_t grow = row0 + (int64_t)row;                    // global row (row0 = first row of this shard)
        const u64 gkey = (key & 0xffffffff00000000ull) | (u64)(uint32_t)(~(uint32_t)grow);
        const int64_t id = ids ? ids[row] : grow;
        if (push.world == 0) {
            if (out.scores) out.scores[(int64_t)q * out.stride + i] = key_score(key);
            if (out.keys) out.keys[(int64_t)q * out.stride + i] = gkey;
            out.ids[(int64_t)q * out.stride + i] = id;
        }
        for (int p = 0; p < push.world; ++p) {                       // the record, straight into every rank's window
            u64* r = static_cast<u64*>(push.dst[p]) + (int64_t)q * out.stride;
            r[i] = gkey; r[out.cap + i] = (u64)id;
        }
    }
    if (tid == 0) {
        const int32_t c0 = cnt | (cnt < full ? REFINE_COUNT_TRUNCATED : 0), c1 = sc.ver[q];
        if (out_count) { out_count[0] = c0; if (mode & REFINE_PARTIAL) out_count[1] = c1; }
        for (int p = 0; p < push.world; ++p) {
            int32_t* cw = reinterpret_cast<int32_t*>(static_cast<u64*>(push.dst[p]) + (int64_t)q * out.stride + 2 * out.cap);
            cw[0] = c0; cw[1] = c1;
        }
    }
    batch_publish(push);
}

// ---------------------------------------------------------------------------------------------
// refine, lean and fused (the default for k <= 409): select -> exact re-score -> sort -> emit / push in ONE kernel, one CTA
// of 256 threads per query with 21 KB static + 12-15 KB dynamic shared memory; MINB = 5 (48 registers) puts five CTAs on
// an SM when the query and the lists are small, MINB = 3 otherwise, so that the latency-bound phases of some overlap the
// HBM-bound re-score of the others.  At a 125 k-row shard it replaces the split's three launches (15 + 60 + 11 us, each
// with dependent global round trips at its head) by one of 74-86 us (profiles/r02_launches_c3_virtual_n8.md); on one GPU
// with the bench's data it is no faster than round 1's kernel.
// Shared memory for keys[cap_s] / rows[cap_s] is sized by the host from k; a query with more survivors keeps the rest in
// global lists (RefineScratch) and sorts there -- slower, exact all the same; more than REFINE_SURVIVOR_CAP: flag 8.
// ---------------------------------------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(RS_THREADS, MINB)
refine_lean_kernel(const float* __restrict__ M, int64_t n, int d4, const int64_t* __restrict__ ids, int64_t row0,
                   const float* __restrict__ Q, int ldq, int k, const u64* __restrict__ cand, const int32_t* __restrict__ cand_cnt,
                   int cand_cap, const float* __restrict__ eps, const float* __restrict__ thr, int32_t* __restrict__ flags, int mode,
                   int cap_s, RefineScratch spill, RefineOut out, int32_t* __restrict__ stats, const __grid_constant__ BatchPush push)
{
    extern __shared__ __align__(16) unsigned char rl_smem_raw[];
    __shared__ uint32_t cached[RS_CACHE];
    __shared__ uint32_t hist[RF_BINS];
    __shared__ uint32_t scratch[72];
    __shared__ uint32_t small[RF_SMALL];
    __shared__ u64 sel_sorted[RS_THREADS];
    __shared__ u64 sel_bcast;
    __shared__ uint32_t counter, vcount, lcount;
    float4* sq = reinterpret_cast<float4*>(rl_smem_raw);
    u64* skeys = reinterpret_cast<u64*>(rl_smem_raw + (size_t)d4 * 16);
    uint32_t* srows = reinterpret_cast<uint32_t*>(skeys + cap_s);
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = RS_THREADS / 32;
    const int kk = (int)min((int64_t)k, n);
    const bool partial = (mode & REFINE_PARTIAL) != 0;
    const int flag_in = flags[q];                                   // independent loads first: one round trip
    const int total = cand_cnt[q];
    const float eps_q = eps[q];
    const float thr_q = thr ? thr[q] : 0.f;
    int32_t* out_count = push.world ? nullptr : out.counts + (int64_t)q * out.count_stride;
    int flag_out = flag_in;
    if (flag_in == 0 && (total > cand_cap || (!partial && total < kk))) flag_out = total > cand_cap ? 2 : 4;
    if (tid == 0) { counter = 0; vcount = 0; lcount = 0; if (stats) stats[q] = 0; }
    const u64* cq = cand + (size_t)q * cand_cap;
    // survivors beyond cap_s (a score distribution packed more tightly than the host's sizing rule assumes) spill to global lists
    uint32_t* grows = spill.rows ? spill.rows + (size_t)q * REFINE_SURVIVOR_CAP : nullptr;
    u64* gkeys = spill.keys ? spill.keys + (size_t)q * REFINE_SURVIVOR_CAP : nullptr;
    const int cap_all = grows && gkeys ? REFINE_SURVIVOR_CAP : cap_s;
    int C = 0, ver = 0;
    if (flag_out == 0) {
        for (int c = tid; c < d4; c += RS_THREADS) sq[c] = reinterpret_cast<const float4*>(Q + (size_t)q * ldq)[c];
        const int lim = min(total, RS_CACHE);
        for (int i0 = tid; i0 < lim; i0 += RS_THREADS * 4) {
            u64 kv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * RS_THREADS; kv[u] = i < lim ? cq[i] : 0ull; }
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * RS_THREADS; if (i < lim) cached[i] = (uint32_t)(kv[u] >> 32); }
        }
        __syncthreads();
        auto score_o = [&](int64_t i) { return i < RS_CACHE ? cached[i] : (uint32_t)(cq[i] >> 32); };
        float cutoff = __int_as_float(0xff800000);                  // -inf: keep every candidate (partial, total < kk)
        if (total >= kk) {
            uint32_t tau_o = 0;
            bool have = false;
            if (kk <= RS_THREADS / 2) {                             // thread maxima + exact refilter (see sample_order_kernel)
                uint32_t m = 0;
                for (int i = tid; i < total; i += RS_THREADS) m = max(m, score_o(i));
                const uint32_t tau0 = (uint32_t)(block_select_unique<u64>(((u64)m << 32) | (u64)(RS_THREADS - 1 - tid), kk, sel_sorted, &sel_bcast) >> 32);
                for (int i = tid; i < total; i += RS_THREADS) {
                    const uint32_t o = score_o(i);
                    if (o >= tau0) { const uint32_t p = atomicAdd(&lcount, 1u); if (p < (uint32_t)RS_LIST) hist[p] = o; }
                }
                __syncthreads();
                const int c = (int)lcount;
                if (c <= RS_LIST) {
                    for (int i = tid; i < c; i += RS_THREADS) {
                        const uint32_t mine = hist[i];
                        uint32_t gt = 0, ge = 0;
                        for (int j = 0; j < c; ++j) { const uint32_t o = hist[j]; gt += o > mine ? 1u : 0u; ge += o >= mine ? 1u : 0u; }
                        if (gt < (uint32_t)kk && ge >= (uint32_t)kk) scratch[7] = mine;   // equal values write the same word
                    }
                    __syncthreads();
                    tau_o = scratch[7];
                    have = true;
                }
                __syncthreads();
            }
            if (!have) tau_o = block_kth_largest_o32(score_o, total, kk, hist, scratch, small);
            cutoff = ordered_to_f32(tau_o) - 2.0f * eps_q;
            // the list holds exactly the rows with coarse >= thr and at least kk of them, so tau_o IS the kk-th largest coarse
            // score of all (local) rows; rows with coarse in [cutoff, thr) would be missing from the list
            if (!partial && thr && !(cutoff >= thr_q)) flag_out = REFINE_FLAG_THRESHOLD_HIGH;
        }
        if (flag_out == 0) {
            const float vthr = partial ? __fadd_ru(thr_q, __fmul_ru(2.0f, eps_q)) : 0.f;
            for (int i0 = 0; i0 < total; i0 += RS_THREADS) {
                const int i = i0 + tid;
                const float s = i < total ? ordered_to_f32(score_o(i)) : 0.f;
                if (i < total && s >= cutoff) {
                    const uint32_t p = atomicAdd(&counter, 1u);
                    if (p < (uint32_t)cap_s) srows[p] = key_row(cq[i]);
                    else if (p < (uint32_t)cap_all) grows[p] = key_row(cq[i]);
                }
                if (partial) {
                    const uint32_t v = __ballot_sync(0xffffffffu, i < total && s >= vthr);
                    if (lane == 0 && v) atomicAdd(&vcount, (uint32_t)__popc(v));
                }
            }
            __syncthreads();
            C = (int)counter; ver = (int)vcount;
            if (C > cap_all) flag_out = 8;
        }
    }
    if (flag_out != 0) {                                            // the exact path answers this query
        if (tid == 0) {
            if (flag_out != flag_in) flags[q] = flag_out;
            const int32_t c0 = (mode & REFINE_DEFER) ? -1 : 0;
            if (out_count) { out_count[0] = c0; if (partial) out_count[1] = 0; }
            for (int p = 0; p < push.world; ++p) {
                int32_t* cw = reinterpret_cast<int32_t*>(static_cast<u64*>(push.dst[p]) + (int64_t)q * out.stride + 2 * out.cap);
                cw[0] = c0; cw[1] = 0;
            }
        }
        batch_publish(push);
        return;
    }
    if (tid == 0 && stats) stats[q] = C;

    // exact fp32 re-score, one warp per survivor, two survivors per warp iteration (summation order == gemv_tma_kernel)
    const float4* M4 = reinterpret_cast<const float4*>(M);
    for (int i = warp; i < C; i += 2 * nwarps) {
        const int i2 = i + nwarps;
        const bool two = i2 < C;
        const int ib = two ? i2 : i;
        const uint32_t rowa = i < cap_s ? srows[i] : grows[i], rowb = ib < cap_s ? srows[ib] : grows[ib];
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
        if (lane == 0) {
            const u64 ka = make_key(sa, rowa), kb = make_key(sb, rowb);
            if (i < cap_s) skeys[i] = ka; else gkeys[i] = ka;
            if (two) { if (i2 < cap_s) skeys[i2] = kb; else gkeys[i2] = kb; }
        }
    }
    __syncthreads();

    const u64* sorted;
    if (C <= RANK_SORT_MAX) {                                       // sel_sorted doubles as the rank sort's destination
        block_rank_sort_desc(skeys, sel_sorted, C);
        sorted = sel_sorted;
    } else if (C <= cap_s) {
        int np2 = 1; while (np2 < C) np2 <<= 1;                     // np2 <= cap_s (a power of two)
        for (int i = C + tid; i < np2; i += RS_THREADS) skeys[i] = 0ull;
        __syncthreads();
        block_bitonic_desc<false>(skeys, nullptr, np2);
        sorted = skeys;
    } else {                                                        // spilled: sort the whole list where it is, in global memory
        int np2 = 1; while (np2 < C) np2 <<= 1;                     // <= REFINE_SURVIVOR_CAP
        for (int i = tid; i < np2; i += RS_THREADS) { if (i < cap_s) gkeys[i] = skeys[i]; else if (i >= C) gkeys[i] = 0ull; }
        __syncthreads();
        block_bitonic_desc<false>(gkeys, nullptr, np2);
        sorted = gkeys;
    }
    const int full = min(kk, C);
    const int cnt = (out.cap > 0 && partial) ? min(full, out.cap) : full;
    for (int i = tid; i < cnt; i += RS_THREADS) {
        const u64 key = sorted[i];
        const uint32_t row = key_row(key);
        const int64_t grow = row0 + (int64_t)row;                    // global row (row0 = first row of this shard)
        const u64 gkey = (key & 0xffffffff00000000ull) | (u64)(uint32_t)(~(uint32_t)grow);
        const int64_t id = ids ? ids[row] : grow;
        if (push.world == 0) {
            if (out.scores) out.scores[(int64_t)q * out.stride + i] = key_score(key);
            if (out.keys) out.keys[(int64_t)q * out.stride + i] = gkey;
            out.ids[(int64_t)q * out.stride + i] = id;
        }
        for (int p = 0; p < push.world; ++p) {                       // the record, straight into every rank's window
            u64* r = static_cast<u64*>(push.dst[p]) + (int64_t)q * out.stride;
            r[i] = gkey; r[out.cap + i] = (u64)id;
        }
    }
    if (tid == 0) {
        const int32_t c0 = cnt | (cnt < full ? REFINE_COUNT_TRUNCATED : 0);
        if (out_count) { out_count[0] = c0; if (partial) out_count[1] = ver; }
        for (int p = 0; p < push.world; ++p) {
            int32_t* cw = reinterpret_cast<int32_t*>(static_cast<u64*>(push.dst[p]) + (int64_t)q * out.stride + 2 * out.cap);
            cw[0] = c0; cw[1] = ver;
        }
    }
    batch_publish(push);
}

__global__ void __launch_bounds__(32)
wait_flags_kernel(const u64* flags, int world, u64 seq, u64 timeout_ns, int* status)
{
    const int tid = threadIdx.x;
    if (tid < world) {
        unsigned ns = 32;
        u64 t0 = 0;
        while (ld_acquire_sys(flags + tid) < seq) {
            __nanosleep(ns);
            if (ns < 1024) ns <<= 1;
            else {
                u64 now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > timeout_ns) { atomicOr(status, 1); break; }
            }
        }
    }
}

cudaError_t launch_wait_flags(cudaStream_t st, const u64* flags, int world, unsigned long long seq, unsigned long long timeout_ns, int* status)
{
    if (!flags || !status || world < 1 || world > 32) return cudaErrorInvalidValue;
    wait_flags_kernel<<<1, 32, 0, st>>>(flags, world, seq, timeout_ns, status);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_refine(cudaStream_t st, const float* M, int64_t n, int ld, const int64_t* ids, int64_t row0,
                          const float* Q, int b, int ldq, int k, const u64* cand, const int32_t* cand_cnt, int cand_cap,
                          const float* eps, const float* thr, int32_t* flags, RefineOut out, int32_t* stats,
                          const RefineScratch* scratch, int mode, const BatchPush* push)
{
    if (b <= 0) return cudaSuccess;
    if (push) {
        if (push->world < 1 || push->world > XCHG_MAX_RANKS || !push->done || out.cap < 1 || out.stride < 2 * (int64_t)out.cap + 1 ||
            !(mode & REFINE_PARTIAL)) return cudaErrorInvalidValue;
    } else if (!out.ids || !out.counts) return cudaErrorInvalidValue;
    if (k < 1 || (ld & 3) || ldq < ld) return cudaErrorInvalidValue;
    const int64_t kk = k < n ? k : n;
    if (kk > REFINE_SURVIVOR_CAP) return cudaErrorInvalidValue;
    // SVSB_REFINE = lean (default) | split | fused ; SVSB_REFINE_FUSED=1 is the old spelling of "fused"
    static const int variant = [] {
        const char* v = getenv("SVSB_REFINE");
        const char* f = getenv("SVSB_REFINE_FUSED");
        if (f && atoi(f) != 0) return 2;
        if (v && !strcmp(v, "split")) return 1;
        if (v && !strcmp(v, "fused")) return 2;
        return 0;
    }();
    // survivors the lean kernel holds in shared memory: a power of two >= max(1024, 2.5 kk) (tau~ - 2 eps keeps kk plus a margin
    // band whose population depends on how tightly the scores are packed; what does not fit spills to global lists).
    // Beyond 1024 (kk > 409) its CTAs get so fat that two share an SM; the split form -- re-score grid-wide from global
    // lists -- is then the faster one (1M x 3072, k = 1000, 256 queries: 2.16 ms against 2.92 ms).
    int cap_s = 1024;
    while (cap_s < 4096 && cap_s < (int)((5 * kk + 1) / 2)) cap_s <<= 1;
    const bool have_scratch = scratch && scratch->rows && scratch->keys && scratch->cnt && scratch->ver;
    const bool fused = variant == 2 && mode == 0 && !push;
    const bool split = have_scratch && (variant == 1 || (variant == 0 && cap_s > 1024));
    if ((mode & REFINE_PARTIAL) && !thr) return cudaErrorInvalidValue;
    static bool attr_set[64] = {false};
    int dev = 0; cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(refine_lean_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(refine_lean_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024);
        // the CTAs are small: ask for the largest shared-memory carve-out so that registers, not the default split, bound residency
        if (e == cudaSuccess) e = cudaFuncSetAttribute(refine_lean_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(refine_lean_kernel<5>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    if (!fused && !split) {
        const size_t smem = (size_t)ld * 4 + (size_t)cap_s * 12;
        if (smem > 180 * 1024) return cudaErrorInvalidValue;
        // five resident CTAs per SM (48 registers) when the query and the lists leave room, else three (72 registers)
        static const int minb_env = [] { const char* v = getenv("SVSB_REFINE_MINB"); return v ? atoi(v) : 0; }();
        const bool five = minb_env ? minb_env >= 5 : smem <= 20 * 1024;
        const RefineScratch sp = have_scratch ? *scratch : RefineScratch{nullptr, nullptr, nullptr, nullptr};
        if (five)
            refine_lean_kernel<5><<<b, RS_THREADS, smem, st>>>(M, n, ld / 4, ids, row0, Q, ldq, k, cand, cand_cnt, cand_cap, eps, thr, flags, mode,
                                                               cap_s, sp, out, stats, push ? *push : BatchPush());
        else
            refine_lean_kernel<3><<<b, RS_THREADS, smem, st>>>(M, n, ld / 4, ids, row0, Q, ldq, k, cand, cand_cnt, cand_cap, eps, thr, flags, mode,
                                                               cap_s, sp, out, stats, push ? *push : BatchPush());
        count_launch();
        return cudaGetLastError();
    }
    if (fused) {
        const size_t smem = ((sizeof(RefineSmem) + 15) & ~(size_t)15) + (size_t)ld * 4;
        if (smem > 200 * 1024) return cudaErrorInvalidValue;
        refine_kernel<<<b, RF_THREADS, smem, st>>>(M, n, ld / 4, ids, row0, Q, ldq, k, cand, cand_cnt, cand_cap, eps, thr, flags, out, stats);
        count_launch();
        return cudaGetLastError();
    }
    if ((size_t)ld * 4 > 200 * 1024) return cudaErrorInvalidValue;
    refine_select_kernel<<<b, RS_THREADS, 0, st>>>(n, k, cand, cand_cnt, cand_cap, eps, thr, flags, mode, *scratch, stats);
    static const int rs_split = [] { const char* v = getenv("SVSB_RESCORE_SPLIT"); const int x = v ? atoi(v) : 0; return x >= 1 && x <= 64 ? x : 4; }();
    rescore_kernel<<<dim3(b, rs_split), RS_THREADS, (size_t)ld * 4, st>>>(M, ld / 4, Q, ldq, *scratch);
    refine_sort_kernel<<<b, RS_THREADS, (size_t)REFINE_SURVIVOR_CAP * 8, st>>>(n, k, ids, row0, flags, mode, *scratch, out,
                                                                               push ? *push : BatchPush());
    count_launch(3);
    return cudaGetLastError();
}

cudaError_t preload_batch_kernels()
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, sample_order_kernel<true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, sample_order_kernel<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, union_threshold_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, refine_select_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, rescore_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, refine_sort_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, refine_lean_kernel<3>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, refine_lean_kernel<5>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, wait_flags_kernel);
    return e;
}

}  // namespace svsb
