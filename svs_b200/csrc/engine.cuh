// Internal structures of the engine (engine.cu) shared with the subsystems that live in their own files:
// multi.cu (several devices in one process), mutate.cu (incremental updates), loader.cu (native SQLite scan).
#pragma once
#include "../../include/svsb200.h"
#include "kernels.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

using namespace svsb;

// ------------------------------------------------------------------------------------------------
// errors, launch counter
// ------------------------------------------------------------------------------------------------
extern thread_local std::string g_err;
static inline int env_int(const char* name, int dflt) { const char* s = getenv(name); return s ? atoi(s) : dflt; }
extern std::atomic<int64_t> g_launches;

static inline int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            (void)cudaGetLastError();                                                              \
            return fail(_e == cudaErrorMemoryAllocation ? SVSB_E_NOMEM : SVSB_E_CUDA,              \
                        std::string(#call) + ": " + cudaGetErrorString(_e));                       \
        }                                                                                          \
    } while (0)

// ------------------------------------------------------------------------------------------------
// data structures
// ------------------------------------------------------------------------------------------------
// Device memory of one shard.  Shared by the generations of one lineage: an incremental update (svsb_apply_mutations)
// publishes a NEW generation that appends rows behind the old generation's last row in the same buffers, so the old
// generation -- and the in-flight queries that pin it -- keeps reading exactly the rows it had.
struct ShardBuf {
    int dev = 0;
    float* M = nullptr;          // cap_rows x ld floats
    int64_t* ids = nullptr;      // cap_rows
    int64_t cap_rows = 0;
    ~ShardBuf() {
        if (M || ids) cudaSetDevice(dev);
        if (M) cudaFree(M);
        if (ids) cudaFree(ids);
    }
};

struct Shard {
    int dev = 0;
    int64_t row0 = 0, n = 0;     // n = rows of the buffer this generation uses (physical rows, tombstoned ones included)
    float* M = nullptr;          // == buf->M, buf->ids (raw copies for the hot path)
    int64_t* ids = nullptr;
    std::shared_ptr<ShardBuf> buf;
    // tombstones (svsb_apply_mutations): one byte per physical row, 1 = live; null = every row is live.  Owned by the
    // generation (a delete must not disturb queries still running on the previous generation).
    uint8_t* live = nullptr;
    int64_t n_live = 0;
};

struct Generation {
    uint64_t id = 0;
    int64_t n = 0;               // physical rows over all shards
    int64_t n_live = 0;          // rows a query can return (== n unless rows were tombstoned)
    int d = 0, ld = 0;
    int norm_mode = SVSB_NORM_CHECK;
    bool owns_live = true;       // false for the per-device views of a multi-device generation (child_gen)
    std::vector<Shard> shards;
    float max_dev = 0.f;
    int64_t n_out_of_tol = 0;
    // multi-device engines: one single-shard view per device, what that device's shard engine queries
    std::vector<std::shared_ptr<Generation>> child_gen;
    // fp16 shadow of shard 0 for the batched coarse contraction (built lazily by the first batch)
    std::mutex m16_mu;
    void* M16 = nullptr; int ld16 = 0;
    // set once a batch saw its statistical filter thresholds fail verification (rows not in random order w.r.t. the
    // queries): later batches of this generation use the guaranteed thresholds
    std::atomic<bool> batch_guaranteed{false};
    bool has_tombstones() const { return n_live != n; }
    ~Generation() {
        child_gen.clear();
        if (owns_live)
            for (auto& s : shards) if (s.live) { cudaSetDevice(s.dev); cudaFree(s.live); }
        if (M16 && !shards.empty()) { cudaSetDevice(shards[0].dev); cudaFree(M16); }
    }
};

// per-device scratch for one in-flight query
struct svsb_workspace {
    int dev = 0;
    cudaStream_t st = nullptr;       // owned when created by the engine; null for the stateless API
    bool own_stream = false;
    float* d_q = nullptr;   int q_cap = 0;
    float* scores = nullptr; int64_t n_cap = 0;
    u64* gmax = nullptr;     int64_t g_cap = 0;
    u64* cand = nullptr;     int64_t cand_cap = 0;
    u64* sortbuf = nullptr;  int64_t sort_cap = 0;
    u64* out_keys = nullptr; float* out_scores = nullptr; int64_t* out_ids = nullptr; int64_t out_cap = 0;
    int32_t* out_count = nullptr;
    u64* mscr_keys = nullptr; int64_t* mscr_ids = nullptr; int64_t mscr_cap = 0;   // merge scratch
    cudaEvent_t ev = nullptr, ev0 = nullptr, ev1 = nullptr, ev_sel = nullptr;
    // The similarity kernel leaves group maxima in gmax and only a SUCCESSFUL selection kernel re-zeroes them
    // (select.cu).  Set before the similarity launch, cleared once the selection is enqueued: a call that failed in
    // between leaves it set and the next user of this (pooled) workspace re-zeroes gmax first.
    bool gmax_dirty = false;

    int ensure_rows(int64_t n) {
        cudaSetDevice(dev);
        if (gmax_dirty && gmax && n <= n_cap) {
            CU(cudaMemset(gmax, 0, (size_t)g_cap * 8));
            CU(cudaStreamSynchronize(cudaStreamLegacy));
            gmax_dirty = false;
        }
        if (n > n_cap) {
            if (scores) cudaFree(scores); if (gmax) cudaFree(gmax); if (cand) cudaFree(cand);
            scores = nullptr; gmax = nullptr; cand = nullptr; n_cap = 0;
            const int shift = group_shift_for(n);
            const int64_t G = (n + ((int64_t)1 << shift) - 1) >> shift;
            int64_t cc = (int64_t)K_FAST_MAX << shift; if (cc > n) cc = n;
            cc += K_FAST_MAX;                       // room for the overflow path to re-home the sort buffer
            CU(cudaMalloc(&scores, (size_t)n * 4));
            CU(cudaMalloc(&gmax, (size_t)G * 8));
            CU(cudaMemset(gmax, 0, (size_t)G * 8));
            // cudaMemset runs on the legacy default stream, asynchronously to the host, and the engine's streams are
            // non-blocking: without this wait a similarity kernel could write group maxima BEFORE the zeroing lands
            CU(cudaStreamSynchronize(cudaStreamLegacy));
            CU(cudaMalloc(&cand, (size_t)cc * 8));
            n_cap = n; g_cap = G; cand_cap = cc; gmax_dirty = false;
        }
        return SVSB_OK;
    }
    int ensure_q(int ld) {
        cudaSetDevice(dev);
        if (ld > q_cap) { if (d_q) cudaFree(d_q); d_q = nullptr; CU(cudaMalloc(&d_q, (size_t)ld * 4)); q_cap = ld; }
        return SVSB_OK;
    }
    int ensure_out(int64_t k) {
        cudaSetDevice(dev);
        if (!out_count) CU(cudaMalloc(&out_count, 64));
        if (k > out_cap) {
            if (out_keys) cudaFree(out_keys); if (out_scores) cudaFree(out_scores); if (out_ids) cudaFree(out_ids);
            out_keys = nullptr; out_scores = nullptr; out_ids = nullptr; out_cap = 0;
            CU(cudaMalloc(&out_keys, (size_t)k * 8));
            CU(cudaMalloc(&out_scores, (size_t)k * 4));
            CU(cudaMalloc(&out_ids, (size_t)k * 8));
            out_cap = k;
        }
        return SVSB_OK;
    }
    int ensure_sort(int64_t n) {
        cudaSetDevice(dev);
        int64_t need = next_pow2(n); if (need < 2048) need = 2048;
        if (need > sort_cap) { if (sortbuf) cudaFree(sortbuf); sortbuf = nullptr; CU(cudaMalloc(&sortbuf, (size_t)need * 8)); sort_cap = need; }
        return SVSB_OK;
    }
    int ensure_merge_scratch(int64_t entries) {
        cudaSetDevice(dev);
        if (entries > mscr_cap) {
            if (mscr_keys) cudaFree(mscr_keys); if (mscr_ids) cudaFree(mscr_ids);
            mscr_keys = nullptr; mscr_ids = nullptr; mscr_cap = 0;
            CU(cudaMalloc(&mscr_keys, (size_t)entries * 8));
            CU(cudaMalloc(&mscr_ids, (size_t)entries * 8));
            mscr_cap = entries;
        }
        return SVSB_OK;
    }
    void release() {
        cudaSetDevice(dev);
        void* ptrs[] = {d_q, scores, gmax, cand, sortbuf, out_keys, out_scores, out_ids, out_count, mscr_keys, mscr_ids};
        for (void* p : ptrs) if (p) cudaFree(p);
        if (ev) cudaEventDestroy(ev); if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1);
        if (ev_sel) cudaEventDestroy(ev_sel);
        if (own_stream && st) cudaStreamDestroy(st);
    }
};
typedef svsb_workspace DevWs;

struct QueryCtx {
    std::vector<DevWs> ws;                 // one per engine device
    std::unique_ptr<DevWs> alt;            // second buffer set + the selection stream of the pipelined bench loop (device 0)
    DevWs* last = nullptr;                 // workspace holding the last single-device bench result
    // pinned host staging
    float* h_q = nullptr; int h_q_cap = 0;
    float* h_scores = nullptr; int64_t* h_ids = nullptr; int32_t* h_count = nullptr; int64_t h_out_cap = 0;
    // device-0 gather + merge outputs (multi-device)
    u64* g_keys = nullptr; int64_t* g_ids = nullptr; int32_t* g_counts = nullptr; int64_t g_stride = 0;
    float* m_scores = nullptr; int64_t* m_ids = nullptr; int32_t* m_count = nullptr; int64_t m_cap = 0;
    cudaEvent_t ev_merge = nullptr; bool merge_recorded = false;   // last merge that read g_keys / g_ids
    cudaEvent_t ev_gemv = nullptr;         // svsb_query_submit: this context's similarity pass has finished
};

struct Slab {
    float* rows = nullptr; int64_t* ids = nullptr;
    std::vector<cudaEvent_t> ev;           // per device: last copy out of this slab
    std::vector<char> pending;
};

struct Loading {
    std::shared_ptr<Generation> gen;
    int norm_mode = SVSB_NORM_CHECK;
    int64_t loaded = 0;
    int64_t slab_rows = 0;
    int cur = 0;
    bool borrowed = false;
};

// workspace of the batched path: coarse operands, thresholds, candidate lists, outputs (device) + pinned staging
struct BatchWs {
    int dev = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // capacities
    int cap_b = 0, cap_ld = 0, cap_k = 0; int64_t cap_sample = 0; int cand_cap = 0;
    float* dQ = nullptr; void* dQ16 = nullptr;
    float* eps = nullptr; float* thr = nullptr; int32_t* flags = nullptr; int32_t* cand_cnt = nullptr; int32_t* stats = nullptr;
    float* sample = nullptr; u64* cand = nullptr;
    RefineScratch rs = {nullptr, nullptr, nullptr, nullptr};     // split refine: survivors' rows / exact keys / counts
    float* tops = nullptr;                                        // [cap_b][SAMPLE_TOPX]: this rank's largest sample values (sharded path)
    uint64_t global_gen = 0; int global_b = 0, global_k = 0;      // svsb_batch_sample_tops -> svsb_batch_global_records hand-over
    float* o_scores = nullptr; int64_t* o_ids = nullptr; int32_t* o_counts = nullptr;
    float* h_Q = nullptr; float* h_scores = nullptr; int64_t* h_ids = nullptr; int32_t* h_counts = nullptr;
    int32_t* h_flags = nullptr; int32_t* h_stats = nullptr; int32_t* h_cnt = nullptr;

    void release_device() {
        cudaSetDevice(dev);
        void* ptrs[] = {dQ, dQ16, eps, thr, flags, cand_cnt, stats, sample, cand, o_scores, o_ids, o_counts,
                        rs.rows, rs.keys, rs.cnt, rs.ver, tops};
        for (void* p : ptrs) if (p) cudaFree(p);
        void* hp[] = {h_Q, h_scores, h_ids, h_counts, h_flags, h_stats, h_cnt};
        for (void* p : hp) if (p) cudaFreeHost(p);
        dQ = nullptr; dQ16 = nullptr; eps = thr = nullptr; flags = cand_cnt = stats = nullptr; sample = nullptr; cand = nullptr;
        o_scores = nullptr; o_ids = nullptr; o_counts = nullptr;
        rs = {nullptr, nullptr, nullptr, nullptr}; tops = nullptr;
        h_Q = h_scores = nullptr; h_ids = nullptr; h_counts = h_flags = h_stats = h_cnt = nullptr;
        cap_b = cap_ld = cap_k = 0; cap_sample = 0; cand_cap = 0;
    }
    void release() {
        release_device();
        for (auto& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
        if (st) { cudaStreamDestroy(st); st = nullptr; }
    }
};

// workspace of the small-batch exact path (gemv_tma_mq_kernel + one selection CTA per query)
struct MqWs {
    int dev = 0;
    cudaStream_t st = nullptr;
    int cap_b = 0, cap_ld = 0; int64_t cap_n = 0, G = 0, cand_cap = 0, cap_out = 0;
    float* dQ = nullptr; float* scores = nullptr; u64* gmax = nullptr; u64* cand = nullptr;
    u64* o_keys = nullptr; float* o_scores = nullptr; int64_t* o_ids = nullptr; int32_t* o_counts = nullptr;
    float* h_Q = nullptr;
    bool gmax_dirty = false;

    int ensure(int b, int64_t n, int ld, int64_t k) {
        cudaSetDevice(dev);
        if (!st) CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (b > cap_b || n > cap_n) {
            CU(cudaStreamSynchronize(st));
            const int nb = std::max(b, cap_b); const int64_t nn = std::max(n, cap_n);
            void* ptrs[] = {scores, gmax, cand, o_counts};
            for (void* p : ptrs) if (p) cudaFree(p);
            scores = nullptr; gmax = nullptr; cand = nullptr; o_counts = nullptr; cap_b = 0; cap_n = 0;
            const int shift = group_shift_for(nn);
            G = (nn + ((int64_t)1 << shift) - 1) >> shift;
            cand_cap = std::min<int64_t>((int64_t)K_FAST_MAX << shift, nn) + K_FAST_MAX;
            CU(cudaMalloc(&scores, (size_t)nb * nn * 4));
            CU(cudaMalloc(&gmax, (size_t)nb * G * 8));
            CU(cudaMalloc(&cand, (size_t)nb * cand_cap * 8));
            CU(cudaMalloc(&o_counts, (size_t)nb * 4));
            cap_b = nb; cap_n = nn; gmax_dirty = true; cap_out = 0; cap_ld = 0;
        }
        if (gmax_dirty) {
            CU(cudaMemsetAsync(gmax, 0, (size_t)cap_b * G * 8, st));
            gmax_dirty = false;
        }
        if (ld > cap_ld) {
            CU(cudaStreamSynchronize(st));
            if (dQ) cudaFree(dQ);
            if (h_Q) cudaFreeHost(h_Q);
            dQ = nullptr; h_Q = nullptr; cap_ld = 0;
            CU(cudaMalloc(&dQ, (size_t)cap_b * ld * 4));
            CU(cudaMallocHost(&h_Q, (size_t)cap_b * ld * 4));
            cap_ld = ld;
        }
        if ((int64_t)cap_b * k > cap_out) {
            CU(cudaStreamSynchronize(st));
            void* ptrs[] = {o_keys, o_scores, o_ids};
            for (void* p : ptrs) if (p) cudaFree(p);
            o_keys = nullptr; o_scores = nullptr; o_ids = nullptr; cap_out = 0;
            const int64_t e = (int64_t)cap_b * k;
            CU(cudaMalloc(&o_keys, (size_t)e * 8));
            CU(cudaMalloc(&o_scores, (size_t)e * 4));
            CU(cudaMalloc(&o_ids, (size_t)e * 8));
            cap_out = e;
        }
        return SVSB_OK;
    }
    void release() {
        cudaSetDevice(dev);
        void* ptrs[] = {dQ, scores, gmax, cand, o_keys, o_scores, o_ids, o_counts};
        for (void* p : ptrs) if (p) cudaFree(p);
        if (h_Q) cudaFreeHost(h_Q);
        if (st) cudaStreamDestroy(st);
    }
};

// Batch window of the one-process-per-GPU deployment (svsb_bxchg_*, svsb_batch_peer): per slot and source rank a region
// for the rank's sample maxima and one for its candidate records, written by the peers' kernels over NVLink peer memory.
//   [flags: 2 phases x slots x world u64, padded to 256 B][tops: slots x world x B_MAX x 32 f32][records: slots x world x B_MAX x (2 rec_cap + 1) i64]
struct BXchg {
    static constexpr int SLOTS = 2, B_MAX = COARSE_MAX_BATCH;
    int world = 0, rank = 0, rec_cap = 0;
    unsigned char* block = nullptr;
    size_t flags_bytes = 0, tops_bytes = 0, bytes = 0;
    std::vector<unsigned char*> peer_block;
    std::vector<void*> ipc_opened;
    bool connected = false;
    unsigned long long seq = 0;
    unsigned long long timeout_ns = 30ull * 1000000000ull;
    unsigned int* done = nullptr;            // two last-block tickets (sample maxima, records)
    int* status = nullptr;                   // device word: 1 = a wait timed out
    // svsb_batch_peer with flag bit 0: the verifying merge of batch j is enqueued behind the FIRST phase of batch j+1 (or by
    // svsb_batch_peer_flush), so the peers have that long to deliver their records before this rank's stream waits for them
    struct Deferred {
        bool pending = false;
        int slot = 0, rec_cap = 0, k = 0, b = 0; unsigned long long seq = 0;
        float* out_scores = nullptr; int64_t* out_ids = nullptr; int32_t* out_counts = nullptr;
    } deferred;
    int64_t tops_region() const { return (int64_t)B_MAX * SAMPLE_TOPX; }                       // floats per (slot, source)
    int64_t rec_region() const { return (int64_t)B_MAX * (2 * (int64_t)rec_cap + 1); }         // words per (slot, source)
    u64* flag(unsigned char* blk, int phase, int slot, int src) const { return reinterpret_cast<u64*>(blk) + ((int64_t)phase * SLOTS + slot) * world + src; }
    float* tops(unsigned char* blk, int slot, int src) const { return reinterpret_cast<float*>(blk + flags_bytes) + ((int64_t)slot * world + src) * tops_region(); }
    int64_t* recs(unsigned char* blk, int slot, int src) const { return reinterpret_cast<int64_t*>(blk + flags_bytes + tops_bytes) + ((int64_t)slot * world + src) * rec_region(); }
};

// Peer exchange of the one-process-per-GPU deployment (kernels.cuh "peer exchange"): this rank's gather window,
// the peers' windows opened over CUDA IPC (or plain pointers inside one process), and a synchronous query context.
struct Xchg {
    int world = 0, rank = 0, cap = 0, slots = 4;
    int64_t rec_words = 0;
    unsigned char* block = nullptr;                      // [flags: slots*world u64, padded to 256 B][window]
    size_t flags_bytes = 0;
    std::vector<unsigned char*> peer_block;              // per rank (own entry = block)
    std::vector<void*> ipc_opened;
    bool connected = false;
    unsigned long long seq = 0;
    unsigned long long timeout_ns = 30ull * 1000000000ull;   // merge kernel gives up waiting for a peer (SVSB_XCHG_TIMEOUT_MS)
    // Pipelined path: the merge of query j is enqueued on the side stream BEHIND the selection of query j+1, so a
    // peer has a whole query time to deliver its record before this rank's side stream would wait for it.
    struct DeferredMerge {
        bool pending = false;
        int slot = 0, k = 0; unsigned long long seq = 0;
        u64* sk = nullptr; int64_t* sp = nullptr;
        float* out_scores = nullptr; int64_t* out_ids = nullptr; int32_t* out_count = nullptr;
        cudaEvent_t ev_done = nullptr;                   // recorded behind the merge once it is enqueued (submit / wait tickets)
    } deferred;
    // svsb_query_peer_submit / _wait: host-buffer queries with several in flight.  A ticket owns its pinned query and
    // result buffers and a device copy of the query (the next query is staged while this one's similarity pass runs).
    struct Ticket {
        float* h_q = nullptr; float* d_q = nullptr;
        float* h_scores = nullptr; int64_t* h_ids = nullptr; int32_t* h_count = nullptr;
        cudaEvent_t ev = nullptr;
        bool busy = false; int k = 0; unsigned long long seq = 0;
    };
    static constexpr int N_TICKETS = 3;
    Ticket tk[N_TICKETS];
    int tk_ld = 0, tk_next = 0;
    // measurement aid (SVSB_XCHG_STAMPS=1): %globaltimer stamps of the last STAMP_RING synchronous peer queries --
    // selection kernel phases (16 words) + merge kernel (24 words) per query
    static constexpr int STAMP_RING = 1024, STAMP_WORDS = 40;
    u64* stamps = nullptr;
    cudaEvent_t ev_join = nullptr;
    // synchronous path (svsb_query_peer): own stream, device query, pinned staging; results land in pinned
    // host memory straight from the merge kernel (mapped, no copy back)
    cudaStream_t st = nullptr;
    DevWs ws;
    float* h_q = nullptr; int h_q_cap = 0;
    float* h_scores = nullptr; int64_t* h_ids = nullptr; int32_t* h_count = nullptr; int64_t h_cap = 0;
    u64* flags_of(unsigned char* b, int slot) const { return reinterpret_cast<u64*>(b) + (size_t)slot * world; }
    u64* rec_of(unsigned char* b, int slot, int src) const {
        return reinterpret_cast<u64*>(b + flags_bytes) + ((size_t)slot * world + src) * rec_words;
    }
};

struct svsb_engine {
    std::vector<int> devs;
    std::mutex mu;                          // guards current, loading, pool bookkeeping
    std::condition_variable cv;
    std::shared_ptr<Generation> current;
    uint64_t next_gen = 1;
    std::unique_ptr<Loading> loading;
    std::vector<Slab> slabs; int64_t slab_bytes_rows = 0, slab_ids_cap = 0;
    std::vector<cudaStream_t> copy_st;      // per device
    std::vector<std::unique_ptr<QueryCtx>> pool_free;
    int ctx_total = 0, ctx_max = 4;
    // svsb_query_submit keeps several contexts in flight.  Their similarity passes all go down ONE stream, back to back
    // (pass j+1 starts when pass j ends, so the one SM each pass leaves free really is free for the single-CTA
    // selections, which run on the contexts' own streams) -- the structure of the device-resident pipelined loop.
    std::mutex chain_mu;
    cudaStream_t submit_st = nullptr;
    // batched path (one batch at a time)
    std::mutex batch_mu;
    std::unique_ptr<struct BatchWs> batch_ws;
    std::unique_ptr<MqWs> mq_ws;            // small exact batches (guarded by batch_mu as well)
    // bench state
    std::vector<float*> bench_q; int bench_nq = 0, bench_d = 0, bench_ld = 0;
    std::unique_ptr<QueryCtx> bench_ctx;
    std::vector<cudaEvent_t> bench_kev;                  // similarity-kernel timing events of svsb_bench_run (device 0)
    // sharded deployment (one process per GPU)
    int64_t shard_row0 = 0;                              // global row of this engine's first row
    std::vector<std::unique_ptr<DevWs>> shard_ws;        // workspaces for svsb_enqueue_local_topk (by slot)
    std::vector<cudaEvent_t> kev; size_t kev_used = 0;   // similarity-kernel timing events
    cudaStream_t side_st = nullptr;                      // selection kernels of the pipelined sharded path
    std::vector<char> sel_pending;                       // per slot: a selection is (or was) in flight on side_st
    std::unique_ptr<Xchg> xchg;
    std::unique_ptr<BXchg> bxchg;
    // several devices in ONE process (svsb_create with n_dev > 1, multi.cu): one shard engine per device, driven by one
    // worker thread each, candidate records pushed into device 0's gather window by the selection kernels
    struct Multi* multi = nullptr;
    bool is_kid = false;                                 // a per-device shard engine owned by a multi-device engine
    std::mutex mutate_mu;                                // one svsb_apply_mutations / load publication at a time
};

static inline int round_up4(int d) { return (d + 3) & ~3; }

// a query handle that keeps "the arrays it fetched" alive across an invalidate (svsb_snapshot_*)
struct svsb_snapshot { std::shared_ptr<Generation> gen; };

// ------------------------------------------------------------------------------------------------
// internal interfaces between engine.cu and the subsystem files
// ------------------------------------------------------------------------------------------------
std::shared_ptr<Generation> pin(svsb_engine* e);
int prepare_ws(DevWs& w, const Generation* g, const Shard& s, int64_t kk);
int engine_create(const int* device_ids, int n_dev, bool as_kid, svsb_engine** out);
int alloc_generation(svsb_engine* e, int64_t n, int d, std::shared_ptr<Generation>& out);
int finish_generation(svsb_engine* e, Generation* g, int norm_mode);
// publish `g` as the engine's resident generation (assigns the id; builds the per-device views of a multi-device engine)
int publish_generation(svsb_engine* e, const std::shared_ptr<Generation>& g, uint64_t* generation);
// local top-k records [b][2k+1] of b device-resident queries on `st` (engine.cu; the body of svsb_batch_local_records)
int batch_local_records_gen(svsb_engine* e, const std::shared_ptr<Generation>& g, cudaStream_t st, const float* d_Q, int32_t b,
                            int32_t k, int64_t* d_records, int32_t* n_fallback);

// multi.cu
int  multi_create(svsb_engine* e);
void multi_destroy(svsb_engine* e);
int  multi_publish(svsb_engine* e, const std::shared_ptr<Generation>& g);
int  multi_query(svsb_engine* e, const std::shared_ptr<Generation>& g, const float* q, int32_t d, int64_t kk,
                 float* out_scores, int64_t* out_ids, int32_t* out_count);
struct svsb_ticket;
int  multi_submit(svsb_engine* e, const std::shared_ptr<Generation>& g, const float* q, int32_t d, int64_t kk, svsb_ticket** out);
int  multi_wait(svsb_engine* e, svsb_ticket* t, float* out_scores, int64_t* out_ids, int32_t* out_count);
int  multi_query_batch(svsb_engine* e, const std::shared_ptr<Generation>& g, const float* Q, int32_t b, int32_t d, int32_t k,
                       float* out_scores, int64_t* out_ids, int32_t* out_counts);
int  multi_bench_run(svsb_engine* e, const std::shared_ptr<Generation>& g, int32_t k, int32_t iters, float* total_ms, float* gemv_ms);
int  multi_bench_last_result(svsb_engine* e, int32_t k, float* out_scores, int64_t* out_ids, int32_t* out_count);
