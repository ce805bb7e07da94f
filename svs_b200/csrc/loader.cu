// Native scan of the SQLite `embeddings` table into the device matrix: the load path of
// `_Querier.build_embeddings_matrix` (reference src/svs/kb.py:573-618; table kb.py:80-83; blob codec
// src/svs/embeddings/util.py:15-23) without a per-row trip through the Python interpreter.
//
// SQLite is bound at run time from libsqlite3.so.0 (the image has the library but no header; the handful of entry
// points used are declared here).  The file is opened READ-ONLY on private connections -- the KB's own connection and
// its transaction are not touched; the caller (svs_b200/matrix.py) calls this from inside the transaction the
// reference opens for the rebuild (kb.py:872), when that connection has nothing uncommitted.
//
// Why it is fast.  A 6 KB blob lives in a 4 KB-page database as ~2 KB in the table leaf + one overflow page, so a scan
// is two page fetches and a reassembly per row; through Python that is ~4 us per row (1.3-1.6 GB/s, profiles/
// r01_load_bench.txt).  Here T threads each scan a contiguous range of rowids on their own connection (memory-mapped
// I/O: page fetches are pointers, not pread calls) straight into their own pinned slabs and copy them to the rows'
// final place in the device matrix -- row order == rowid order == the reference's scan order (kb.py:603-609), because
// the ranges are contiguous in rowid and each range's first row index is the sum of the counts of the ranges before it
// (counted by the same threads in a first, index-only pass over the table leaves).
#include "engine.cuh"

#include <dlfcn.h>

#include <chrono>
#include <thread>

using namespace svsb;

namespace {

struct sqlite3;
struct sqlite3_stmt;
constexpr int SQLITE_OK_ = 0, SQLITE_ROW_ = 100, SQLITE_DONE_ = 101, SQLITE_BLOB_ = 4;
constexpr int OPEN_READONLY = 0x1, OPEN_NOMUTEX = 0x8000;

struct Sqlite {
    void* lib = nullptr;
    int (*open_v2)(const char*, sqlite3**, int, const char*) = nullptr;
    int (*close_v2)(sqlite3*) = nullptr;
    int (*prepare_v2)(sqlite3*, const char*, int, sqlite3_stmt**, const char**) = nullptr;
    int (*step)(sqlite3_stmt*) = nullptr;
    long long (*column_int64)(sqlite3_stmt*, int) = nullptr;
    const void* (*column_blob)(sqlite3_stmt*, int) = nullptr;
    int (*column_bytes)(sqlite3_stmt*, int) = nullptr;
    int (*column_type)(sqlite3_stmt*, int) = nullptr;
    int (*finalize)(sqlite3_stmt*) = nullptr;
    int (*bind_int64)(sqlite3_stmt*, int, long long) = nullptr;
    const char* (*errmsg)(sqlite3*) = nullptr;
    int (*exec)(sqlite3*, const char*, int (*)(void*, int, char**, char**), void*, char**) = nullptr;
    int (*busy_timeout)(sqlite3*, int) = nullptr;
    bool ok = false;
    std::string why;
};

const Sqlite& sq() {
    static const Sqlite s = [] {
        Sqlite r;
        const char* names[] = {getenv("SVSB_SQLITE_LIB"), "libsqlite3.so.0", "libsqlite3.so"};
        for (const char* nme : names) {
            if (!nme || !*nme) continue;
            r.lib = dlopen(nme, RTLD_NOW | RTLD_LOCAL);
            if (r.lib) break;
        }
        if (!r.lib) { r.why = "libsqlite3.so.0 not found"; return r; }
        bool all = true;
#define SVSB_SYM(field, name) do { *(void**)(&r.field) = dlsym(r.lib, name); if (!r.field) { all = false; r.why = std::string("missing symbol ") + name; } } while (0)
        SVSB_SYM(open_v2, "sqlite3_open_v2"); SVSB_SYM(close_v2, "sqlite3_close_v2"); SVSB_SYM(prepare_v2, "sqlite3_prepare_v2");
        SVSB_SYM(step, "sqlite3_step"); SVSB_SYM(column_int64, "sqlite3_column_int64"); SVSB_SYM(column_blob, "sqlite3_column_blob");
        SVSB_SYM(column_bytes, "sqlite3_column_bytes"); SVSB_SYM(column_type, "sqlite3_column_type"); SVSB_SYM(finalize, "sqlite3_finalize");
        SVSB_SYM(bind_int64, "sqlite3_bind_int64"); SVSB_SYM(errmsg, "sqlite3_errmsg"); SVSB_SYM(exec, "sqlite3_exec");
        SVSB_SYM(busy_timeout, "sqlite3_busy_timeout");
#undef SVSB_SYM
        r.ok = all;
        return r;
    }();
    return s;
}

struct Conn {                                       // one read-only connection inside a read transaction
    sqlite3* db = nullptr;
    ~Conn() { if (db) { sq().exec(db, "ROLLBACK;", nullptr, nullptr, nullptr); sq().close_v2(db); } }
    int open(const char* path, std::string& err) {
        const Sqlite& S = sq();
        if (S.open_v2(path, &db, OPEN_READONLY | OPEN_NOMUTEX, nullptr) != SQLITE_OK_) {
            err = std::string("cannot open ") + path + ": " + (db ? S.errmsg(db) : "out of memory");
            return SVSB_E_STATE;
        }
        S.busy_timeout(db, 10000);
        S.exec(db, "PRAGMA mmap_size = 1099511627776;", nullptr, nullptr, nullptr);   // page fetches become pointers (capped by the build's maximum)
        S.exec(db, "PRAGMA cache_size = -65536;", nullptr, nullptr, nullptr);
        if (S.exec(db, "BEGIN;", nullptr, nullptr, nullptr) != SQLITE_OK_) { err = std::string("BEGIN: ") + S.errmsg(db); return SVSB_E_STATE; }
        return SVSB_OK;
    }
    // one-row query of integers
    int ints(const char* sql, const long long* binds, int nb, long long* out, int nout, bool* have_row, std::string& err) {
        const Sqlite& S = sq();
        sqlite3_stmt* st = nullptr;
        if (S.prepare_v2(db, sql, -1, &st, nullptr) != SQLITE_OK_) { err = std::string(sql) + ": " + S.errmsg(db); return SVSB_E_STATE; }
        for (int i = 0; i < nb; ++i) S.bind_int64(st, i + 1, binds[i]);
        const int rc = S.step(st);
        if (have_row) *have_row = rc == SQLITE_ROW_;
        if (rc == SQLITE_ROW_) for (int i = 0; i < nout; ++i) out[i] = S.column_type(st, i) == 5 ? 0 : S.column_int64(st, i);
        S.finalize(st);
        if (rc != SQLITE_ROW_ && rc != SQLITE_DONE_) { err = std::string(sql) + ": " + S.errmsg(db); return SVSB_E_STATE; }
        return SVSB_OK;
    }
};

// Where scanned rows go.  place(row, rows, ids, count): rows [row, row + count) of the scan order are at `rows` / `ids`
// (pinned memory owned by the scanning thread's slab) and must be consumed before the slab is reused: done(slab) is
// called before a slab is refilled.
struct Sink {
    virtual ~Sink() {}
    virtual int begin(int64_t n, int d) = 0;
    virtual int thread_begin(int t, float** slab_rows, int64_t** slab_ids, int64_t* slab_cap, int n_slabs) = 0;
    virtual int place(int t, int slab, int64_t row, int64_t count) = 0;
    virtual int wait_slab(int t, int slab) = 0;
    virtual int thread_end(int t) = 0;
};

struct Range { long long lo = 0, hi = -1; int64_t count = 0, row0 = 0; };

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int scan_table(const char* path, int threads, Sink& sink, int64_t* n_out, int32_t* d_out) {
    const Sqlite& S = sq();
    const bool dbg = env_int("SVSB_LOAD_DEBUG", 0) != 0;
    const double t_start = now_s();
    if (!S.ok) return fail(SVSB_E_STATE, "native SQLite scan unavailable: " + S.why);
    std::string err;
    Conn main;
    if (main.open(path, err) != SVSB_OK) return fail(SVSB_E_STATE, err);
    long long mm[2] = {0, 0};
    bool have = false;
    // two statements: SQLite turns a lone MIN() or MAX() into one b-tree descent, but scans the whole table for both at once
    if (main.ints("SELECT MIN(id) FROM embeddings;", nullptr, 0, &mm[0], 1, &have, err) != SVSB_OK) return fail(SVSB_E_STATE, err);
    if (main.ints("SELECT MAX(id) FROM embeddings;", nullptr, 0, &mm[1], 1, &have, err) != SVSB_OK) return fail(SVSB_E_STATE, err);
    long long first_len = -1;
    if (main.ints("SELECT LENGTH(embedding) FROM embeddings LIMIT 1;", nullptr, 0, &first_len, 1, &have, err) != SVSB_OK) return fail(SVSB_E_STATE, err);
    if (!have) {                                                 // empty table: shapes (0, 0) and (0,) (kb.py:595-601)
        int rc = sink.begin(0, 0);
        if (n_out) *n_out = 0;
        if (d_out) *d_out = 0;
        return rc;
    }
    if (first_len < 0 || first_len % 4 != 0 || first_len / 4 > 0x7fffffff)
        return fail(SVSB_E_STATE, "embedding blob length is not a multiple of 4 (embeddings/util.py:20-21)");
    const int d = (int)(first_len / 4);
    if (d == 0) return fail(SVSB_E_STATE, "zero-length embeddings: use the generic load path");
    const __int128 span = (__int128)mm[1] - (__int128)mm[0] + 1;  // rowids are any int64: no overflow here
    int T = std::max(1, std::min(threads, 64));
    if (span < (__int128)4096 * T) T = 1;
    std::vector<Range> ranges((size_t)T);
    for (int t = 0; t < T; ++t) {
        ranges[t].lo = (long long)((__int128)mm[0] + span * t / T);
        ranges[t].hi = (long long)((__int128)mm[0] + span * (t + 1) / T - 1);
    }
    std::vector<std::unique_ptr<Conn>> conns((size_t)T);
    std::vector<std::string> errs((size_t)T);
    std::vector<int> rcs((size_t)T, SVSB_OK);
    // pass 1: rows per range.  Counting in `embeddings` itself walks every table leaf (one 4 KB page per row at d = 1536:
    // as expensive as the scan); the reference's schema has `docs.embedding REFERENCES embeddings(id)` with an index on it
    // (kb.py:90, 96) and every embeddings row belongs to exactly one document, so the count comes from that small index
    // when the table exists.  A wrong count cannot go unnoticed: the scan below checks every range against it (kb.py:616)
    // and the caller falls back to the generic load.
    bool docs_index = env_int("SVSB_LOAD_COUNT_VIA_DOCS", 1) != 0;
    if (docs_index) {
        long long probe = 0; std::string perr;
        long long b0[2] = {mm[0], mm[0]};
        docs_index = main.ints("SELECT COUNT(*) FROM docs WHERE embedding BETWEEN ?1 AND ?2;", b0, 2, &probe, 1, nullptr, perr) == SVSB_OK;
    }
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
            conns[t].reset(new Conn());
            if ((rcs[t] = conns[t]->open(path, errs[t])) != SVSB_OK) return;
            long long b[2] = {ranges[t].lo, ranges[t].hi}, c = 0;
            rcs[t] = conns[t]->ints(docs_index ? "SELECT COUNT(*) FROM docs WHERE embedding BETWEEN ?1 AND ?2;"
                                               : "SELECT COUNT(*) FROM embeddings WHERE id BETWEEN ?1 AND ?2;", b, 2, &c, 1, nullptr, errs[t]);
            ranges[t].count = c;
        });
        for (auto& x : th) x.join();
        for (int t = 0; t < T; ++t) if (rcs[t] != SVSB_OK) return fail(rcs[t], errs[t]);
    }
    int64_t n = 0;
    for (auto& r : ranges) { r.row0 = n; n += r.count; }
    const double t_count = now_s();
    int rc = sink.begin(n, d);
    if (rc != SVSB_OK) return rc;
    const double t_begin = now_s();
    if (n_out) *n_out = n;
    if (d_out) *d_out = d;
    // pass 2: every thread scans its range in rowid order into its own slabs
    std::vector<std::string> terr((size_t)T);                    // g_err is thread-local: carry the workers' messages back
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
            constexpr int NS = 2;
            float* srows[NS]; int64_t* sids[NS]; int64_t cap = 0;
            auto bail = [&](int code) { rcs[t] = code; terr[t] = g_err; };
            int r = sink.thread_begin(t, srows, sids, &cap, NS);
            if (r != SVSB_OK) return bail(r);
            sqlite3_stmt* st = nullptr;
            sqlite3* db = conns[t]->db;
            if (S.prepare_v2(db, "SELECT id, embedding FROM embeddings WHERE id BETWEEN ?1 AND ?2;", -1, &st, nullptr) != SQLITE_OK_) {
                fail(SVSB_E_STATE, std::string("prepare scan: ") + S.errmsg(db)); return bail(SVSB_E_STATE);
            }
            S.bind_int64(st, 1, ranges[t].lo); S.bind_int64(st, 2, ranges[t].hi);
            int slab = 0; int64_t fill = 0, done = 0;
            const size_t row_bytes = (size_t)d * 4;
            int step_rc;
            while ((step_rc = S.step(st)) == SQLITE_ROW_) {
                if (done + fill >= ranges[t].count) { fail(SVSB_E_STATE, "more embedding rows than COUNT(*) reported (kb.py:616)"); r = SVSB_E_STATE; break; }
                if (fill == 0 && (r = sink.wait_slab(t, slab)) != SVSB_OK) break;
                const int bytes = S.column_bytes(st, 1);
                if ((size_t)bytes != row_bytes || S.column_type(st, 1) != SQLITE_BLOB_) {
                    fail(SVSB_E_STATE, "embedding rows of unequal length (kb.py:613)"); r = SVSB_E_STATE; break;
                }
                memcpy(reinterpret_cast<char*>(srows[slab]) + (size_t)fill * row_bytes, S.column_blob(st, 1), row_bytes);
                sids[slab][fill] = S.column_int64(st, 0);
                if (++fill == cap) {
                    if ((r = sink.place(t, slab, ranges[t].row0 + done, fill)) != SVSB_OK) break;
                    done += fill; fill = 0; slab = (slab + 1) % NS;
                }
            }
            if (r == SVSB_OK && step_rc != SQLITE_DONE_ && step_rc != SQLITE_ROW_) { fail(SVSB_E_STATE, std::string("scan: ") + S.errmsg(db)); r = SVSB_E_STATE; }
            S.finalize(st);
            if (r == SVSB_OK && fill > 0) { r = sink.place(t, slab, ranges[t].row0 + done, fill); done += fill; }
            if (r == SVSB_OK && done != ranges[t].count) { fail(SVSB_E_STATE, "fewer embedding rows than COUNT(*) reported (kb.py:616)"); r = SVSB_E_STATE; }
            const int r2 = sink.thread_end(t);
            if (r != SVSB_OK) return bail(r);
            if (r2 != SVSB_OK) return bail(r2);
        });
        for (auto& x : th) x.join();
        for (int t = 0; t < T; ++t) if (rcs[t] != SVSB_OK) return fail(rcs[t], terr[t]);
    }
    if (dbg)
        fprintf(stderr, "[svsb load] %lld rows x %d, %d threads: open + count %.3f s, allocate %.3f s, scan + copy %.3f s\n", (long long)n, d, T,
                t_count - t_start, t_begin - t_count, now_s() - t_begin);
    return SVSB_OK;
}

// ---- sink 1: plain host arrays (verification accessor; also what the CPU tests exercise) ------------------------
struct HostSink : Sink {
    float* rows; int64_t* ids; int64_t cap_rows; int d_expected;
    int64_t n = 0; int d = 0;
    struct Slabs { std::vector<float> r[2]; std::vector<int64_t> i[2]; int64_t cap = 0; };
    std::vector<Slabs> th;
    HostSink(float* r, int64_t* i, int64_t cap, int dexp, int T) : rows(r), ids(i), cap_rows(cap), d_expected(dexp), th((size_t)T) {}
    int begin(int64_t n_, int d_) override {
        n = n_; d = d_;
        if (((rows || ids) && n > cap_rows) || (n > 0 && d_expected >= 0 && d != d_expected))
            return fail(SVSB_E_INVALID, "svsb_sqlite_read: output buffers do not match the table");
        return SVSB_OK;
    }
    int thread_begin(int t, float** sr, int64_t** si, int64_t* cap, int ns) override {
        Slabs& s = th[t];
        s.cap = std::max<int64_t>(1, (4 << 20) / std::max(1, d * 4));
        for (int k = 0; k < ns; ++k) { s.r[k].resize((size_t)s.cap * d); s.i[k].resize((size_t)s.cap); sr[k] = s.r[k].data(); si[k] = s.i[k].data(); }
        *cap = s.cap;
        return SVSB_OK;
    }
    int place(int t, int slab, int64_t row, int64_t count) override {
        if (rows) memcpy(rows + row * d, th[t].r[slab].data(), (size_t)count * d * 4);
        if (ids) memcpy(ids + row, th[t].i[slab].data(), (size_t)count * 8);
        return SVSB_OK;
    }
    int wait_slab(int, int) override { return SVSB_OK; }
    int thread_end(int) override { return SVSB_OK; }
};

// ---- sink 2: the engine's device matrix -----------------------------------------------------------------------
struct DeviceSink : Sink {
    svsb_engine* e; int norm_mode;
    std::shared_ptr<Generation> gen;
    struct Th {
        float* r[2] = {nullptr, nullptr}; int64_t* i[2] = {nullptr, nullptr}; int64_t cap = 0;
        std::vector<cudaStream_t> st;               // per shard device
        cudaEvent_t ev[2] = {nullptr, nullptr}; bool pending[2] = {false, false};
    };
    std::vector<Th> th;
    DeviceSink(svsb_engine* e_, int nm, int T) : e(e_), norm_mode(nm), th((size_t)T) {}
    ~DeviceSink() override {
        for (auto& t : th) {
            for (int k = 0; k < 2; ++k) { if (t.r[k]) cudaFreeHost(t.r[k]); if (t.i[k]) cudaFreeHost(t.i[k]); if (t.ev[k]) cudaEventDestroy(t.ev[k]); }
            for (size_t s = 0; s < t.st.size(); ++s) if (t.st[s]) { cudaSetDevice(gen ? gen->shards[s].dev : 0); cudaStreamDestroy(t.st[s]); }
        }
    }
    int begin(int64_t n, int d) override {
        int rc = alloc_generation(e, n, d, gen);
        if (rc != SVSB_OK) return rc;
        if (gen->ld != d)                           // zero the padding columns once
            for (auto& s : gen->shards) if (s.n) {
                CU(cudaSetDevice(s.dev));
                CU(cudaMemset(s.M, 0, (size_t)s.n * gen->ld * 4));
                CU(cudaStreamSynchronize(cudaStreamLegacy));
            }
        return SVSB_OK;
    }
    int thread_begin(int t, float** sr, int64_t** si, int64_t* cap, int ns) override {
        Th& h = th[t];
        const int d = gen->d;
        h.cap = std::max<int64_t>(16, (16 << 20) / std::max(1, d * 4));
        h.st.assign(gen->shards.size(), nullptr);
        CU(cudaSetDevice(gen->shards[0].dev));
        for (int k = 0; k < ns; ++k) {
            CU(cudaHostAlloc(&h.r[k], (size_t)h.cap * std::max(1, d) * 4, cudaHostAllocPortable));
            CU(cudaHostAlloc(&h.i[k], (size_t)h.cap * 8, cudaHostAllocPortable));
            CU(cudaEventCreateWithFlags(&h.ev[k], cudaEventDisableTiming));
            sr[k] = h.r[k]; si[k] = h.i[k];
        }
        *cap = h.cap;
        return SVSB_OK;
    }
    int wait_slab(int t, int slab) override {
        Th& h = th[t];
        if (h.pending[slab]) { CU(cudaEventSynchronize(h.ev[slab])); h.pending[slab] = false; }
        return SVSB_OK;
    }
    int place(int t, int slab, int64_t row, int64_t count) override {
        Th& h = th[t];
        Generation* g = gen.get();
        int64_t done = 0;
        int last_dev_shard = -1;
        while (done < count) {
            const int64_t grow = row + done + e->shard_row0;
            size_t si = 0;
            while (si + 1 < g->shards.size() && grow >= g->shards[si].row0 + g->shards[si].n) ++si;
            Shard& sh = g->shards[si];
            const int64_t local = grow - sh.row0;
            const int64_t take = std::min(count - done, sh.n - local);
            CU(cudaSetDevice(sh.dev));
            if (!h.st[si]) CU(cudaStreamCreateWithFlags(&h.st[si], cudaStreamNonBlocking));
            const float* src = h.r[slab] + done * g->d;
            if (g->ld == g->d)
                CU(cudaMemcpyAsync(sh.M + local * g->ld, src, (size_t)take * g->d * 4, cudaMemcpyHostToDevice, h.st[si]));
            else
                CU(cudaMemcpy2DAsync(sh.M + local * g->ld, (size_t)g->ld * 4, src, (size_t)g->d * 4, (size_t)g->d * 4, (size_t)take,
                                     cudaMemcpyHostToDevice, h.st[si]));
            CU(cudaMemcpyAsync(sh.ids + local, h.i[slab] + done, (size_t)take * 8, cudaMemcpyHostToDevice, h.st[si]));
            if (last_dev_shard >= 0 && last_dev_shard != (int)si) {
                // the slab's reuse must wait for the copies to EVERY shard it fed: chain the earlier stream into this one
                cudaEvent_t tmp = h.ev[slab];
                CU(cudaSetDevice(g->shards[last_dev_shard].dev));
                CU(cudaEventRecord(tmp, h.st[last_dev_shard]));
                CU(cudaSetDevice(sh.dev));
                CU(cudaStreamWaitEvent(h.st[si], tmp, 0));
            }
            last_dev_shard = (int)si;
            done += take;
        }
        if (last_dev_shard >= 0) {
            CU(cudaSetDevice(g->shards[last_dev_shard].dev));
            CU(cudaEventRecord(h.ev[slab], h.st[last_dev_shard]));
            h.pending[slab] = true;
        }
        return SVSB_OK;
    }
    int thread_end(int t) override {
        Th& h = th[t];
        for (int k = 0; k < 2; ++k) if (h.pending[k]) { CU(cudaEventSynchronize(h.ev[k])); h.pending[k] = false; }
        return SVSB_OK;
    }
};

int default_threads() {
    int t = env_int("SVSB_LOAD_THREADS", 0);
    if (t > 0) return t;
    // measured on the 16-core GPU box, 1M x 1536 (profiles/r02_load_bench.txt): 1 / 2 / 4 / 8 / 16 connections give
    // 2.5 / 4.1 / 5.6 / 4.6 / 3.3 GB/s -- the scan is bound by page-cache faults and SQLite's allocator mutex, not by cores
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(4u, hw / 2));
}

}  // namespace

extern "C" int svsb_sqlite_available(void) { return sq().ok ? 1 : 0; }

extern "C" int svsb_sqlite_read(const char* path, int32_t threads, float* rows, int64_t* emb_ids, int64_t capacity_rows, int32_t d_expected,
                                int64_t* n_out, int32_t* d_out) {
    if (!path) return fail(SVSB_E_INVALID, "svsb_sqlite_read: path is NULL");
    const int T = threads > 0 ? threads : default_threads();
    HostSink sink(rows, emb_ids, capacity_rows, d_expected, std::max(1, std::min(T, 64)));
    return scan_table(path, T, sink, n_out, d_out);
}

extern "C" int svsb_load_sqlite(svsb_t* e, const char* path, int32_t norm_mode, int32_t threads, uint64_t* generation, int64_t* n_out,
                                int32_t* d_out) {
    if (!e || !path) return fail(SVSB_E_INVALID, "svsb_load_sqlite: NULL argument");
    if (norm_mode != SVSB_NORM_CHECK && norm_mode != SVSB_NORM_NORMALIZE) return fail(SVSB_E_INVALID, "svsb_load_sqlite: bad norm_mode");
    std::lock_guard<std::mutex> mlk(e->mutate_mu);
    if (e->loading) return fail(SVSB_E_STATE, "svsb_load_sqlite: another load is in progress");
    const int T = threads > 0 ? threads : default_threads();
    DeviceSink sink(e, norm_mode, std::max(1, std::min(T, 64)));
    int rc = scan_table(path, T, sink, n_out, d_out);
    if (rc != SVSB_OK) return rc;
    if (!sink.gen) return fail(SVSB_E_STATE, "svsb_load_sqlite: nothing was loaded");
    if ((rc = finish_generation(e, sink.gen.get(), norm_mode)) != SVSB_OK) return rc;
    sink.gen->norm_mode = norm_mode;
    return publish_generation(e, sink.gen, generation);
}
