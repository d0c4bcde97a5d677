// Host transpose: []abstract.ChangeItem in row form (tf_rows, include/tfgpu_sink.h) <-> the columnar tf_batch the device
// consumes. The reference keeps a batch as an array of structs whose values are boxed interfaces (pkg/abstract/changeitem/change_item.go:27-78);
// the shim flattens that into one byte image with plain appends and this file turns the image into column buffers in two parallel passes
// (sizes, then fill), into pooled pinned memory so that tfgpu_push_* can start its H2D copies without another staging copy.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

#include "../../include/tfgpu_sink.h"
#include "plan.hpp"
#include "row_image.hpp"
#include "host_internal.hpp"

namespace {

struct Fail { int rc; std::string msg; };

struct Buf {
    uint8_t* p = nullptr; size_t cap = 0; bool pinned = false;
    void release() { if (!p) return; if (pinned) cudaFreeHost(p); else std::free(p); p = nullptr; cap = 0; }
    uint8_t* ensure(size_t n, bool want_pinned) {
        if (n <= cap && p) return p;
        release();
        const size_t want = std::max<size_t>(4096, n + n / 4 + 256) & ~(size_t)255;
        if (want_pinned && cudaHostAlloc((void**)&p, want, cudaHostAllocDefault) == cudaSuccess) pinned = true;
        else { (void)cudaGetLastError(); pinned = false; p = (uint8_t*)std::aligned_alloc(256, want); if (!p) throw Fail{TF_E_RETRY_OOM, "host allocation failed"}; }
        cap = want; return p;
    }
};

inline int fixed_width(int tf) {
    switch (tf) {
    case TF_INT8: case TF_UINT8: case TF_BOOLEAN: return 1;
    case TF_INT16: case TF_UINT16: return 2;
    case TF_INT32: case TF_UINT32: case TF_FLOAT: return 4;
    case TF_INT64: case TF_UINT64: case TF_DOUBLE: case TF_INTERVAL: case TF_DATE: case TF_DATETIME: case TF_TIMESTAMP: return 8;
    default: return 0;
    }
}
inline bool is_time(int tf) { return tf == TF_DATE || tf == TF_DATETIME || tf == TF_TIMESTAMP; }

// the canonical Go type of a YT type in a strict ChangeItem (typesystem/values/type_checkers.go:39-84) as a value tag
inline int canonical_tag(int tf) {
    switch (tf) {
    case TF_INT8: return TF_V_INT8; case TF_INT16: return TF_V_INT16; case TF_INT32: return TF_V_INT32; case TF_INT64: return TF_V_INT64;
    case TF_UINT8: return TF_V_UINT8; case TF_UINT16: return TF_V_UINT16; case TF_UINT32: return TF_V_UINT32; case TF_UINT64: return TF_V_UINT64;
    case TF_FLOAT: return TF_V_FLOAT32; case TF_DOUBLE: return TF_V_FLOAT64; case TF_BOOLEAN: return TF_V_BOOL;
    case TF_BYTES: return TF_V_BYTES; case TF_UTF8: return TF_V_STRING; case TF_ANY: return TF_V_JSON;
    case TF_INTERVAL: return TF_V_DURATION; default: return TF_V_TIME;
    }
}
// the fixed-width column type a numeric tag travels as (a loose column when it differs from the schema's)
static const uint8_t TAG_TF[18] = {0, TF_BOOLEAN, TF_INT8, TF_INT16, TF_INT32, TF_INT64, TF_UINT8, TF_UINT16, TF_UINT32, TF_UINT64, TF_FLOAT, TF_DOUBLE, 0, 0, 0, TF_INTERVAL, TF_DOUBLE, 0};
inline int tag_tf_slow(int tag) {
    switch (tag) {
    case TF_V_BOOL: return TF_BOOLEAN;
    case TF_V_INT8: return TF_INT8; case TF_V_INT16: return TF_INT16; case TF_V_INT32: return TF_INT32; case TF_V_INT64: return TF_INT64;
    case TF_V_UINT8: return TF_UINT8; case TF_V_UINT16: return TF_UINT16; case TF_V_UINT32: return TF_UINT32; case TF_V_UINT64: return TF_UINT64;
    case TF_V_FLOAT32: return TF_FLOAT; case TF_V_FLOAT64: case TF_V_JSONNUM: return TF_DOUBLE; case TF_V_DURATION: return TF_INTERVAL;
    default: return 0;
    }
}
inline int tag_tf(int tag) { return TAG_TF[tag]; }
// text an `any` cell holds for a scalar value: what json.Marshal writes for Go ints / bools
inline uint32_t any_scalar_text(const Val& v, char* out) {
    if (v.tag == TF_V_BOOL) { const char* t = v.p[0] ? "true" : "false"; const uint32_t n = v.p[0] ? 4 : 5; std::memcpy(out, t, n); return n; }
    if (v.tag == TF_V_UINT64) { uint64_t x; std::memcpy(&x, v.p, 8); return (uint32_t)std::snprintf(out, 24, "%llu", (unsigned long long)x); }
    return (uint32_t)std::snprintf(out, 24, "%lld", (long long)val_i64(v));
}

struct ColPlan {
    int schema_tf = 0, phys_tf = 0, width = 0, lens_width = 0;
    bool has_nil = false, has_nsec = false, has_anytag = false;
    uint32_t tagmask = 0; uint64_t heap_total = 0; uint32_t max_len = 0;
    uint8_t *values = nullptr, *validity = nullptr, *lens = nullptr, *heap = nullptr, *aux = nullptr;
};
struct ChunkStat { uint32_t tagmask = 0; uint64_t heap = 0; uint32_t max_len = 0; bool nil = false, nsec = false; };

constexpr uint32_t M_SIGNED = (1u << TF_V_INT8) | (1u << TF_V_INT16) | (1u << TF_V_INT32) | (1u << TF_V_INT64) | (1u << TF_V_BOOL);
constexpr uint32_t M_UNSIGNED = (1u << TF_V_UINT8) | (1u << TF_V_UINT16) | (1u << TF_V_UINT32) | (1u << TF_V_UINT64);
constexpr uint32_t M_FLOAT = (1u << TF_V_FLOAT32) | (1u << TF_V_FLOAT64) | (1u << TF_V_JSONNUM);   // json.Number: the strict form of `double` (type_checkers.go:63-65)
constexpr uint32_t M_TEXT = (1u << TF_V_STRING) | (1u << TF_V_BYTES);

}  // namespace

struct tfgpu_columnar {
    std::string err;
    bool want_pinned = true;
    std::map<std::string, std::vector<int>> schemas;      // schema_json -> tf types
    std::vector<Buf> bufs; size_t next_buf = 0;
    Buf* take() { if (next_buf == bufs.size()) bufs.emplace_back(); return &bufs[next_buf++]; }
    std::vector<Buf> tmp_bufs; size_t next_tmp = 0;       // scratch of the strict transposer path (never handed out, never pinned)
    uint8_t* take_tmp(size_t n) { if (next_tmp == tmp_bufs.size()) tmp_bufs.emplace_back(); return tmp_bufs[next_tmp++].ensure(n, false); }
    // results of the last call
    std::vector<tf_col> cols, old_cols; tf_batch batch{}, old_batch{}; tf_row_meta meta{}; tf_old_keys old{};
    std::vector<uint8_t> present;
    std::vector<uint8_t> text_mixed;      // [col] of the last tfgpu_rows_to_batch: a text cell carried Go's other text type (see host_internal.hpp)
    void* workers = nullptr;              // Workers, created on first use
    ~tfgpu_columnar();
};

namespace {

// Persistent host workers (spawning 32 threads per pass costs more than the pass itself on a 1 M-row gather): run(count, width, fn)
// hands task indexes [0, count) to `width` workers (the caller is one of them) and returns when all are done.
class Workers {
public:
    ~Workers() { { std::lock_guard<std::mutex> g(m_); stop_ = true; } cv_.notify_all(); for (auto& t : ts_) t.join(); }
    void run(uint64_t count, int width, const std::function<void(uint64_t)>& fn) {
        if (count == 0) return;
        width = (int)std::min<uint64_t>((uint64_t)std::max(1, width), count);
        if (width <= 1) { for (uint64_t k = 0; k < count; k++) fn(k); return; }
        while ((int)ts_.size() < width - 1) ts_.emplace_back([this] { loop(); });
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn; count_ = count; next_.store(0); failed_.store(false); active_ = width - 1; wanted_ = width - 1; gen_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [this] { return active_ == 0; });
        fn_ = nullptr;
        if (failed_.load()) throw first_;
    }
private:
    void work() {
        for (;;) {
            const uint64_t k = next_.fetch_add(1); if (k >= count_ || failed_.load()) return;
            auto fail = [&](const Fail& e) { std::lock_guard<std::mutex> g(m_); if (!failed_.exchange(true)) first_ = e; };
            try { (*fn_)(k); }
            catch (const Fail& e) { fail(e); return; }
            catch (const std::bad_alloc&) { fail(Fail{TF_E_RETRY_OOM, "host allocation failed"}); return; }       // (an exception leaving a worker thread would end the process)
            catch (const std::exception& e) { fail(Fail{TF_E_FATAL_ARG, e.what()}); return; }
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || (gen_ != seen && wanted_ > 0); });
                if (stop_) return;
                seen = gen_; wanted_--;
            }
            work();
            { std::lock_guard<std::mutex> g(m_); if (--active_ == 0) done_cv_.notify_one(); }
        }
    }
    std::vector<std::thread> ts_; std::mutex m_; std::condition_variable cv_, done_cv_;
    const std::function<void(uint64_t)>* fn_ = nullptr; uint64_t count_ = 0, gen_ = 0; int active_ = 0, wanted_ = 0; bool stop_ = false;
    std::atomic<uint64_t> next_{0}; std::atomic<bool> failed_{false}; Fail first_{0, ""};
};

struct PoolWorkers { Workers w; };
Workers& workers_of(tfgpu_columnar* pool);

template <class F> void parallel_chunks(tfgpu_columnar* pool, uint64_t n, uint64_t chunk, int threads, F&& f) {
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    const std::function<void(uint64_t)> fn = std::forward<F>(f);
    workers_of(pool).run(nchunks, threads, fn);
}

Workers& workers_of(tfgpu_columnar* pool) { if (!pool->workers) pool->workers = new Workers(); return *(Workers*)pool->workers; }

// Columnar image of `n` value lists. get(j, at, end, sparse, nvals): where row j's values start / end.
// which = nullptr: every schema column; else only the listed columns are laid out (OldKeys), the others stay empty.
struct Transposer {
    tfgpu_columnar* pool; const std::vector<int>& tfs; uint64_t n; int threads;
    std::vector<ColPlan> cp; uint64_t chunk = 2048, nchunks = 0;      // rows per task: a multiple of 8 (validity bytes), small enough to keep 64 workers busy on 100 k-row batches
    std::vector<ChunkStat> stats;      // [chunk][col]
    std::vector<uint64_t> heap_base;   // [chunk][col]

    template <class Get, class On> void walk_row(Get& get, uint64_t j, bool keyed, On&& on) {
        const uint8_t *at, *end; bool sparse; uint32_t nvals;
        get(j, at, end, sparse, nvals);
        const uint32_t nc = (uint32_t)tfs.size();
        if (!sparse && !keyed && nvals != nc) throw Fail{TF_E_FATAL_ARG, "an item's value count differs from its table schema (set TF_ITEM_SPARSE for a column subset)"};
        for (uint32_t k = 0; k < nvals; k++) {
            uint32_t c = k;
            if (sparse || keyed) { if (end - at < 2) throw Fail{TF_E_FATAL_ARG, "truncated value image"}; uint16_t ci; std::memcpy(&ci, at, 2); at += 2; c = ci; }
            if (c >= nc) throw Fail{TF_E_FATAL_ARG, "column index outside the table schema"};
            Val v; if (!read_val(at, end, v)) throw Fail{TF_E_FATAL_ARG, "malformed value image"};
            on(c, v);
        }
    }

    // The strict path (every value carries the canonical Go type of its column, type_checkers.go:39-84, or is nil — what a source that honours
    // the contract sends): ONE decode of the row image. Fixed-width values go straight to their columns; of a var-width cell the pass only
    // notes where it lies in the image and how long it is, and a second pass copies the cells column by column (sequential writes, no
    // decoding). A value of another type makes the call start over on the general two-pass path below, which lays out loose columns.
    template <class Get> bool run_strict(Get get, std::vector<tf_col>& out_cols, const uint8_t* image, uint64_t image_len) {
        const uint32_t nc = (uint32_t)tfs.size();
        if (n == 0 || image_len >= (1ull << 32) || std::getenv("TFGPU_TRANSPOSE_GENERAL")) return false;
        const bool trace = std::getenv("TFGPU_TRANSPOSE_TRACE") != nullptr; const auto t_start = std::chrono::steady_clock::now();
        if (threads > 1 && n / ((uint64_t)threads * 4) < chunk) chunk = std::max<uint64_t>(256, n / ((uint64_t)threads * 4) & ~7ull);
        nchunks = (n + chunk - 1) / chunk;
        struct SCol { uint8_t want = 0, also = 0, w = 0; bool time = false; uint8_t *values = nullptr, *validity = nullptr, *aux = nullptr; uint32_t vi = 0; };
        std::vector<SCol> sc(nc); uint32_t nvar = 0;
        for (uint32_t c = 0; c < nc; c++) {
            SCol& x = sc[c]; x.want = x.also = (uint8_t)canonical_tag(tfs[c]); x.w = (uint8_t)fixed_width(tfs[c]); x.time = is_time(tfs[c]);
            if (tfs[c] == TF_UTF8) x.also = TF_V_BYTES; else if (tfs[c] == TF_BYTES) x.also = TF_V_STRING;          // a text cell is its bytes either way (M_TEXT below)
            else if (tfs[c] == TF_ANY) { x.also = TF_V_STRING; x.aux = pool->take()->ensure(n + 16, pool->want_pinned); }   // a raw Go string inside `any`: the cell's aux tag
            if (x.w) { x.values = pool->take()->ensure((size_t)x.w * n + 16, pool->want_pinned); if (x.time) x.aux = pool->take()->ensure(4 * n + 16, pool->want_pinned); }
            else x.vi = nvar++;
            x.validity = pool->take()->ensure((n + 7) / 8 + 16, pool->want_pinned);
        }
        // where every var-width cell lies in the image and how long it is: (offset, length) per (row, var column), row-major — written and
        // read back as one stream
        uint32_t* cells = (uint32_t*)pool->take_tmp((size_t)8 * nvar * n + 16);
        stats.assign((size_t)nchunks * nc, ChunkStat());
        std::atomic<bool> loose{false};
        parallel_chunks(pool, n, chunk, threads, [&](uint64_t k) {
            if (loose.load(std::memory_order_relaxed)) return;
            const uint64_t r0 = k * chunk, r1 = std::min(n, (k + 1) * chunk);
            ChunkStat* st = &stats[(size_t)k * nc];
            for (uint32_t c = 0; c < nc; c++) {                  // a row without a value for the column: zero value, zero length, validity bit cleared below
                SCol& x = sc[c];
                std::memset(x.validity + r0 / 8, 0xff, (r1 - r0 + 7) / 8);
                if (x.values) std::memset(x.values + (size_t)x.w * r0, 0, (size_t)x.w * (r1 - r0));
                if (x.aux) { const size_t aw = x.time ? 4 : 1; std::memset(x.aux + aw * r0, 0, aw * (r1 - r0)); }
            }
            std::memset(cells + (size_t)2 * nvar * r0, 0, (size_t)8 * nvar * (r1 - r0));
            std::vector<uint8_t> seen(nc);
            for (uint64_t j = r0; j < r1; j++) {
                const uint8_t *at, *end; bool sparse; uint32_t nvals; get(j, at, end, sparse, nvals);
                if (!sparse && nvals != nc) throw Fail{TF_E_FATAL_ARG, "an item's value count differs from its table schema (set TF_ITEM_SPARSE for a column subset)"};
                if (sparse) std::fill(seen.begin(), seen.end(), 0);
                const uint8_t vbit = (uint8_t)(1u << (j & 7)); const uint64_t vbyte = j >> 3;
                for (uint32_t v = 0; v < nvals; v++) {
                    uint32_t c = v;
                    if (sparse) {
                        if (end - at < 2) throw Fail{TF_E_FATAL_ARG, "truncated value image"};
                        uint16_t ci; std::memcpy(&ci, at, 2); at += 2; c = ci;
                        if (c >= nc) throw Fail{TF_E_FATAL_ARG, "column index outside the table schema"};
                        seen[c] = 1;
                    }
                    if (at >= end) throw Fail{TF_E_FATAL_ARG, "malformed value image"};
                    const uint8_t tag = *at++; SCol& x = sc[c];
                    if (tag == x.want || tag == x.also) {
                        if (x.w) {
                            const uint32_t pw = x.time ? 12u : x.w;
                            if ((size_t)(end - at) < pw) throw Fail{TF_E_FATAL_ARG, "malformed value image"};
                            uint8_t* d = x.values + (size_t)x.w * j;
                            switch (x.w) { case 1: d[0] = at[0]; break; case 2: std::memcpy(d, at, 2); break; case 4: std::memcpy(d, at, 4); break; default: std::memcpy(d, at, 8); }
                            if (x.time) { uint32_t ns; std::memcpy(&ns, at + 8, 4); if (ns) { std::memcpy(x.aux + 4 * j, &ns, 4); st[c].nsec = true; } }
                            at += pw;
                        } else {
                            if (end - at < 4) throw Fail{TF_E_FATAL_ARG, "malformed value image"};
                            uint32_t len; std::memcpy(&len, at, 4); at += 4;
                            if ((size_t)(end - at) < len) throw Fail{TF_E_FATAL_ARG, "malformed value image"};
                            uint32_t* rec = cells + ((size_t)j * nvar + x.vi) * 2; rec[0] = (uint32_t)(at - image); rec[1] = len; st[c].heap += len; if (len > st[c].max_len) st[c].max_len = len;
                            at += len;
                            if (x.aux && tag == TF_V_STRING) { x.aux[j] = 1; st[c].nsec = true; }
                            else if (tag != x.want) st[c].tagmask = 1;                                     // the column's other text type was seen
                        }
                    } else if (tag == TF_V_NIL) { x.validity[vbyte] &= (uint8_t)~vbit; st[c].nil = true; }
                    else { if (trace && !loose.exchange(true)) fprintf(stderr, "[transpose strict] column %u (type %d) holds a value of tag %d: general path\n", c, tfs[c], (int)tag); loose.store(true, std::memory_order_relaxed); return; }
                }
                if (sparse) for (uint32_t c = 0; c < nc; c++) if (!seen[c]) { sc[c].validity[vbyte] &= (uint8_t)~vbit; st[c].nil = true; }   // absent = nil in the columnar view
            }
            if (r1 == n && (n & 7)) for (uint32_t c = 0; c < nc; c++) sc[c].validity[n >> 3] &= (uint8_t)((1u << (n & 7)) - 1);   // no stray bits behind the last row
        });
        if (loose.load()) return false;
        const auto t_p1 = std::chrono::steady_clock::now();
        // ---- layout
        cp.assign(nc, ColPlan()); heap_base.assign((size_t)nchunks * nc, 0); out_cols.assign(nc, tf_col{}); pool->text_mixed.assign(nc, 0);
        for (uint32_t c = 0; c < nc; c++) {
            ColPlan& p = cp[c]; SCol& x = sc[c]; tf_col& o = out_cols[c];
            p.schema_tf = p.phys_tf = tfs[c]; p.width = x.w; o.type = tfs[c];
            for (uint64_t k = 0; k < nchunks; k++) {
                const ChunkStat& s = stats[(size_t)k * nc + c];
                heap_base[(size_t)k * nc + c] = p.heap_total;
                p.heap_total += s.heap; p.max_len = std::max(p.max_len, s.max_len); p.has_nil |= s.nil; p.has_nsec |= s.nsec; p.tagmask |= s.tagmask;
            }
            pool->text_mixed[c] = p.tagmask != 0;
            if (p.has_nil) o.validity = x.validity;
            if (x.w) { o.values = x.values; if (p.has_nsec) o.aux = x.aux; continue; }
            if (x.aux && p.has_nsec) { p.has_anytag = true; p.has_nsec = false; o.aux = x.aux; }
            if (p.heap_total >= (1ull << 32)) throw Fail{TF_E_FATAL_UNSUPPORTED, "column " + std::to_string(c) + " (" + tfplan::tf_to_yt(tfs[c]) + "): heap over 4 GiB"};
            p.lens_width = p.max_len < 256 ? 1 : p.max_len < 65536 ? 2 : 4;
            p.lens = pool->take()->ensure((size_t)p.lens_width * (n + 1) + 16, pool->want_pinned);
            p.heap = pool->take()->ensure((size_t)p.heap_total + 16, pool->want_pinned);
            o.offsets = (const uint32_t*)p.lens; o.heap = p.heap; o.heap_len = p.heap_total;
            o.flags = p.lens_width == 1 ? TF_COL_LENS8 : p.lens_width == 2 ? TF_COL_LENS16 : 0;
        }
        // ---- pass 2: the var-width cells, row by row (a row's cells lie together in the image), one running heap offset per column
        std::vector<uint32_t> var_cols; for (uint32_t c = 0; c < nc; c++) if (!cp[c].width) var_cols.push_back(c);
        parallel_chunks(pool, n, chunk, threads, [&](uint64_t k) {
            const uint64_t r0 = k * chunk, r1 = std::min(n, (k + 1) * chunk); const size_t nv = var_cols.size();
            std::vector<uint64_t> off(nv), hend(nv);
            for (size_t i = 0; i < nv; i++) { const uint32_t c = var_cols[i]; off[i] = heap_base[(size_t)k * nc + c]; hend[i] = k + 1 < nchunks ? heap_base[(size_t)(k + 1) * nc + c] : cp[c].heap_total; }
            for (uint64_t j = r0; j < r1; j++) {
                const uint32_t* rec = cells + (size_t)j * nv * 2;
                for (size_t i = 0; i < nv; i++) {
                    const ColPlan& p = cp[var_cols[i]];
                    const uint32_t at = rec[2 * i], len = rec[2 * i + 1]; const uint8_t* src = image + at; const uint64_t o = off[i];
                    if (p.lens_width == 1) p.lens[j] = (uint8_t)len;
                    else if (p.lens_width == 2) { const uint16_t l = (uint16_t)len; std::memcpy(p.lens + 2 * j, &l, 2); }
                    else { const uint32_t o32 = (uint32_t)o; std::memcpy(p.lens + 4 * j, &o32, 4); }
                    if (len <= 16 && o + 16 <= hend[i] && (uint64_t)at + 16 <= image_len) std::memcpy(p.heap + o, src, 16);   // one 16-byte move inside this chunk's share; the next cell overwrites the excess
                    else std::memcpy(p.heap + o, src, len);
                    off[i] = o + len;
                }
            }
            if (r1 == n) for (uint32_t c : var_cols) if (cp[c].lens_width == 4) { const uint32_t tot = (uint32_t)cp[c].heap_total; std::memcpy(cp[c].lens + 4 * n, &tot, 4); }
        });
        if (trace) {
            const auto t_end = std::chrono::steady_clock::now(); auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "[transpose strict] n=%llu decode %.2f ms, cells %.2f ms\n", (unsigned long long)n, ms(t_start, t_p1), ms(t_p1, t_end));
        }
        return true;
    }

    template <class Get> void run(Get get, bool keyed, std::vector<tf_col>& out_cols, const uint8_t* only /* ncols flags or nullptr */) {
        const uint32_t nc = (uint32_t)tfs.size(); const auto t_start = std::chrono::steady_clock::now();
        cp.assign(nc, ColPlan()); for (uint32_t c = 0; c < nc; c++) cp[c].schema_tf = tfs[c];
        // four tasks per worker at least (98 tasks of 2048 rows leave a third of 64 workers idle in the second round of a 200 k-row batch), never
        // below 256 rows (the per-task setup is per column), a multiple of 8 (validity bytes)
        if (threads > 1 && n / ((uint64_t)threads * 4) < chunk) chunk = std::max<uint64_t>(256, n / ((uint64_t)threads * 4) & ~7ull);
        nchunks = (n + chunk - 1) / chunk;
        stats.assign((size_t)nchunks * nc, ChunkStat());
        if (!keyed && !only) pool->text_mixed.assign(nc, 0);
        // ---- pass 1: tags, heap bytes, longest cell per (chunk, column)
        std::vector<uint8_t> fw(nc); for (uint32_t c = 0; c < nc; c++) fw[c] = (uint8_t)fixed_width(tfs[c]);
        parallel_chunks(pool, n, chunk, threads, [&](uint64_t k) {
            ChunkStat* st = &stats[(size_t)k * nc];
            std::vector<uint8_t> seen(nc);
            for (uint64_t j = k * chunk; j < std::min(n, (k + 1) * chunk); j++) {
                const uint8_t *at0, *end0; bool sparse0; uint32_t nv0; get(j, at0, end0, sparse0, nv0);
                const bool track = sparse0 || keyed;                  // a dense row lists every column: nothing can be absent
                if (track) std::fill(seen.begin(), seen.end(), 0);
                walk_row(get, j, keyed, [&](uint32_t c, const Val& v) {
                    if (track) seen[c] = 1;
                    ChunkStat& s = st[c];
                    if (v.tag == TF_V_NIL) { s.nil = true; return; }
                    s.tagmask |= 1u << v.tag;
                    if (!fw[c]) {
                        uint32_t len = v.n;
                        if (tfs[c] == TF_ANY && payload_fixed(v.tag) != 0xffffffffu) { char tmp[32]; len = any_scalar_text(v, tmp); }
                        s.heap += len; if (len > s.max_len) s.max_len = len;
                    } else if (v.tag == TF_V_TIME) { uint32_t ns; std::memcpy(&ns, v.p + 8, 4); if (ns) s.nsec = true; }
                });
                if (track) for (uint32_t c = 0; c < nc; c++) if (!seen[c] && (!only || only[c])) st[c].nil = true;     // absent = nil in the columnar view
            }
        });
        const bool trace = std::getenv("TFGPU_TRANSPOSE_TRACE") != nullptr; const auto t_p1 = std::chrono::steady_clock::now();
        // ---- layout decisions per column
        heap_base.assign((size_t)nchunks * nc, 0);
        for (uint32_t c = 0; c < nc; c++) {
            ColPlan& p = cp[c];
            if (only && !only[c]) continue;
            for (uint64_t k = 0; k < nchunks; k++) {
                const ChunkStat& s = stats[(size_t)k * nc + c];
                heap_base[(size_t)k * nc + c] = p.heap_total;
                p.tagmask |= s.tagmask; p.heap_total += s.heap; p.max_len = std::max(p.max_len, s.max_len); p.has_nil |= s.nil; p.has_nsec |= s.nsec;
            }
            const int tf = p.schema_tf; const uint32_t m = p.tagmask;
            auto refuse = [&](const char* why) { throw Fail{TF_E_FATAL_UNSUPPORTED, "column " + std::to_string(c) + " (" + tfplan::tf_to_yt(tf) + "): " + why}; };
            if (fixed_width(tf)) {
                if (m == 0 || m == (1u << canonical_tag(tf)) || (tf == TF_DOUBLE && m == (1u << TF_V_JSONNUM))) p.phys_tf = tf;
                else if (is_time(tf) && (m & (1u << TF_V_TIME))) refuse("time.Time mixed with other value types");
                else if (m & ~(M_SIGNED | M_UNSIGNED | M_FLOAT | (1u << TF_V_DURATION))) refuse("a text / JSON value in a fixed-width column stays on the Go path (strictify casts of strings)");
                else if ((m & (m - 1)) == 0) { int t = 0; while (!(m >> t & 1)) t++; p.phys_tf = t == TF_V_DURATION ? TF_INT64 : tag_tf(t); }   // one foreign type: a loose column of that type
                else if (!(m & ~M_SIGNED) || !(m & ~(M_SIGNED | (1u << TF_V_DURATION)))) p.phys_tf = TF_INT64;
                else if (!(m & ~M_UNSIGNED)) p.phys_tf = TF_UINT64;
                else if (!(m & ~M_FLOAT)) p.phys_tf = TF_DOUBLE;
                else refuse("signed, unsigned and float values mixed in one column");
                p.width = fixed_width(p.phys_tf);
            } else {
                if (tf != TF_ANY && (m & ~M_TEXT)) refuse("a non-text value in a string column");
                if (tf == TF_ANY && (m & (M_FLOAT | (1u << TF_V_TIME) | (1u << TF_V_DURATION) | (1u << TF_V_BYTES)))) refuse("float / time / []byte inside `any` needs encoding/json's formatter");
                if (p.heap_total >= (1ull << 32)) refuse("heap over 4 GiB");
                p.phys_tf = tf; p.lens_width = p.max_len < 256 ? 1 : p.max_len < 65536 ? 2 : 4;
                if (!keyed && !only) pool->text_mixed[c] = tf != TF_ANY && (m & (1u << (tf == TF_UTF8 ? TF_V_BYTES : TF_V_STRING))) != 0;
                p.has_anytag = tf == TF_ANY && (m & (1u << TF_V_STRING));
            }
        }
        // ---- buffers
        out_cols.assign(nc, tf_col{});
        for (uint32_t c = 0; c < nc; c++) {
            ColPlan& p = cp[c]; tf_col& o = out_cols[c];
            o.type = p.phys_tf ? p.phys_tf : p.schema_tf;
            if (only && !only[c]) continue;
            if (p.width) { p.values = pool->take()->ensure((size_t)p.width * n + 16, pool->want_pinned); o.values = p.values; }
            else {
                p.lens = pool->take()->ensure((size_t)p.lens_width * (n + 1) + 16, pool->want_pinned);
                p.heap = pool->take()->ensure((size_t)p.heap_total + 16, pool->want_pinned);
                o.offsets = (const uint32_t*)p.lens; o.heap = p.heap; o.heap_len = p.heap_total;
                o.flags = p.lens_width == 1 ? TF_COL_LENS8 : p.lens_width == 2 ? TF_COL_LENS16 : 0;
            }
            if (p.has_nil) { p.validity = pool->take()->ensure((n + 7) / 8 + 16, pool->want_pinned); o.validity = p.validity; }
            if (p.has_nsec) { p.aux = pool->take()->ensure(4 * n + 16, pool->want_pinned); o.aux = p.aux; }
            if (p.has_anytag) { p.aux = pool->take()->ensure(n + 16, pool->want_pinned); o.aux = p.aux; }
        }
        std::vector<uint32_t> wide; for (uint32_t c = 0; c < nc; c++) if (cp[c].lens && cp[c].lens_width == 4) wide.push_back(c);
        const auto t_buf = std::chrono::steady_clock::now();
        // ---- pass 2: fill (chunks are multiples of 8 rows, so validity bytes never straddle two workers)
        parallel_chunks(pool, n, chunk, threads, [&](uint64_t k) {
            const uint64_t r0 = k * chunk, r1 = std::min(n, (k + 1) * chunk);
            std::vector<uint64_t> hp(nc), hend(nc);
            for (uint32_t c = 0; c < nc; c++) { hp[c] = heap_base[(size_t)k * nc + c]; hend[c] = k + 1 < nchunks ? heap_base[(size_t)(k + 1) * nc + c] : cp[c].heap_total; }
            for (uint32_t c = 0; c < nc; c++) {                  // defaults: nil rows keep zero values / zero lengths
                ColPlan& p = cp[c]; if (only && !only[c]) continue;
                if (p.values) std::memset(p.values + (size_t)p.width * r0, 0, (size_t)p.width * (r1 - r0));
                if (p.validity) std::memset(p.validity + r0 / 8, 0, (r1 - r0 + 7) / 8);
                if (p.aux) std::memset(p.aux + (p.has_nsec ? 4 : 1) * r0, 0, (p.has_nsec ? 4 : 1) * (r1 - r0));
                if (p.lens && p.lens_width != 4) std::memset(p.lens + (size_t)p.lens_width * r0, 0, (size_t)p.lens_width * (r1 - r0));
            }
            for (uint64_t j = r0; j < r1; j++) {
                for (uint32_t c : wide) { const uint32_t o = (uint32_t)hp[c]; std::memcpy(cp[c].lens + 4 * j, &o, 4); }   // u32 offsets: a row without a value repeats the running offset
                const uint8_t *at0, *vend_all; bool sparse0; uint32_t nv0; get(j, at0, vend_all, sparse0, nv0);
                walk_row(get, j, keyed, [&](uint32_t c, const Val& v) {
                    ColPlan& p = cp[c];
                    if (v.tag == TF_V_NIL) return;
                    if (p.validity) p.validity[j >> 3] |= (uint8_t)(1u << (j & 7));
                    if (p.width) {
                        uint8_t* d = p.values + (size_t)p.width * j;
                        if (v.tag == TF_V_TIME) { std::memcpy(d, v.p, 8); if (p.aux) std::memcpy(p.aux + 4 * j, v.p + 8, 4); }
                        else if (v.tag == TF_V_JSONNUM) { const std::string t((const char*)v.p, v.n); const double x = std::strtod(t.c_str(), nullptr); std::memcpy(d, &x, 8); }
                        else if (tag_tf(v.tag) == p.phys_tf || (v.tag == TF_V_DURATION && p.width == 8)) {
                            switch (p.width) { case 1: d[0] = v.p[0]; break; case 2: std::memcpy(d, v.p, 2); break; case 4: std::memcpy(d, v.p, 4); break; default: std::memcpy(d, v.p, 8); }
                        }
                        else if (p.phys_tf == TF_DOUBLE) { float f; std::memcpy(&f, v.p, 4); const double x = f; std::memcpy(d, &x, 8); }
                        else { const int64_t x = val_i64(v); std::memcpy(d, &x, 8); }                     // widened into INT64 / UINT64
                    } else {
                        uint32_t len = v.n; const uint8_t* src = v.p; char tmp[32];
                        if (p.schema_tf == TF_ANY && payload_fixed(v.tag) != 0xffffffffu) { len = any_scalar_text(v, tmp); src = (const uint8_t*)tmp; }
                        if (len <= 16 && hp[c] + 16 <= hend[c] && (src == (const uint8_t*)tmp || src + 16 <= vend_all)) std::memcpy(p.heap + hp[c], src, 16);   // one 16-byte move inside this chunk's share of the heap; the cells behind it overwrite the excess
                        else std::memcpy(p.heap + hp[c], src, len);
                        if (p.lens_width == 1) p.lens[j] = (uint8_t)len;
                        else if (p.lens_width == 2) { const uint16_t l = (uint16_t)len; std::memcpy(p.lens + 2 * j, &l, 2); }
                        hp[c] += len;
                        if (p.has_anytag && v.tag == TF_V_STRING) p.aux[j] = 1;
                    }
                });
            }
            if (r1 == n) for (uint32_t c : wide) { const uint32_t tot = (uint32_t)cp[c].heap_total; std::memcpy(cp[c].lens + 4 * n, &tot, 4); }
        });
        if (trace) {
            const auto t_end = std::chrono::steady_clock::now(); auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "[transpose] n=%llu pass1 %.2f ms, layout+buffers %.2f ms, pass2 %.2f ms\n", (unsigned long long)n, ms(t_start, t_p1), ms(t_p1, t_buf), ms(t_buf, t_end));
        }
    }
};

const std::vector<int>& schema_types(tfgpu_columnar* p, const char* schema_json) {
    auto it = p->schemas.find(schema_json);
    if (it != p->schemas.end()) return it->second;
    std::vector<int> tfs;
    try { for (auto& c : tfplan::parse_schema(schema_json)) tfs.push_back(c.tf); }
    catch (const tfplan::FatalError& f) { throw Fail{f.code, f.what()}; }
    catch (const std::exception& e) { throw Fail{TF_E_FATAL_CONFIG, e.what()}; }
    return p->schemas.emplace(schema_json, std::move(tfs)).first->second;
}

}  // namespace

tfgpu_columnar::~tfgpu_columnar() { for (auto& b : bufs) b.release(); for (auto& b : tmp_bufs) b.release(); delete (Workers*)workers; }

bool tfgpu_columnar_text_was_mixed(const tfgpu_columnar* pool, uint32_t col) { return pool && col < pool->text_mixed.size() && pool->text_mixed[col]; }

int tfgpu_columnar_rewrite_text(tfgpu_columnar* pool, uint32_t col, const std::function<tf_text_fn()>& make, int threads) {
    if (!pool || col >= pool->cols.size()) return TF_E_FATAL_ARG;
    try {
        tf_col& o = pool->cols[col]; const uint64_t n = pool->batch.nrows;
        if (o.values || (!o.offsets && n)) throw Fail{TF_E_FATAL_ARG, "not a var-width column"};
        if (!n) return TF_OK;
        if (threads <= 0) threads = (int)std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency()));
        const int lw = o.flags & TF_COL_LENS8 ? 1 : o.flags & TF_COL_LENS16 ? 2 : 4;
        const uint8_t* lens = (const uint8_t*)o.offsets;
        auto cell_len = [&](uint64_t j) -> uint32_t { if (lw == 1) return lens[j]; if (lw == 2) { uint16_t l; std::memcpy(&l, lens + 2 * j, 2); return l; } uint32_t a, b; std::memcpy(&a, lens + 4 * j, 4); std::memcpy(&b, lens + 4 * j + 4, 4); return b - a; };
        const uint64_t chunk = std::max<uint64_t>(512, std::min<uint64_t>(8192, n / ((uint64_t)threads * 4) & ~7ull));
        const uint64_t nchunks = (n + chunk - 1) / chunk;
        std::vector<uint64_t> in_base(nchunks + 1, 0);                       // where every task's cells start in the old heap
        for (uint64_t k = 0; k < nchunks; k++) { uint64_t sum = 0; for (uint64_t j = k * chunk; j < std::min(n, (k + 1) * chunk); j++) sum += cell_len(j); in_base[k + 1] = in_base[k] + sum; }
        struct Out { std::string heap; std::vector<uint32_t> len; uint32_t max_len = 0; };
        std::vector<Out> outs(nchunks);
        parallel_chunks(pool, n, chunk, threads, [&](uint64_t k) {
            const tf_text_fn fn = make(); Out& w = outs[k]; std::string cell;
            const uint64_t r0 = k * chunk, r1 = std::min(n, (k + 1) * chunk); uint64_t at = in_base[k];
            w.len.resize(r1 - r0); w.heap.reserve((size_t)(in_base[k + 1] - in_base[k]) + 64);
            for (uint64_t j = r0; j < r1; j++) {
                const uint32_t len = cell_len(j);
                if (o.validity && !(o.validity[j >> 3] >> (j & 7) & 1)) { w.len[j - r0] = 0; at += len; continue; }      // nil stays nil
                fn(o.heap + at, len, cell); at += len;
                if (cell.size() >= (1ull << 32)) throw Fail{TF_E_FATAL_UNSUPPORTED, "a replaced cell over 4 GiB"};
                w.len[j - r0] = (uint32_t)cell.size(); w.max_len = std::max(w.max_len, (uint32_t)cell.size()); w.heap += cell;
            }
        });
        uint64_t total = 0; uint32_t max_len = 0; std::vector<uint64_t> out_base(nchunks);
        for (uint64_t k = 0; k < nchunks; k++) { out_base[k] = total; total += outs[k].heap.size(); max_len = std::max(max_len, outs[k].max_len); }
        if (total >= (1ull << 32)) throw Fail{TF_E_FATAL_UNSUPPORTED, "column " + std::to_string(col) + ": heap over 4 GiB"};
        const int nlw = max_len < 256 ? 1 : max_len < 65536 ? 2 : 4;
        uint8_t* nl = pool->take()->ensure((size_t)nlw * (n + 1) + 16, pool->want_pinned);
        uint8_t* nh = pool->take()->ensure((size_t)total + 16, pool->want_pinned);
        parallel_chunks(pool, n, chunk, threads, [&](uint64_t k) {
            const Out& w = outs[k]; const uint64_t r0 = k * chunk; uint64_t off = out_base[k];
            std::memcpy(nh + off, w.heap.data(), w.heap.size());
            for (size_t i = 0; i < w.len.size(); i++) {
                const uint64_t j = r0 + i;
                if (nlw == 1) nl[j] = (uint8_t)w.len[i];
                else if (nlw == 2) { const uint16_t l = (uint16_t)w.len[i]; std::memcpy(nl + 2 * j, &l, 2); }
                else { const uint32_t o32 = (uint32_t)off; std::memcpy(nl + 4 * j, &o32, 4); }
                off += w.len[i];
            }
        });
        if (nlw == 4) { const uint32_t t32 = (uint32_t)total; std::memcpy(nl + 4 * n, &t32, 4); }
        o.offsets = (const uint32_t*)nl; o.heap = nh; o.heap_len = total; o.flags = nlw == 1 ? TF_COL_LENS8 : nlw == 2 ? TF_COL_LENS16 : 0;
        return TF_OK;
    } catch (const Fail& f) { pool->err = f.msg; return f.rc; }
    catch (const std::bad_alloc&) { pool->err = "host allocation failed"; return TF_E_RETRY_OOM; }
}

extern "C" {

int tfgpu_columnar_create(tfgpu_columnar** out) {
    if (!out) return TF_E_FATAL_ARG;
    auto* p = new tfgpu_columnar();
    int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { (void)cudaGetLastError(); p->want_pinned = false; }   // pageable memory without a device: the layout is the same
    *out = p; return TF_OK;
}
int tfgpu_columnar_destroy(tfgpu_columnar* p) { if (!p) return TF_E_FATAL_ARG; delete p; return TF_OK; }
const char* tfgpu_columnar_last_error(const tfgpu_columnar* p) { return p ? p->err.c_str() : "null pool"; }

int tfgpu_rows_to_batch(tfgpu_columnar* pool, const tf_rows* rows, uint32_t table, const uint64_t* item_idx, uint64_t n, int threads,
                        const tf_batch** batch, const tf_row_meta** meta, const tf_old_keys** old) {
    if (!pool || !rows || !batch) return TF_E_FATAL_ARG;
    try {
        if (table >= rows->n_tables) throw Fail{TF_E_FATAL_ARG, "table index out of range"};
        const std::vector<int>& tfs = schema_types(pool, rows->tables[table].schema_json);
        std::vector<uint64_t> all;
        if (!item_idx) {
            for (uint64_t i = 0; i < rows->n_items; i++) if (rows->items[i].table == table && TF_KIND_IS_ROW(rows->items[i].kind)) all.push_back(i);
            item_idx = all.data(); n = all.size();
        }
        for (uint64_t j = 0; j < n; j++) {
            if (item_idx[j] >= rows->n_items) throw Fail{TF_E_FATAL_ARG, "item index out of range"};
            const tf_item& it = rows->items[item_idx[j]];
            if (it.table != table || !TF_KIND_IS_ROW(it.kind)) throw Fail{TF_E_FATAL_ARG, "tfgpu_rows_to_batch takes the row events of ONE table"};
            if (it.values_off > rows->values_len) throw Fail{TF_E_FATAL_ARG, "values offset outside the image"};
        }
        if (threads <= 0) threads = (int)std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency()));
        pool->next_buf = 0;
        const uint8_t* vend = rows->values + rows->values_len;
        auto value_list = [&](uint64_t j, const uint8_t*& at, const uint8_t*& end, bool& sparse, uint32_t& nvals) {
            const tf_item& it = rows->items[item_idx[j]];
            at = rows->values + it.values_off; end = vend; sparse = it.flags & TF_ITEM_SPARSE; nvals = it.n_values;
        };
        pool->next_tmp = 0;
        bool strict;
        { Transposer ts{pool, tfs, n, threads}; strict = ts.run_strict(value_list, pool->cols, rows->values, rows->values_len); }
        if (!strict) { pool->next_buf = 0; Transposer tr{pool, tfs, n, threads}; tr.run(value_list, false, pool->cols, nullptr); }
        // kinds + meta
        uint8_t* kinds = pool->take()->ensure(n + 16, pool->want_pinned);
        uint32_t* ids = (uint32_t*)pool->take()->ensure(4 * n + 16, pool->want_pinned);
        uint64_t* lsn = (uint64_t*)pool->take()->ensure(8 * n + 16, pool->want_pinned);
        uint64_t* ct = (uint64_t*)pool->take()->ensure(8 * n + 16, pool->want_pinned);
        uint32_t* tx_off = (uint32_t*)pool->take()->ensure(4 * (n + 1) + 16, pool->want_pinned);
        uint64_t tx_total = 0; bool any_noninsert = false, any_old = false;
        for (uint64_t j = 0; j < n; j++) {
            const tf_item& it = rows->items[item_idx[j]];
            kinds[j] = it.kind; ids[j] = it.id; lsn[j] = it.lsn; ct[j] = it.commit_time; tx_off[j] = (uint32_t)tx_total;
            if ((uint64_t)it.txid_off + it.txid_len > rows->strings_len) throw Fail{TF_E_FATAL_ARG, "TxID outside the strings heap"};
            tx_total += it.txid_len; any_noninsert |= it.kind != TF_KIND_INSERT; any_old |= it.old_keys_off != UINT64_MAX;
        }
        if (tx_total >= (1ull << 32)) throw Fail{TF_E_FATAL_UNSUPPORTED, "TxID heap over 4 GiB"};
        tx_off[n] = (uint32_t)tx_total;
        uint8_t* tx_heap = pool->take()->ensure(tx_total + 16, pool->want_pinned);
        for (uint64_t j = 0; j < n; j++) { const tf_item& it = rows->items[item_idx[j]]; std::memcpy(tx_heap + tx_off[j], rows->strings + it.txid_off, it.txid_len); }
        pool->batch = tf_batch{n, (uint32_t)tfs.size(), TF_MEM_HOST, pool->cols.data(), any_noninsert ? kinds : nullptr};
        pool->meta = tf_row_meta{ids, lsn, ct, tx_total ? tx_off : nullptr, tx_total ? tx_heap : nullptr};
        *batch = &pool->batch; if (meta) *meta = &pool->meta;
        // OldKeys (old_keys.go:3-7): KeyNames / KeyValues of the update and delete rows as a second batch over the same rows
        if (old) {
            *old = nullptr;
            if (any_old) {
                pool->present.assign(tfs.size(), 0);
                uint8_t* row_has = pool->take()->ensure(n + 16, pool->want_pinned);
                for (uint64_t j = 0; j < n; j++) {
                    const tf_item& it = rows->items[item_idx[j]]; row_has[j] = it.old_keys_off != UINT64_MAX;
                    if (!row_has[j]) continue;
                    if (it.old_keys_off + 2 > rows->values_len) throw Fail{TF_E_FATAL_ARG, "OldKeys offset outside the image"};
                    const uint8_t* at = rows->values + it.old_keys_off; uint16_t cnt; std::memcpy(&cnt, at, 2); at += 2;
                    for (uint16_t k = 0; k < cnt; k++) {
                        if (vend - at < 2) throw Fail{TF_E_FATAL_ARG, "truncated OldKeys image"};
                        uint16_t c; std::memcpy(&c, at, 2); at += 2; Val v;
                        if (c >= tfs.size() || !read_val(at, vend, v)) throw Fail{TF_E_FATAL_ARG, "malformed OldKeys image"};
                        pool->present[c] = 1;
                    }
                }
                Transposer to{pool, tfs, n, threads};
                static const uint8_t none[2] = {0, 0};
                to.run([&](uint64_t j, const uint8_t*& at, const uint8_t*& end, bool& sparse, uint32_t& nvals) {
                    const tf_item& it = rows->items[item_idx[j]]; sparse = true; end = vend;
                    const uint8_t* p = it.old_keys_off != UINT64_MAX ? rows->values + it.old_keys_off : none;
                    uint16_t cnt; std::memcpy(&cnt, p, 2); nvals = cnt; at = p + 2;
                }, true, pool->old_cols, pool->present.data());
                pool->old_batch = tf_batch{n, (uint32_t)tfs.size(), TF_MEM_HOST, pool->old_cols.data(), nullptr};
                pool->old = tf_old_keys{&pool->old_batch, pool->present.data(), row_has};
                *old = &pool->old;
            }
        }
        return TF_OK;
    } catch (const Fail& f) { pool->err = f.msg; return f.rc; }
    catch (const std::bad_alloc&) { pool->err = "host allocation failed"; return TF_E_RETRY_OOM; }
}

// The rows sel[0..m) of a host batch (ascending row indexes), in order, in the pool's buffers: the host half of a two-phase push
// (tfgpu_push_encode_selective) — when filter_rows keeps a fraction of the rows only that fraction has to cross PCIe. Fixed-width values,
// validity bits and aux go through the selection vector per chunk of OUTPUT rows; var-width columns per chunk of INPUT rows: narrow length
// arrays have no random access, so a task first turns its chunk's lengths into a local prefix (a tight loop) and then copies only the kept cells.
static int gather_core(tfgpu_columnar* pool, const tf_batch* in, const uint32_t* sel, uint64_t m, int threads, const tf_batch** out) {
    const uint64_t n = in->nrows; const uint32_t nc = in->ncols;
    const uint64_t CH = 32768, nch = (n + CH - 1) / CH;
    std::vector<uint64_t> kept(nch + 1, 0);                       // kept[k] = first output row of input chunk k
    for (uint64_t k = 1; k <= nch; k++) kept[k] = (uint64_t)(std::lower_bound(sel, sel + m, (uint32_t)std::min<uint64_t>(k * CH, 0xffffffffull)) - sel);
    kept[nch] = m;
    pool->cols.assign(nc, tf_col{});
    struct Var { uint32_t c; int lw; uint8_t* lens; uint8_t* heap; std::vector<uint64_t> in_base, out_base; };
    std::vector<Var> vars;
    for (uint32_t c = 0; c < nc; c++) {
        const tf_col& ic = in->cols[c]; tf_col& oc = pool->cols[c]; oc.type = ic.type; oc.flags = ic.flags;
        const int w = fixed_width(ic.type);
        if (w) { if (ic.values) oc.values = pool->take()->ensure((size_t)w * m + 16, pool->want_pinned); }
        else if (ic.offsets) {
            Var v; v.c = c; v.lw = (ic.flags & TF_COL_LENS8) ? 1 : (ic.flags & TF_COL_LENS16) ? 2 : 4;
            v.lens = pool->take()->ensure((size_t)v.lw * (m + 1) + 16, pool->want_pinned); v.heap = nullptr;
            v.in_base.assign(nch + 1, 0); v.out_base.assign(nch + 1, 0);
            oc.offsets = (const uint32_t*)v.lens; vars.push_back(std::move(v));
        }
        if (ic.validity) oc.validity = pool->take()->ensure((m + 7) / 8 + 16, pool->want_pinned);
        if (ic.aux) oc.aux = pool->take()->ensure((size_t)(is_time(ic.type) ? 4 : 1) * m + 16, pool->want_pinned);
    }
    // pass A: bytes per (var column, input chunk): all rows (where the chunk's cells start in the heap) and kept rows (where they go)
    const uint64_t nv = vars.size();
    parallel_chunks(pool, nv * nch, 1, threads, [&](uint64_t t) {
        Var& v = vars[t / nch]; const uint64_t k = t % nch; const tf_col& ic = in->cols[v.c];
        const uint64_t r0 = k * CH, r1 = std::min(n, (k + 1) * CH); uint64_t all = 0, kb = 0;
        if (v.lw == 1) { const uint8_t* l = (const uint8_t*)ic.offsets; uint32_t a = 0; for (uint64_t r = r0; r < r1; r++) a += l[r]; all = a; for (uint64_t j = kept[k]; j < kept[k + 1]; j++) kb += l[sel[j]]; }
        else if (v.lw == 2) { const uint16_t* l = (const uint16_t*)ic.offsets; for (uint64_t r = r0; r < r1; r++) all += l[r]; for (uint64_t j = kept[k]; j < kept[k + 1]; j++) kb += l[sel[j]]; }
        else { all = ic.offsets[r1] - ic.offsets[r0]; for (uint64_t j = kept[k]; j < kept[k + 1]; j++) kb += ic.offsets[sel[j] + 1] - ic.offsets[sel[j]]; }
        v.in_base[k + 1] = all; v.out_base[k + 1] = kb;
    });
    for (Var& v : vars) {
        for (uint64_t k = 0; k < nch; k++) { v.in_base[k + 1] += v.in_base[k]; v.out_base[k + 1] += v.out_base[k]; }
        v.heap = pool->take()->ensure(v.out_base[nch] + 32, pool->want_pinned);
        tf_col& oc = pool->cols[v.c]; oc.heap = v.heap; oc.heap_len = v.out_base[nch];
    }
    // pass B
    const uint64_t OC = 16384, noc = (m + OC - 1) / OC;
    parallel_chunks(pool, nv * nch + (uint64_t)nc * noc, 1, threads, [&](uint64_t t) {
        if (t < nv * nch) {
            Var& v = vars[t / nch]; const uint64_t k = t % nch; const tf_col& ic = in->cols[v.c];
            const uint64_t r0 = k * CH, r1 = std::min(n, (k + 1) * CH), j0 = kept[k], j1 = kept[k + 1];
            if (v.lw == 4 && r1 == n) { const uint32_t x = (uint32_t)v.out_base[nch]; std::memcpy(v.lens + 4 * m, &x, 4); }
            if (j0 == j1) return;
            uint64_t op = v.out_base[k]; const uint64_t op_end = v.out_base[k + 1];
            // a cell of up to 16 bytes is copied as one 16-byte move while that stays inside this task's share of the output heap
            auto put = [&](const uint8_t* src, uint32_t l) {
                if (l <= 16 && op + 16 <= op_end) std::memcpy(v.heap + op, src, 16); else std::memcpy(v.heap + op, src, l);
                op += l;
            };
            if (v.lw == 4) {
                for (uint64_t j = j0; j < j1; j++) {
                    const uint32_t r = sel[j], o = ic.offsets[r], l = ic.offsets[r + 1] - o; const uint32_t x = (uint32_t)op;
                    std::memcpy(v.lens + 4 * j, &x, 4);
                    if ((uint64_t)o + 16 <= ic.heap_len) put(ic.heap + o, l); else { std::memcpy(v.heap + op, ic.heap + o, l); op += l; }
                }
                return;
            }
            thread_local std::vector<uint32_t> pre; pre.resize(CH + 1);
            const uint64_t base = v.in_base[k]; uint32_t acc = 0;
            if (v.lw == 1) { const uint8_t* l = (const uint8_t*)ic.offsets + r0; for (uint64_t i = 0; i < r1 - r0; i++) { pre[i] = acc; acc += l[i]; } }
            else { const uint16_t* l = (const uint16_t*)ic.offsets + r0; for (uint64_t i = 0; i < r1 - r0; i++) { pre[i] = acc; acc += l[i]; } }
            // the heap of a narrow column may hold trailing bytes past the last cell read 16 at a time: heap_len bounds the fast path
            const uint64_t in_end = ic.heap_len;
            for (uint64_t j = j0; j < j1; j++) {
                const uint32_t r = sel[j]; const uint64_t ip = base + pre[r - r0];
                const uint32_t l = v.lw == 1 ? ((const uint8_t*)ic.offsets)[r] : ((const uint16_t*)ic.offsets)[r];
                if (v.lw == 1) v.lens[j] = (uint8_t)l; else { const uint16_t x = (uint16_t)l; std::memcpy(v.lens + 2 * j, &x, 2); }
                if (ip + 16 <= in_end) put(ic.heap + ip, l); else { std::memcpy(v.heap + op, ic.heap + ip, l); op += l; }
            }
            return;
        }
        const uint64_t u = t - nv * nch; const uint32_t c = (uint32_t)(u / noc); const uint64_t j0 = (u % noc) * OC, j1 = std::min(m, j0 + OC);
        const tf_col& ic = in->cols[c]; tf_col& oc = pool->cols[c];
        const int w = fixed_width(ic.type);
        if (w && ic.values) {
            uint8_t* o = (uint8_t*)oc.values; const uint8_t* s = (const uint8_t*)ic.values;
            switch (w) {
            case 1: for (uint64_t j = j0; j < j1; j++) o[j] = s[sel[j]]; break;
            case 2: for (uint64_t j = j0; j < j1; j++) ((uint16_t*)o)[j] = ((const uint16_t*)s)[sel[j]]; break;
            case 4: for (uint64_t j = j0; j < j1; j++) ((uint32_t*)o)[j] = ((const uint32_t*)s)[sel[j]]; break;
            default: for (uint64_t j = j0; j < j1; j++) ((uint64_t*)o)[j] = ((const uint64_t*)s)[sel[j]]; break;
            }
        }
        if (ic.validity) {
            uint8_t* o = (uint8_t*)oc.validity;
            for (uint64_t j = j0; j < j1; j += 8) { uint8_t b = 0; for (uint64_t q = j; q < std::min(j1, j + 8); q++) { const uint32_t r = sel[q]; b |= (uint8_t)(((ic.validity[r >> 3] >> (r & 7)) & 1) << (q - j)); } o[j >> 3] = b; }
        }
        if (ic.aux) {
            if (is_time(ic.type)) for (uint64_t j = j0; j < j1; j++) ((uint32_t*)oc.aux)[j] = ((const uint32_t*)ic.aux)[sel[j]];
            else for (uint64_t j = j0; j < j1; j++) ((uint8_t*)oc.aux)[j] = ((const uint8_t*)ic.aux)[sel[j]];
        }
    });
    uint8_t* kinds = nullptr;
    if (in->kinds) { kinds = pool->take()->ensure(m + 16, pool->want_pinned); for (uint64_t j = 0; j < m; j++) kinds[j] = in->kinds[sel[j]]; }
    for (uint32_t c = 0; c < nc; c++) if (!fixed_width(in->cols[c].type) && in->cols[c].offsets && !pool->cols[c].heap) pool->cols[c].heap = (const uint8_t*)sel;   // empty heap: any valid pointer
    pool->batch = tf_batch{m, nc, TF_MEM_HOST, pool->cols.data(), kinds};
    *out = &pool->batch;
    return TF_OK;
}

int tfgpu_batch_gather(tfgpu_columnar* pool, const tf_batch* in, const uint8_t* keep, int threads, const tf_batch** out, const uint32_t** sel_out) {
    if (!pool || !in || !keep || !out || in->mem != TF_MEM_HOST) return TF_E_FATAL_ARG;
    try {
        const uint64_t n = in->nrows;
        if (threads <= 0) threads = (int)std::min<unsigned>(32, std::max(1u, std::thread::hardware_concurrency()));
        pool->next_buf = 0;
        const uint64_t CH = 32768, nch = (n + CH - 1) / CH;
        std::vector<uint64_t> kept(nch + 1, 0);
        parallel_chunks(pool, n, CH, threads, [&](uint64_t k) { uint64_t c = 0; for (uint64_t r = k * CH; r < std::min(n, (k + 1) * CH); r++) c += keep[r] != 0; kept[k + 1] = c; });
        for (uint64_t k = 0; k < nch; k++) kept[k + 1] += kept[k];
        const uint64_t m = kept[nch];
        uint32_t* sel = (uint32_t*)pool->take()->ensure(4 * m + 16, pool->want_pinned);
        parallel_chunks(pool, n, CH, threads, [&](uint64_t k) { uint64_t j = kept[k]; for (uint64_t r = k * CH; r < std::min(n, (k + 1) * CH); r++) if (keep[r]) sel[j++] = (uint32_t)r; });
        if (sel_out) *sel_out = sel;
        return gather_core(pool, in, sel, m, threads, out);
    } catch (const Fail& f) { pool->err = f.msg; return f.rc; }
    catch (const std::bad_alloc&) { pool->err = "host allocation failed"; return TF_E_RETRY_OOM; }
}

// Same with the selection vector already at hand (the device compacts the keep flags of phase one itself): sel = m ascending row indexes.
int tfgpu_batch_gather_sel(tfgpu_columnar* pool, const tf_batch* in, const uint32_t* sel, uint64_t m, int threads, const tf_batch** out) {
    if (!pool || !in || (!sel && m) || !out || in->mem != TF_MEM_HOST) return TF_E_FATAL_ARG;
    try {
        if (threads <= 0) threads = (int)std::min<unsigned>(32, std::max(1u, std::thread::hardware_concurrency()));
        pool->next_buf = 0;
        if (m && sel[m - 1] >= in->nrows) throw Fail{TF_E_FATAL_ARG, "selection names a row outside the batch"};
        return gather_core(pool, in, sel, m, threads, out);
    } catch (const Fail& f) { pool->err = f.msg; return f.rc; }
    catch (const std::bad_alloc&) { pool->err = "host allocation failed"; return TF_E_RETRY_OOM; }
}

int tfgpu_batch_to_rows(const tf_batch* b, uint8_t* out, uint64_t cap, uint64_t* row_off, uint64_t* need) {
    if (!b || !row_off || (b->mem != TF_MEM_HOST)) return TF_E_FATAL_ARG;
    const uint64_t n = b->nrows; const uint32_t nc = b->ncols;
    auto cell_len = [&](const tf_col& c, uint64_t r) -> uint32_t {
        if (c.flags & TF_COL_LENS8) return ((const uint8_t*)c.offsets)[r];
        if (c.flags & TF_COL_LENS16) return ((const uint16_t*)c.offsets)[r];
        return c.offsets[r + 1] - c.offsets[r];
    };
    std::vector<int> width(nc); std::vector<uint8_t> narrow(nc);
    for (uint32_t ci = 0; ci < nc; ci++) { width[ci] = fixed_width(b->cols[ci].type); narrow[ci] = !width[ci] && (b->cols[ci].flags & (TF_COL_LENS8 | TF_COL_LENS16)); }
    // Rows [r0, r1) written at `at` (or only measured when dst is null); cursor[ci]: where the rows' cells start in a heap whose column
    // carries lengths instead of offsets. Returns the end position.
    auto walk = [&](uint64_t r0, uint64_t r1, uint64_t at, uint64_t* cursor, uint8_t* dst, bool offsets_out) -> uint64_t {
        auto put = [&](const void* p, size_t k) { if (dst) std::memcpy(dst + at, p, k); at += k; };
        auto put8 = [&](uint8_t v) { if (dst) dst[at] = v; at++; };
        for (uint64_t r = r0; r < r1; r++) {
            if (offsets_out) row_off[r] = at;
            for (uint32_t ci = 0; ci < nc; ci++) {
                const tf_col& c = b->cols[ci];
                const bool valid = !c.validity || (c.validity[r >> 3] >> (r & 7) & 1);
                const int w = width[ci];
                uint32_t len = 0; const uint8_t* src = nullptr;
                if (!w) {
                    len = cell_len(c, r);
                    src = c.heap + (narrow[ci] ? cursor[ci] : c.offsets[r]);
                    cursor[ci] += len;
                }
                if (!valid) { put8(TF_V_NIL); continue; }
                if (w) {
                    if (is_time(c.type)) { put8(TF_V_TIME); put((const uint8_t*)c.values + 8 * r, 8); const uint32_t ns = c.aux ? ((const uint32_t*)c.aux)[r] : 0; put(&ns, 4); }
                    else { put8((uint8_t)canonical_tag(c.type)); put((const uint8_t*)c.values + (size_t)w * r, w); }
                } else {
                    const bool go_string = c.type == TF_UTF8 || (c.type == TF_ANY && c.aux && ((const uint8_t*)c.aux)[r] == 1);
                    put8(go_string ? TF_V_STRING : c.type == TF_BYTES ? TF_V_BYTES : TF_V_JSON); put(&len, 4); put(src, len);
                }
            }
        }
        return at;
    };
    int threads = (int)std::min<unsigned>(16, std::max(1u, std::thread::hardware_concurrency()));
    if (n < 8192 || threads < 2 || std::getenv("TFGPU_INVERSE_SEQUENTIAL")) {
        std::vector<uint64_t> cursor(nc, 0);
        const uint64_t total = walk(0, n, 0, cursor.data(), nullptr, false);                    // measure, then write when it fits
        const bool fits = out && total <= cap;
        std::fill(cursor.begin(), cursor.end(), 0);
        walk(0, n, 0, cursor.data(), fits ? out : nullptr, true);
        row_off[n] = total; if (need) *need = total;
        return fits ? TF_OK : TF_E_FATAL_ARG;
    }
    // in parallel: every task measures its rows, a prefix over the tasks gives each its place in the image (and in the heaps of the columns
    // that carry lengths), then every task writes its rows
    const uint64_t chunk = std::max<uint64_t>(1024, (n + (uint64_t)threads * 4 - 1) / ((uint64_t)threads * 4));
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    std::vector<uint64_t> bytes(nchunks + 1, 0), cur((nchunks + 1) * (size_t)nc, 0);
    auto run = [&](const std::function<void(uint64_t)>& task) {
        std::atomic<uint64_t> next{0}; std::vector<std::thread> ts;
        auto loop = [&] { for (;;) { const uint64_t k = next.fetch_add(1); if (k >= nchunks) return; task(k); } };
        for (int t = 1; t < threads; t++) ts.emplace_back(loop);
        loop();
        for (auto& t : ts) t.join();
    };
    run([&](uint64_t k) {
        std::vector<uint64_t> c0(nc, 0);
        bytes[k + 1] = walk(k * chunk, std::min(n, (k + 1) * chunk), 0, c0.data(), nullptr, false);
        for (uint32_t ci = 0; ci < nc; ci++) cur[(k + 1) * (size_t)nc + ci] = c0[ci];
    });
    for (uint64_t k = 0; k < nchunks; k++) { bytes[k + 1] += bytes[k]; for (uint32_t ci = 0; ci < nc; ci++) cur[(k + 1) * (size_t)nc + ci] += cur[k * (size_t)nc + ci]; }
    const uint64_t total = bytes[nchunks]; const bool fits = out && total <= cap;
    run([&](uint64_t k) {
        std::vector<uint64_t> c0(cur.begin() + k * (size_t)nc, cur.begin() + (k + 1) * (size_t)nc);
        walk(k * chunk, std::min(n, (k + 1) * chunk), bytes[k], c0.data(), fits ? out : nullptr, true);
    });
    row_off[n] = total; if (need) *need = total;
    return fits ? TF_OK : TF_E_FATAL_ARG;
}

}  // extern "C"
