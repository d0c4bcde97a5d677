// ClickHouse JSONEachRow rows on the device.
//   reference: MarshalCItoJSON pkg/providers/clickhouse/httpuploader/marshal.go:88-253 (+ marshalTime :63-78,
//   questionableQuoter :264-266): `{"col":value,...}\n`, schema order, nil columns omitted, strings escape only
//   `\` and `"` (bytes otherwise untouched, non-UTF-8 included), ints/floats bare ('f', -1) unless the target column is
//   String, time by the target ClickHouse type, non-string `any` values as a JSON-quoted string of their JSON text.
// Row text has no fixed width: pass 1 sizes every kept row (CountSink), a tile scan places them, pass 2 writes them.
#pragma once
#include "device_types.cuh"
#include "kernels_encode.cuh"
#include "kernels_fmt.cuh"
#include "kernels_mask.cuh"

namespace tfk {

enum JsonChClass : int32_t { JC_OTHER = 0, JC_STRING = 1, JC_DATE = 2, JC_DATETIME = 3, JC_DT64 = 4 };
struct JsonCol { int32_t col; int32_t name_off, name_len; int32_t ch_class; int32_t prec; int32_t result_tf; int32_t pad0, pad1; };

template <typename Inner> struct EscSink {     // questionableQuoter: `\` -> `\\`, `"` -> `\"`
    Inner* in;
    __device__ __forceinline__ void put(uint8_t b) { if (b == '\\' || b == '"') in->put('\\'); in->put(b); }
};

template <typename Sink> __device__ __forceinline__ void json_time(Sink& s, int64_t sec, uint32_t nsec, const JsonCol& jc) {
    switch (jc.ch_class) {
    case JC_STRING: {      // v.Format("2006-01-02 15:04:05.999999999 -0700 MST") in UTC
        s.put('"');
        int64_t days = sec / 86400; int64_t sod = sec - days * 86400; if (sod < 0) { sod += 86400; days--; }
        int64_t y; unsigned m, d; civil_from_days_d(days, y, m, d);
        fmt_pad(s, y, 4); s.put('-'); fmt_pad(s, m, 2); s.put('-'); fmt_pad(s, d, 2); s.put(' ');
        fmt_pad(s, sod / 3600, 2); s.put(':'); fmt_pad(s, (sod / 60) % 60, 2); s.put(':'); fmt_pad(s, sod % 60, 2);
        if (nsec) { char b[9]; uint32_t v = nsec; for (int i = 8; i >= 0; i--) { b[i] = (char)('0' + v % 10); v /= 10; } int n = 9; while (n > 0 && b[n - 1] == '0') n--; s.put('.'); for (int i = 0; i < n; i++) s.put((uint8_t)b[i]); }
        fmt_lit(s, " +0000 UTC\""); break;
    }
    case JC_DT64: {        // UnixNano() / 10^(9-p) for 0 < p < 9 (Go integer division: toward zero)
        int64_t full = sec * 1000000000LL + (int64_t)nsec;
        if (jc.prec > 0 && jc.prec < 9) { int64_t div = 1; for (int i = 0; i < 9 - jc.prec; i++) div *= 10; full = full / div; }
        fmt_i64(s, full); break;
    }
    case JC_DATE: s.put('"'); fmt_time(s, sec, 0, true); s.put('"'); break;
    default: fmt_i64(s, sec);
    }
}

// one column's value; returns false when the column is omitted (nil, or a JSON `null`)
template <typename Sink> __device__ bool json_value(Sink& s, const DCol& c, uint64_t r, const JsonCol& jc, const MaskKey* keys, bool sizing) {
    const bool str = jc.ch_class == JC_STRING;
    if (c.out_kind == OK_MASK) {                 // hex digest, a Go string: never nil
        s.put('"');
        if (sizing) { for (int i = 0; i < 64; i++) s.put('0'); }
        else { uint8_t hx[64]; mask_digest_hex(c, r, keys[c.mask_slot], hx); for (int i = 0; i < 64; i++) s.put(hx[i]); }
        s.put('"'); return true;
    }
    if (c.out_kind == OK_TOSTR) { s.put('"'); EscSink<Sink> es{&s}; fmt_value(es, c, r); s.put('"'); return true; }
    if (c.out_kind == OK_TODT) {
        int64_t sec = 0; if (row_valid(c, r)) sec = c.type == TF_INT32 ? (int64_t)((const int32_t*)c.values)[r] : (int64_t)((const uint32_t*)c.values)[r];
        json_time(s, sec, 0, jc); return true;
    }
    if (!row_valid(c, r)) return false;
    switch (c.type) {
    case TF_INT8: if (str) s.put('"'); fmt_i64(s, ((const int8_t*)c.values)[r]); if (str) s.put('"'); break;
    case TF_INT16: if (str) s.put('"'); fmt_i64(s, ((const int16_t*)c.values)[r]); if (str) s.put('"'); break;
    case TF_INT32: if (str) s.put('"'); fmt_i64(s, ((const int32_t*)c.values)[r]); if (str) s.put('"'); break;
    case TF_INT64: if (str) s.put('"'); fmt_i64(s, ((const int64_t*)c.values)[r]); if (str) s.put('"'); break;
    case TF_UINT8: if (str) s.put('"'); fmt_u64(s, c.values[r]); if (str) s.put('"'); break;
    case TF_UINT16: if (str) s.put('"'); fmt_u64(s, ((const uint16_t*)c.values)[r]); if (str) s.put('"'); break;
    case TF_UINT32: if (str) s.put('"'); fmt_u64(s, ((const uint32_t*)c.values)[r]); if (str) s.put('"'); break;
    case TF_UINT64: if (str) s.put('"'); fmt_u64(s, ((const uint64_t*)c.values)[r]); if (str) s.put('"'); break;
    case TF_FLOAT: if (str) s.put('"'); fmt_float_bits(s, ((const uint32_t*)c.values)[r], true, FM_F); if (str) s.put('"'); break;
    case TF_DOUBLE: if (str) s.put('"'); fmt_float_bits(s, ((const uint64_t*)c.values)[r], false, FM_F); if (str) s.put('"'); break;
    case TF_BOOLEAN: fmt_lit(s, c.values[r] ? "true" : "false"); break;           // DataType == boolean (marshal.go:187-192)
    case TF_INTERVAL: {    // json.Marshal(time.Duration) is an integer; a non-`any` column re-marshals that text as a string
        s.put('"'); fmt_i64(s, ((const int64_t*)c.values)[r]); s.put('"'); break;
    }
    case TF_DATE: case TF_DATETIME: case TF_TIMESTAMP:
        json_time(s, ((const int64_t*)c.values)[r], c.aux ? ((const uint32_t*)c.aux)[r] : 0, jc); break;
    case TF_BYTES: case TF_UTF8: {
        const uint8_t* p = c.heap + c.offsets[r]; const uint32_t L = c.offsets[r + 1] - c.offsets[r];
        s.put('"'); for (uint32_t k = 0; k < L; k++) { const uint8_t b = p[k]; if (b == '\\' || b == '"') s.put('\\'); s.put(b); } s.put('"'); break;
    }
    case TF_ANY: {
        const uint8_t* p = c.heap + c.offsets[r]; const uint32_t L = c.offsets[r + 1] - c.offsets[r];
        if (c.aux && c.aux[r] == 1) { s.put('"'); for (uint32_t k = 0; k < L; k++) { const uint8_t b = p[k]; if (b == '\\' || b == '"') s.put('\\'); s.put(b); } s.put('"'); break; }
        if (L == 4 && p[0] == 'n' && p[1] == 'u' && p[2] == 'l' && p[3] == 'l') return false;      // :229-233
        fmt_json_string(s, p, L);      // any -> String column: json.Marshal(string(r)) (:234-239)
        break;
    }
    }
    return true;
}

template <typename Sink> __device__ void json_row(Sink& s, const DCol* cols, const JsonCol* jcols, int njc, const uint8_t* names, const MaskKey* keys, uint64_t r, bool sizing) {
    s.put('{');
    bool first = true;
    for (int k = 0; k < njc; k++) {
        const JsonCol jc = jcols[k]; const DCol& c = cols[jc.col];
        // is the column present? (decide before writing its name)
        bool present;
        if (c.out_kind == OK_MASK || c.out_kind == OK_TOSTR || c.out_kind == OK_TODT) present = true;
        else if (!row_valid(c, r)) present = false;
        else if (c.type == TF_ANY && !(c.aux && c.aux[r] == 1)) { const uint8_t* p = c.heap + c.offsets[r]; present = !((c.offsets[r + 1] - c.offsets[r]) == 4 && p[0] == 'n' && p[1] == 'u' && p[2] == 'l' && p[3] == 'l'); }
        else present = true;
        if (!present) continue;
        if (!first) s.put(',');
        first = false;
        s.put('"'); for (int i = 0; i < jc.name_len; i++) s.put(names[jc.name_off + i]); s.put('"'); s.put(':');
        json_value(s, c, r, jc, keys, sizing);
    }
    s.put('}'); s.put('\n');
}

struct JsonArgs {
    const DCol* cols; const JsonCol* jcols; int njc; const uint8_t* names; const MaskKey* keys;
    const uint32_t* sel; DState* st; uint8_t* raw; uint32_t* row_size; uint32_t* tile_sum; const uint64_t* tile_base; const uint64_t* col_bytes;
};

#define TF_JSON_TILE 256

__global__ void __launch_bounds__(TF_JSON_TILE) k_json_sizes(JsonArgs a) {
    __shared__ uint32_t sm[33];
    const uint64_t n = a.st->n_kept;
    const uint64_t j = (uint64_t)blockIdx.x * TF_JSON_TILE + threadIdx.x;
    if ((uint64_t)blockIdx.x * TF_JSON_TILE >= n) return;
    uint32_t sz = 0;
    if (j < n) { const uint64_t r = a.sel ? a.sel[j] : j; CountSink cs; cs.n = 0; json_row(cs, a.cols, a.jcols, a.njc, a.names, a.keys, r, true); sz = cs.n; a.row_size[j] = sz; }
    uint32_t tot; block_excl_scan(sz, &tot, sm);
    if (threadIdx.x == 0) a.tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(TF_JSON_TILE) k_json_write(JsonArgs a) {
    __shared__ uint32_t sm[33];
    const uint64_t n = a.st->n_kept;
    if (blockIdx.x == 0 && threadIdx.x == 0) { a.st->raw_total = n ? a.col_bytes[0] : 0; a.st->n_frames = 0; }
    const uint64_t j = (uint64_t)blockIdx.x * TF_JSON_TILE + threadIdx.x;
    if ((uint64_t)blockIdx.x * TF_JSON_TILE >= n) return;
    const uint32_t sz = j < n ? a.row_size[j] : 0;
    uint32_t tot; const uint32_t ex = block_excl_scan(sz, &tot, sm);
    if (j >= n) return;
    const uint64_t r = a.sel ? a.sel[j] : j;
    MemSink ms; ms.p = a.raw + a.tile_base[blockIdx.x] + ex;
    json_row(ms, a.cols, a.jcols, a.njc, a.names, a.keys, r, false);
}

}  // namespace tfk
