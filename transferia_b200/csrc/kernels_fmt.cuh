// Value -> text on the device, bit-exact with the Go standard library forms the reference relies on:
//   fmt.Sprintf("%v", x)                 to_string.go:170 (mask_field / convert_to_string)
//   strconv.FormatFloat(f,'f',-1,bits)   httpuploader/marshal.go:150-152
//   encoding/json floats                 serializer/json.go:63-66
//   time.Time.Format(RFC3339Nano / DateOnly), time.Duration.String()
// Shortest round-trip float digits come from Ryu (Ulf Adams, 2018), written here for the GPU: one thread per value,
// 64x128-bit multiplies through __umul64hi, tables in ryu_tables.cuh (generated, validated by scripts/ryu_proto.py
// against CPython repr and numpy Dragon4; the float32 interval runs through the same 64-bit core).
// Every formatter writes through a Sink (`put(uint8_t)`): the SHA-256 sink for mask_field, a counting sink and a memory
// sink for text columns.
#pragma once
#include "device_types.cuh"
#include "kernels_encode.cuh"
#include "ryu_tables.cuh"

namespace tfk {

struct CountSink { uint32_t n; __device__ __forceinline__ void put(uint8_t) { n++; } };
struct MemSink { uint8_t* p; __device__ __forceinline__ void put(uint8_t b) { *p++ = b; } };
// MemSink that gathers 8 bytes in a register and stores aligned 64-bit words: row text is written by one thread per row, and a
// byte store per character makes the LSU the bottleneck. flush() must follow the last put().
struct WordSink {
    uint8_t* p; uint64_t acc; uint32_t nb;
    __device__ __forceinline__ explicit WordSink(uint8_t* dst) : p(dst), acc(0), nb(0) {}
    __device__ __forceinline__ void put(uint8_t b) {
        if (nb == 0 && ((uintptr_t)p & 7)) { *p++ = b; return; }          // head: up to the first 8-byte boundary
        acc |= (uint64_t)b << (8 * nb);
        if (++nb == 8) { *(uint64_t*)p = acc; p += 8; acc = 0; nb = 0; }
    }
    __device__ __forceinline__ void flush() { for (uint32_t i = 0; i < nb; i++) p[i] = (uint8_t)(acc >> (8 * i)); p += nb; nb = 0; acc = 0; }
};

// ------------------------------------------------------------------ Ryu core
// decimal digits without a local array (digits are packed into registers, least significant first, and read back from the top):
// a dynamically indexed local buffer here was placed outside the thread's stack by ptxas 12.9 in the mask translation unit
struct DigitRegs { uint64_t a, b, c; int n; };
__device__ __forceinline__ DigitRegs digits_of(uint64_t u) {
    DigitRegs d{0, 0, 0, 0};
    do {
        const uint64_t q = u / 10; const uint64_t dg = u - q * 10; u = q;
        if (d.n < 8) d.a |= dg << (8 * d.n); else if (d.n < 16) d.b |= dg << (8 * (d.n - 8)); else d.c |= dg << (8 * (d.n - 16));
        d.n++;
    } while (u);
    return d;
}
__device__ __forceinline__ uint8_t digit_at(const DigitRegs& d, int i) {
    const uint64_t w = i < 8 ? d.a >> (8 * i) : (i < 16 ? d.b >> (8 * (i - 8)) : d.c >> (8 * (i - 16)));
    return (uint8_t)('0' + (uint32_t)(w & 0xff));
}
struct DecF { uint64_t digits; int32_t exp10; };   // value = digits * 10^exp10

__device__ __forceinline__ uint32_t ryu_pow5bits(int32_t e) { return (uint32_t)(((uint32_t)e * 1217359u) >> 19) + 1; }
__device__ __forceinline__ uint32_t ryu_log10Pow2(int32_t e) { return ((uint32_t)e * 78913u) >> 18; }
__device__ __forceinline__ uint32_t ryu_log10Pow5(int32_t e) { return ((uint32_t)e * 732923u) >> 20; }
__device__ __forceinline__ uint64_t ryu_mulshift(uint64_t m, const uint64_t* mul, int32_t j) {   // (m * mul) >> j, j > 64
    const uint64_t b0hi = __umul64hi(m, mul[0]);
    const uint64_t b2lo = m * mul[1], b2hi = __umul64hi(m, mul[1]);
    const uint64_t sum = b0hi + b2lo; const uint64_t hi = b2hi + (sum < b0hi ? 1 : 0);
    const int32_t s = j - 64;
    return s == 0 ? sum : ((sum >> s) | (hi << (64 - s)));
}
__device__ __forceinline__ uint32_t ryu_pow5factor(uint64_t v) { uint32_t c = 0; while (v && v % 5 == 0) { v /= 5; c++; } return c; }

static __device__ DecF ryu_d2d(uint64_t mant, uint32_t expo, int mbits, int bias) {
    int32_t e2; uint64_t m2;
    if (expo == 0) { e2 = 1 - bias - mbits - 2; m2 = mant; } else { e2 = (int32_t)expo - bias - mbits - 2; m2 = (1ull << mbits) | mant; }
    const bool accept = (m2 & 1) == 0;
    const uint64_t mv = 4 * m2;
    const uint32_t mmShift = (mant != 0 || expo <= 1) ? 1 : 0;
    uint64_t vr, vp, vm; int32_t e10; bool vmTZ = false, vrTZ = false;
    if (e2 >= 0) {
        const uint32_t q = ryu_log10Pow2(e2) - (e2 > 3 ? 1 : 0);
        e10 = (int32_t)q;
        const int32_t k = 125 + (int32_t)ryu_pow5bits((int32_t)q) - 1;
        const int32_t i = -e2 + (int32_t)q + k;
        const uint64_t* mul = RYU_POW5_INV_SPLIT[q];
        vr = ryu_mulshift(mv, mul, i); vp = ryu_mulshift(mv + 2, mul, i); vm = ryu_mulshift(mv - 1 - mmShift, mul, i);
        if (q <= 21) {
            if (mv % 5 == 0) vrTZ = ryu_pow5factor(mv) >= q;
            else if (accept) vmTZ = ryu_pow5factor(mv - 1 - mmShift) >= q;
            else vp -= (ryu_pow5factor(mv + 2) >= q) ? 1 : 0;
        }
    } else {
        const uint32_t q = ryu_log10Pow5(-e2) - (-e2 > 1 ? 1 : 0);
        e10 = (int32_t)q + e2;
        const int32_t i = -e2 - (int32_t)q;
        const int32_t k = (int32_t)ryu_pow5bits(i) - 125;
        const int32_t j = (int32_t)q - k;
        const uint64_t* mul = RYU_POW5_SPLIT[i];
        vr = ryu_mulshift(mv, mul, j); vp = ryu_mulshift(mv + 2, mul, j); vm = ryu_mulshift(mv - 1 - mmShift, mul, j);
        if (q <= 1) { vrTZ = true; if (accept) vmTZ = mmShift == 1; else --vp; }
        else if (q < 63) vrTZ = (mv & ((1ull << q) - 1)) == 0;
    }
    int32_t removed = 0; uint32_t last = 0; uint64_t out;
    if (vmTZ || vrTZ) {
        while (vp / 10 > vm / 10) { vmTZ &= vm % 10 == 0; vrTZ &= last == 0; last = (uint32_t)(vr % 10); vr /= 10; vp /= 10; vm /= 10; removed++; }
        if (vmTZ) while (vm % 10 == 0) { vrTZ &= last == 0; last = (uint32_t)(vr % 10); vr /= 10; vp /= 10; vm /= 10; removed++; }
        if (vrTZ && last == 5 && vr % 2 == 0) last = 4;
        out = vr + (((vr == vm && (!accept || !vmTZ)) || last >= 5) ? 1 : 0);
    } else {
        bool roundUp = false;
        while (vp / 10 > vm / 10) { roundUp = vr % 10 >= 5; vr /= 10; vp /= 10; vm /= 10; removed++; }
        out = vr + ((vr == vm || roundUp) ? 1 : 0);
    }
    DecF r; r.digits = out; r.exp10 = e10 + removed;
    while (r.digits && r.digits % 10 == 0) { r.digits /= 10; r.exp10++; }     // shortest form: no trailing zeros
    return r;
}

// strconv layouts over the shortest digits (see oracle/go_strconv.hpp layout_e / layout_f)
enum FloatMode { FM_V = 0 /* fmt %v: 'g', eprec 6 */, FM_F = 1 /* 'f', -1 */, FM_JSON = 2 /* encoding/json */ };

template <typename Sink> __device__ void fmt_float_bits(Sink& s, uint64_t bits, bool is32, int mode) {
    const int mbits = is32 ? 23 : 52, ebits = is32 ? 8 : 11, bias = is32 ? 127 : 1023;
    const bool neg = (bits >> (mbits + ebits)) & 1;
    const uint32_t expo = (uint32_t)((bits >> mbits) & ((1u << ebits) - 1));
    const uint64_t mant = bits & ((1ull << mbits) - 1);
    if (expo == (1u << ebits) - 1) {
        const char* t = mant ? "NaN" : (neg ? "-Inf" : "+Inf");
        while (*t) s.put((uint8_t)*t++);
        return;
    }
    DigitRegs dr{0, 0, 0, 0}; int nd = 0, dp = 0;      // digit i (most significant first) = digit_at(dr, nd - 1 - i)
    if (expo || mant) {
        DecF r = ryu_d2d(mant, expo, mbits, bias);
        dr = digits_of(r.digits);
        nd = dr.n; dp = nd + r.exp10;
    }
#define TF_D(i) digit_at(dr, nd - 1 - (i))
    bool use_e = false;
    if (mode == FM_V) { const int ex = dp - 1; use_e = (ex < -4 || ex >= 6); }      // zero: nd = dp = 0 -> ex = -1 -> 'f' -> "0"
    else if (mode == FM_JSON && nd) {        // |f| < 1e-6 || |f| >= 1e21  <=> decimal exponent < -6 or >= 21
        const int ex = dp - 1; use_e = (ex < -6 || ex >= 21);
    }
    if (neg) s.put('-');
    if (use_e) {       // d.ddde±XX  (at least two exponent digits; encoding/json strips a leading zero: e-09 -> e-9)
        s.put(nd ? TF_D(0) : (uint8_t)'0');
        if (nd > 1) { s.put('.'); for (int i = 1; i < nd; i++) s.put(TF_D(i)); }
        s.put('e');
        int ex = nd ? dp - 1 : 0;
        if (ex < 0) { s.put('-'); ex = -ex; } else s.put('+');
        if (ex < 10) { if (mode != FM_JSON) s.put('0'); s.put((uint8_t)('0' + ex)); }
        else if (ex < 100) { s.put((uint8_t)('0' + ex / 10)); s.put((uint8_t)('0' + ex % 10)); }
        else { s.put((uint8_t)('0' + ex / 100)); s.put((uint8_t)('0' + (ex / 10) % 10)); s.put((uint8_t)('0' + ex % 10)); }
    } else {           // %f with the shortest precision
        if (dp > 0) { int m = nd < dp ? nd : dp; for (int i = 0; i < m; i++) s.put(TF_D(i)); for (; m < dp; m++) s.put('0'); }
        else s.put('0');
        const int prec = nd - dp > 0 ? nd - dp : 0;
        if (prec > 0) { s.put('.'); for (int i = 0; i < prec; i++) { const int j = dp + i; s.put((j >= 0 && j < nd) ? TF_D(j) : (uint8_t)'0'); } }
    }
#undef TF_D
}

// ------------------------------------------------------------------ integers, time, duration
template <typename Sink> __device__ __forceinline__ void fmt_u64(Sink& s, uint64_t u) {
    const DigitRegs d = digits_of(u);
    for (int i = d.n; i-- > 0;) s.put(digit_at(d, i));
}
template <typename Sink> __device__ __forceinline__ void fmt_i64(Sink& s, int64_t v) { if (v < 0) { s.put('-'); fmt_u64(s, (uint64_t)0 - (uint64_t)v); } else fmt_u64(s, (uint64_t)v); }
template <typename Sink> __device__ __forceinline__ void fmt_pad(Sink& s, int64_t v, int wdt) {   // Go appendInt(b, v, width)
    if (v < 0) { s.put('-'); v = -v; }
    const DigitRegs d = digits_of((uint64_t)v);
    for (int i = d.n; i < wdt; i++) s.put('0');
    for (int i = d.n; i-- > 0;) s.put(digit_at(d, i));
}

template <typename Sink> __device__ __forceinline__ void fmt_lit(Sink& s, const char* t) { while (*t) s.put((uint8_t)*t++); }

__device__ inline void civil_from_days_d(int64_t z, int64_t& y, unsigned& m, unsigned& d) {
    z += 719468;
    const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
    const unsigned doe = (unsigned)(z - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    y = (int64_t)yoe + era * 400;
    const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doy + 2) / 153;
    d = doy - (153 * mp + 2) / 5 + 1;
    m = mp < 10 ? mp + 3 : mp - 9;
    y += (m <= 2);
}
// time.Time.UTC().Format("2006-01-02") / RFC3339Nano
template <typename Sink> __device__ void fmt_time(Sink& s, int64_t sec, uint32_t nsec, bool date_only) {
    int64_t days = sec / 86400; int64_t sod = sec - days * 86400; if (sod < 0) { sod += 86400; days--; }
    int64_t y; unsigned m, d; civil_from_days_d(days, y, m, d);
    fmt_pad(s, y, 4); s.put('-'); fmt_pad(s, m, 2); s.put('-'); fmt_pad(s, d, 2);
    if (date_only) return;
    s.put('T'); fmt_pad(s, sod / 3600, 2); s.put(':'); fmt_pad(s, (sod / 60) % 60, 2); s.put(':'); fmt_pad(s, sod % 60, 2);
    if (nsec) {
        uint32_t v = nsec; int n = 9; while (v % 10 == 0) { v /= 10; n--; }        // n significant fraction digits, v = those digits
        s.put('.'); fmt_pad(s, v, n);
    }
    s.put('Z');
}
// time.Duration.String(): "72h3m0.5s", "1.5µs", "0s" (Go time/time.go Duration.String, fmtFrac, fmtInt)
template <typename Sink> __device__ void fmt_duration(Sink& s, int64_t d) {
    uint64_t u = (uint64_t)d; const bool neg = d < 0; if (neg) u = (uint64_t)0 - u;
    if (u == 0) { s.put('0'); s.put('s'); return; }
    if (neg) s.put('-');
    // fmtFrac: v / 10^prec as "int[.frac]" with the fraction's trailing zeros dropped
    auto int_frac = [&](uint64_t ip, uint64_t fr, int prec) {
        fmt_u64(s, ip);
        if (fr) { while (fr % 10 == 0) { fr /= 10; prec--; } s.put('.'); fmt_pad(s, (int64_t)fr, prec); }
    };
    if (u < 1000000000ull) {
        if (u < 1000ull) { fmt_u64(s, u); s.put('n'); }
        else if (u < 1000000ull) { int_frac(u / 1000, u % 1000, 3); s.put(0xC2); s.put(0xB5); }     // "µ"
        else { int_frac(u / 1000000, u % 1000000, 6); s.put('m'); }
        s.put('s');
    } else {
        const uint64_t fr = u % 1000000000ull; uint64_t v = u / 1000000000ull;
        const uint64_t sec = v % 60; v /= 60;
        if (v > 0) { const uint64_t mn = v % 60; v /= 60; if (v > 0) { fmt_u64(s, v); s.put('h'); } fmt_u64(s, mn); s.put('m'); }
        int_frac(sec, fr, 9); s.put('s');
    }
}

// encoding/json string encoder, escapeHTML = true (json.Marshal of a Go string inside an `any` column)
template <typename Sink> __device__ void fmt_json_string(Sink& s, const uint8_t* p, uint32_t n, bool html = true) {
    const char* hex = "0123456789abcdef";
    s.put('"');
    uint32_t i = 0;
    while (i < n) {
        const uint8_t b = p[i];
        if (b < 0x80) {
            if (b >= 0x20 && b != '"' && b != '\\' && (!html || (b != '<' && b != '>' && b != '&'))) { s.put(b); i++; continue; }
            s.put('\\');
            switch (b) {
            case '\\': case '"': s.put(b); break;
            case '\b': s.put('b'); break; case '\f': s.put('f'); break; case '\n': s.put('n'); break; case '\r': s.put('r'); break; case '\t': s.put('t'); break;
            default: s.put('u'); s.put('0'); s.put('0'); s.put((uint8_t)hex[b >> 4]); s.put((uint8_t)hex[b & 15]);
            }
            i++; continue;
        }
        uint32_t r = 0xFFFD, w = 1;
        if (b >= 0xC2 && b <= 0xDF && i + 1 < n && (p[i + 1] & 0xC0) == 0x80) { r = ((b & 0x1Fu) << 6) | (p[i + 1] & 0x3Fu); w = 2; }
        else if (b >= 0xE0 && b <= 0xEF && i + 2 < n && (p[i + 1] & 0xC0) == 0x80 && (p[i + 2] & 0xC0) == 0x80) {
            const uint32_t t = ((b & 0x0Fu) << 12) | ((p[i + 1] & 0x3Fu) << 6) | (p[i + 2] & 0x3Fu);
            if (t >= 0x800 && !(t >= 0xD800 && t <= 0xDFFF)) { r = t; w = 3; }
        } else if (b >= 0xF0 && b <= 0xF4 && i + 3 < n && (p[i + 1] & 0xC0) == 0x80 && (p[i + 2] & 0xC0) == 0x80 && (p[i + 3] & 0xC0) == 0x80) {
            const uint32_t t = ((b & 0x07u) << 18) | ((p[i + 1] & 0x3Fu) << 12) | ((p[i + 2] & 0x3Fu) << 6) | (p[i + 3] & 0x3Fu);
            if (t >= 0x10000 && t <= 0x10FFFF) { r = t; w = 4; }
        }
        if (r == 0xFFFD && w == 1) { fmt_lit(s, "\\ufffd"); i++; continue; }
        if (r == 0x2028 || r == 0x2029) { fmt_lit(s, "\\u202"); s.put((uint8_t)hex[r & 0xF]); i += w; continue; }
        for (uint32_t k = 0; k < w; k++) s.put(p[i + k]);
        i += w;
    }
    s.put('"');
}

// to_string.SerializeToString (pkg/transformer/registry/to_string/to_string.go:145-171) for a typed column value
template <typename Sink> __device__ void fmt_value(Sink& s, const DCol& c, uint64_t r) {
    if (!row_valid(c, r)) { if (c.type == TF_ANY) fmt_lit(s, "null"); else fmt_lit(s, "<nil>"); return; }   // json.Marshal(nil) / %v of nil
    switch (c.type) {
    case TF_INT8: fmt_i64(s, ((const int8_t*)c.values)[r]); break;
    case TF_INT16: fmt_i64(s, ((const int16_t*)c.values)[r]); break;
    case TF_INT32: fmt_i64(s, ((const int32_t*)c.values)[r]); break;
    case TF_INT64: fmt_i64(s, ((const int64_t*)c.values)[r]); break;
    case TF_UINT8: fmt_u64(s, c.values[r]); break;
    case TF_UINT16: fmt_u64(s, ((const uint16_t*)c.values)[r]); break;
    case TF_UINT32: fmt_u64(s, ((const uint32_t*)c.values)[r]); break;
    case TF_UINT64: fmt_u64(s, ((const uint64_t*)c.values)[r]); break;
    case TF_FLOAT: fmt_float_bits(s, ((const uint32_t*)c.values)[r], true, FM_V); break;
    case TF_DOUBLE: fmt_float_bits(s, ((const uint64_t*)c.values)[r], false, FM_V); break;
    case TF_BOOLEAN: fmt_lit(s, c.values[r] ? "true" : "false"); break;
    case TF_INTERVAL: fmt_duration(s, ((const int64_t*)c.values)[r]); break;
    case TF_DATE: fmt_time(s, ((const int64_t*)c.values)[r], 0, true); break;
    case TF_DATETIME: case TF_TIMESTAMP: fmt_time(s, ((const int64_t*)c.values)[r], c.aux ? ((const uint32_t*)c.aux)[r] : 0, false); break;
    case TF_BYTES: case TF_UTF8: { const uint8_t* p = c.heap + c.offsets[r]; const uint32_t L = c.offsets[r + 1] - c.offsets[r]; for (uint32_t k = 0; k < L; k++) s.put(p[k]); break; }
    case TF_ANY: {
        const uint8_t* p = c.heap + c.offsets[r]; const uint32_t L = c.offsets[r + 1] - c.offsets[r];
        if (c.aux && c.aux[r] == 1) fmt_json_string(s, p, L); else for (uint32_t k = 0; k < L; k++) s.put(p[k]);
        break;
    }
    }
}

}  // namespace tfk
