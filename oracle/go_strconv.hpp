// ORACLE — test infrastructure only. Never linked into or called from the product path.
//
// CPU restatement of the Go standard-library text formatting the reference's hot path
// relies on.  The reference has no first-party code here; it calls the Go stdlib:
//   fmt.Sprintf("%v", v)                 pkg/transformer/registry/to_string/to_string.go:170
//   strconv.FormatFloat(f,'f',-1,bits)   pkg/providers/clickhouse/httpuploader/marshal.go:150-152
//   encoding/json float encoder          pkg/serializer/json.go:63-66
//   time.Time.Format(RFC3339Nano|DateOnly) to_string.go:164,168
//   time.Duration.String()               (mask golden: time.Minute -> "1m0s", hmac_hasher_test.go:92)
// Algorithms follow the published Go 1.25 stdlib behaviour (strconv/ftoa.go shortest
// round-trip digits: Steele-White/dragon4 free-format with round-half-even boundaries;
// %e/%f/%g layout rules; time layout rules).  Pinned by the mask golden
// (pkg/transformer/registry/mask/gotest/canondata/result.json) and cross-checked in
// tests against CPython repr() / numpy Dragon4, which implement the same
// shortest-closest definition.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <string>
#include <algorithm>

namespace orc {

// ---------------------------------------------------------------- integers
inline std::string fmt_i64(int64_t v) {
    char buf[24]; int n = 0;
    uint64_t u = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    do { buf[n++] = char('0' + u % 10); u /= 10; } while (u);
    if (v < 0) buf[n++] = '-';
    std::string s(buf, n); std::reverse(s.begin(), s.end()); return s;
}
inline std::string fmt_u64(uint64_t u) {
    char buf[24]; int n = 0;
    do { buf[n++] = char('0' + u % 10); u /= 10; } while (u);
    std::string s(buf, n); std::reverse(s.begin(), s.end()); return s;
}

// ---------------------------------------------------------------- tiny bignum (exact dragon4)
struct Big {
    static constexpr int N = 84;            // 84*32 = 2688 bits > 1074 + 1100
    uint32_t w[N]; int n;                   // little-endian limbs, n = used limbs
    Big() : n(0) { std::memset(w, 0, sizeof w); }
    explicit Big(uint64_t v) : n(0) { std::memset(w, 0, sizeof w); w[0] = (uint32_t)v; w[1] = (uint32_t)(v >> 32); n = w[1] ? 2 : (w[0] ? 1 : 0); }
    void trim() { while (n > 0 && w[n - 1] == 0) --n; }
    void mul_small(uint32_t m) {
        uint64_t c = 0;
        for (int i = 0; i < n; i++) { uint64_t t = (uint64_t)w[i] * m + c; w[i] = (uint32_t)t; c = t >> 32; }
        if (c) w[n++] = (uint32_t)c;
    }
    void shl(int bits) {
        int ls = bits / 32, bs = bits % 32;
        if (n == 0) return;
        if (bs) {
            uint32_t c = 0;
            for (int i = 0; i < n; i++) { uint32_t t = w[i]; w[i] = (t << bs) | c; c = t >> (32 - bs); }
            if (c) w[n++] = c;
        }
        if (ls) {
            for (int i = n - 1; i >= 0; i--) w[i + ls] = w[i];
            for (int i = 0; i < ls; i++) w[i] = 0;
            n += ls;
        }
    }
    static int cmp(const Big& a, const Big& b) {
        if (a.n != b.n) return a.n < b.n ? -1 : 1;
        for (int i = a.n - 1; i >= 0; i--) if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
        return 0;
    }
    void add(const Big& b) {
        int m = std::max(n, b.n); uint64_t c = 0;
        for (int i = 0; i < m; i++) { uint64_t t = (uint64_t)w[i] + b.w[i] + c; w[i] = (uint32_t)t; c = t >> 32; }
        n = m; if (c) w[n++] = (uint32_t)c;
    }
    void sub(const Big& b) {   // requires *this >= b
        int64_t c = 0;
        for (int i = 0; i < n; i++) { int64_t t = (int64_t)w[i] - b.w[i] + c; if (t < 0) { t += (1LL << 32); c = -1; } else c = 0; w[i] = (uint32_t)t; }
        trim();
    }
};

// Shortest decimal digits d[0..nd) and decimal point position dp such that
// value = 0.d0d1... * 10^dp, exactly Go's strconv "shortest" (%v / prec = -1).
struct Digits { char d[24]; int nd; int dp; };

// mant: integer significand, e2: binary exponent (value = mant * 2^e2),
// lower_closer: the gap below is half the gap above (mantissa is a power of two).
inline Digits shortest_digits(uint64_t mant, int e2, bool lower_closer) {
    Digits out; out.nd = 0; out.dp = 0;
    if (mant == 0) return out;
    const bool even = (mant & 1) == 0;      // round-half-even: boundaries are inclusive
    // value = r/s, m+ and m- are the half-gaps, all scaled by 2 (and 2 more if lower_closer)
    Big r(mant), s(1), mp(1), mm(1);
    int sh = lower_closer ? 2 : 1;
    if (e2 >= 0) { r.shl(e2 + sh); s.shl(sh); mp.shl(e2 + (lower_closer ? 1 : 0)); mm.shl(e2); }
    else         { r.shl(sh); s.shl(-e2 + sh); mp.shl(lower_closer ? 1 : 0); /* mm = 1 */ }
    // estimate k = ceil(log10(value))
    double lg = (std::log10((double)mant) + e2 * 0.30102999566398119521);
    int k = (int)std::ceil(lg - 1e-10);
    if (k >= 0) { for (int i = 0; i < k; i++) s.mul_small(10); }
    else        { for (int i = 0; i < -k; i++) { r.mul_small(10); mp.mul_small(10); mm.mul_small(10); } }
    // fix-up so that (r + m+)/s < 1 (or <= when boundaries exclusive) and *10 >= 1
    auto high_ge_s = [&]() { Big t = r; t.add(mp); int c = Big::cmp(t, s); return even ? c >= 0 : c > 0; };
    while (high_ge_s()) { s.mul_small(10); k++; }
    for (;;) {
        Big r10 = r; r10.mul_small(10); Big mp10 = mp; mp10.mul_small(10);
        Big t = r10; t.add(mp10); int c = Big::cmp(t, s);
        bool ge = even ? c >= 0 : c > 0;
        if (ge) break;
        r = r10; mp = mp10; mm.mul_small(10); k--;
    }
    out.dp = k;
    for (;;) {
        r.mul_small(10); mp.mul_small(10); mm.mul_small(10);
        int d = 0;
        while (Big::cmp(r, s) >= 0) { r.sub(s); d++; }
        int c1 = Big::cmp(r, mm); bool tc1 = even ? c1 <= 0 : c1 < 0;
        Big t = r; t.add(mp); int c2 = Big::cmp(t, s); bool tc2 = even ? c2 >= 0 : c2 > 0;
        if (!tc1 && !tc2) { out.d[out.nd++] = char('0' + d); continue; }
        if (tc1 && !tc2) { out.d[out.nd++] = char('0' + d); break; }
        if (!tc1 && tc2) { d++; }
        else {
            Big r2 = r; r2.shl(1); int c = Big::cmp(r2, s);
            if (c > 0 || (c == 0 && (d & 1))) d++;
        }
        // d may be 10 only if rounding carried; propagate
        if (d == 10) {
            int i = out.nd - 1;
            while (i >= 0 && out.d[i] == '9') { i--; }
            if (i < 0) { out.d[0] = '1'; out.nd = 1; out.dp++; }
            else { out.d[i]++; out.nd = i + 1; }
        } else out.d[out.nd++] = char('0' + d);
        break;
    }
    while (out.nd > 1 && out.d[out.nd - 1] == '0') out.nd--;   // cannot happen for shortest, defensive
    return out;
}

inline Digits shortest_f64(double v) {   // v finite, > 0
    uint64_t bits; std::memcpy(&bits, &v, 8);
    uint64_t frac = bits & ((1ULL << 52) - 1); int ex = (int)((bits >> 52) & 0x7FF);
    uint64_t mant; int e2;
    if (ex == 0) { mant = frac; e2 = -1074; } else { mant = frac | (1ULL << 52); e2 = ex - 1075; }
    bool lc = (frac == 0 && ex > 1);
    return shortest_digits(mant, e2, lc);
}
inline Digits shortest_f32(float v) {
    uint32_t bits; std::memcpy(&bits, &v, 4);
    uint32_t frac = bits & ((1u << 23) - 1); int ex = (int)((bits >> 23) & 0xFF);
    uint64_t mant; int e2;
    if (ex == 0) { mant = frac; e2 = -149; } else { mant = frac | (1u << 23); e2 = ex - 150; }
    bool lc = (frac == 0 && ex > 1);
    return shortest_digits(mant, e2, lc);
}

// strconv fmtE with prec = nd-1 (shortest): d.ddde±XX
inline std::string layout_e(bool neg, const Digits& g) {
    std::string s; if (neg) s += '-';
    s += g.nd ? g.d[0] : '0';
    if (g.nd > 1) { s += '.'; s.append(g.d + 1, g.nd - 1); }
    s += 'e';
    int exp = g.nd ? g.dp - 1 : 0;
    if (exp < 0) { s += '-'; exp = -exp; } else s += '+';
    if (exp < 10) { s += '0'; s += char('0' + exp); }
    else if (exp < 100) { s += char('0' + exp / 10); s += char('0' + exp % 10); }
    else { s += char('0' + exp / 100); s += char('0' + (exp / 10) % 10); s += char('0' + exp % 10); }
    return s;
}
// strconv fmtF with prec = max(nd-dp,0) (shortest)
inline std::string layout_f(bool neg, const Digits& g) {
    std::string s; if (neg) s += '-';
    if (g.dp > 0) {
        int m = std::min(g.nd, g.dp);
        s.append(g.d, m);
        for (; m < g.dp; m++) s += '0';
    } else s += '0';
    int prec = std::max(g.nd - g.dp, 0);
    if (prec > 0) {
        s += '.';
        for (int i = 0; i < prec; i++) { int j = g.dp + i; s += (j >= 0 && j < g.nd) ? g.d[j] : '0'; }
    }
    return s;
}

enum FloatFmt { FMT_G_V = 0 /* fmt %v */, FMT_F = 1 /* 'f',-1 */, FMT_JSON = 2 /* encoding/json */ };

template <typename F> inline std::string fmt_float_t(F v, int bits, FloatFmt f) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "+Inf" : "-Inf";
    bool neg = std::signbit(v);
    F a = neg ? -v : v;
    Digits g; g.nd = 0; g.dp = 0;
    if (a != 0) g = (bits == 32) ? shortest_f32((float)a) : shortest_f64((double)a);
    switch (f) {
    case FMT_F: return layout_f(neg, g);
    case FMT_G_V: {
        int exp = g.dp - 1;                       // strconv %g, shortest => eprec = 6
        if (exp < -4 || exp >= 6) return layout_e(neg, g);
        return layout_f(neg, g);
    }
    case FMT_JSON: {
        bool use_e = false;
        if (a != 0) {
            if (bits == 64) use_e = ((double)a < 1e-6 || (double)a >= 1e21);
            else use_e = ((float)a < (float)1e-6 || (float)a >= (float)1e21);
        }
        if (!use_e) return layout_f(neg, g);
        std::string s = layout_e(neg, g);
        size_t n = s.size();                      // clean up e-09 to e-9
        if (n >= 4 && s[n - 4] == 'e' && (s[n - 3] == '-' || s[n - 3] == '+') && s[n - 2] == '0') { s[n - 2] = s[n - 1]; s.pop_back(); }
        return s;
    }
    }
    return "";
}
inline std::string fmt_f64(double v, FloatFmt f) { return fmt_float_t<double>(v, 64, f); }
inline std::string fmt_f32(float v, FloatFmt f) { return fmt_float_t<float>(v, 32, f); }

// ---------------------------------------------------------------- time
// civil date from days since 1970-01-01 (proleptic Gregorian), H. Hinnant's algorithm
inline void civil_from_days(int64_t z, int64_t& y, unsigned& m, unsigned& d) {
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
inline int64_t days_from_civil(int64_t y, unsigned m, unsigned d) {
    y -= m <= 2;
    const int64_t era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m > 2 ? m - 3 : m + 9) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (int64_t)doe - 719468;
}
inline int64_t floor_div(int64_t a, int64_t b) { int64_t q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) q--; return q; }

inline void pad(std::string& s, int64_t v, int w) {
    // Go appendInt(b, x, width): zero padded, '-' for negatives
    if (v < 0) { s += '-'; v = -v; }
    std::string t = fmt_i64(v);
    for (int i = (int)t.size(); i < w; i++) s += '0';
    s += t;
}
// time.Time{sec,nsec}.UTC().Format("2006-01-02")
inline std::string fmt_date_only(int64_t sec) {
    int64_t days = floor_div(sec, 86400); int64_t y; unsigned m, d; civil_from_days(days, y, m, d);
    std::string s; pad(s, y, 4); s += '-'; pad(s, m, 2); s += '-'; pad(s, d, 2); return s;
}
// time.Time.UTC().Format(time.RFC3339Nano)  "2006-01-02T15:04:05.999999999Z07:00"
inline std::string fmt_rfc3339nano_utc(int64_t sec, uint32_t nsec) {
    int64_t days = floor_div(sec, 86400); int64_t sod = sec - days * 86400;
    std::string s = fmt_date_only(sec);
    s += 'T'; pad(s, sod / 3600, 2); s += ':'; pad(s, (sod / 60) % 60, 2); s += ':'; pad(s, sod % 60, 2);
    if (nsec) {
        char b[10]; uint32_t v = nsec; for (int i = 8; i >= 0; i--) { b[i] = char('0' + v % 10); v /= 10; }
        int n = 9; while (n > 0 && b[n - 1] == '0') n--;
        s += '.'; s.append(b, n);
    }
    s += 'Z';
    return s;
}

// time.Duration.String()  (Go time/time.go Duration.String / fmtFrac / fmtInt)
inline std::string fmt_duration(int64_t d) {
    char buf[32]; int w = 32;
    uint64_t u = (uint64_t)d; bool neg = d < 0; if (neg) u = (uint64_t)0 - u;
    auto fmt_frac = [&](uint64_t v, int prec, uint64_t& outv) {
        bool print = false;
        for (int i = 0; i < prec; i++) { int digit = (int)(v % 10); print = print || digit != 0; if (print) { buf[--w] = char('0' + digit); } v /= 10; }
        if (print) buf[--w] = '.';
        outv = v;
    };
    auto fmt_int = [&](uint64_t v) { if (v == 0) buf[--w] = '0'; else while (v > 0) { buf[--w] = char('0' + v % 10); v /= 10; } };
    if (u < 1000000000ULL) {
        int prec;
        buf[--w] = 's';
        if (u == 0) return "0s";
        else if (u < 1000ULL) { prec = 0; buf[--w] = 'n'; }
        else if (u < 1000000ULL) { prec = 3; w--; buf[w] = (char)0xB5; w--; buf[w] = (char)0xC2; }  // "µ" U+00B5
        else { prec = 6; buf[--w] = 'm'; }
        uint64_t v; fmt_frac(u, prec, v); fmt_int(v);
    } else {
        buf[--w] = 's';
        uint64_t v; fmt_frac(u, 9, v);
        fmt_int(v % 60); v /= 60;
        if (v > 0) { buf[--w] = 'm'; fmt_int(v % 60); v /= 60; if (v > 0) { buf[--w] = 'h'; fmt_int(v); } }
    }
    if (neg) buf[--w] = '-';
    return std::string(buf + w, 32 - w);
}

}  // namespace orc
