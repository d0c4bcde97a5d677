// ORACLE — test infrastructure only. Never linked into or called from the product path.
//
// CPU restatement of the reference's generic JSON parser (format "json"):
//   message -> lines                 pkg/parsers/generic/generic_parser.go:519-555 (bufio.ScanLines: split at '\n', one
//                                    trailing '\r' dropped, the unterminated last line IS a line; empty lines skipped
//                                    and do not advance idx)
//   line -> map[string]interface{}   generic_parser.go:672-730 Unmarshal (fastjson parse + per-key extraction by the
//                                    DECLARED column type) and :603-633 wrapIntoEmptyInterface
//   value -> column value            generic_parser.go:888-1123 ParseVal, :792-886 extractTimeValue
//   map -> ChangeItem / unparsed     generic_parser.go:297-404 makeChangeItem (flat keys only), aux columns :115-164
// Third party, NOT under /root/reference (go.mod:70 github.com/valyala/fastjson v1.6.4 incl. fastfloat): its parser and
// number routines are restated from the published source: whitespace = {0x20,\n,\t,\r}; a number token is the maximal
// run of [0-9.+-eE] (or inf/nan), strings end at the first '"' preceded by an even number of backslashes, bad escapes
// are kept verbatim (unescapeStringBestEffort), Get* return 0 on a type mismatch, ParseInt64BestEffort / ParseUint64-
// BestEffort return 0 unless the WHOLE token is a decimal integer (falling back to strconv past 18 characters),
// ParseBestEffort = uint64 mantissa / 10^frac * math.Pow10(exp) with strconv.ParseFloat past 18 integer characters,
// 16 mantissa characters with a fraction, or |exp| > 300.
// PINNED by pkg/parsers/generic/gotest/canondata/result.json (TestParserNumberTypes both modes, TestBase64Unpack) —
// see tests/test_json_parser.py; everything the canon data does not exercise (lenient tokens, the exact fall-back
// thresholds, escapes) is PARITY UNPINNED and says so in DESIGN.md.
// Values the tf_batch layout cannot carry, or that need a Go library this file does not restate, mark the row
// JSN_HOST in BOTH oracle and device (the shim re-parses that line with the Go parser):
//   NaN / Inf or an invalid json.Number inside `any`;  a string value in a `datetime` column (araddon/dateparse);
//   a string value in an `any` column that starts with '{' or 'n' (goccy/go-json re-parse of JSON-in-a-string);
//   a line with a key named like one of the aux columns.
// The device additionally hands over lines it cannot decide (nesting > 24 inside `any`, floats whose correct rounding
// Eisel-Lemire leaves open, numeric strings > 96 bytes); tests/test_json_parser.py covers those by code only.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <string>
#include <string_view>
#include <vector>
#include <algorithm>
#include "go_strconv.hpp"

namespace jsn { std::string go_quote(const uint8_t* s, size_t n); }   // oracle.cpp: encoding/json appendString, escapeHTML = true

namespace jsn {
using sv = std::string_view;

enum JsnErr { JSN_OK = 0, JSN_PARSE = 32 /* fastjson error -> unparsed row :548-553 */, JSN_SKIP = 33 /* valid JSON, not an object or no keys: the line
                 yields nothing :536 */, JSN_NIL_REQUIRED = 34 /* "ParseVal nil" :369-371 */, JSN_PARSEVAL = 35 /* "ParseVal error" :361-366 */,
              JSN_HOST = 36 /* see header */ };

// ------------------------------------------------------------------ fastjson v1.6.4 parser.go
enum FT { F_NULL, F_OBJECT, F_ARRAY, F_STRING /* raw, escapes intact */, F_NUMBER, F_TRUE, F_FALSE };
struct FV { FT t = F_NULL; sv s; std::vector<std::pair<sv, FV>> kvs; std::vector<FV> arr; };
constexpr int MAX_DEPTH = 300;

inline sv skip_ws(sv s) { size_t i = 0; while (i < s.size() && (s[i] == 0x20 || s[i] == 0x0A || s[i] == 0x09 || s[i] == 0x0D)) i++; return s.substr(i); }
inline bool fold_eq(sv a, const char* b) { size_t n = std::strlen(b); if (a.size() != n) return false; for (size_t i = 0; i < n; i++) if ((a[i] | 0x20) != b[i]) return false; return true; }

inline bool parse_raw_string(sv s, sv& out, sv& tail) {       // s starts after the opening quote
    size_t i = 0;
    while (i < s.size()) { if (s[i] == '\\') { i += 2; continue; } if (s[i] == '"') { out = s.substr(0, i); tail = s.substr(i + 1); return true; } i++; }
    return false;                                              // missing closing '"'
}
inline bool parse_raw_number(sv s, sv& out, sv& tail) {
    for (size_t i = 0; i < s.size(); i++) {
        const char ch = s[i];
        if ((ch >= '0' && ch <= '9') || ch == '.' || ch == '-' || ch == 'e' || ch == 'E' || ch == '+') continue;
        if (i == 0 || (i == 1 && (s[0] == '-' || s[0] == '+'))) {
            if (s.size() - i >= 3) { sv xs = s.substr(i, 3); if (fold_eq(xs, "inf") || fold_eq(xs, "nan")) { out = s.substr(0, i + 3); tail = s.substr(i + 3); return true; } }
            return false;
        }
        out = s.substr(0, i); tail = s.substr(i); return true;
    }
    out = s; tail = sv(); return true;
}
bool parse_value(sv s, FV& v, sv& tail, int depth);
inline bool parse_object(sv s, FV& v, sv& tail, int depth) {
    v.t = F_OBJECT; s = skip_ws(s);
    if (s.empty()) return false;
    if (s[0] == '}') { tail = s.substr(1); return true; }
    for (;;) {
        s = skip_ws(s);
        if (s.empty() || s[0] != '"') return false;
        sv k; if (!parse_raw_string(s.substr(1), k, s)) return false;
        s = skip_ws(s);
        if (s.empty() || s[0] != ':') return false;
        s = skip_ws(s.substr(1));
        v.kvs.emplace_back(k, FV());
        FV child; if (!parse_value(s, child, s, depth)) return false;
        v.kvs.back().second = std::move(child);
        s = skip_ws(s);
        if (s.empty()) return false;
        if (s[0] == ',') { s = s.substr(1); continue; }
        if (s[0] == '}') { tail = s.substr(1); return true; }
        return false;
    }
}
inline bool parse_array(sv s, FV& v, sv& tail, int depth) {
    v.t = F_ARRAY; s = skip_ws(s);
    if (s.empty()) return false;
    if (s[0] == ']') { tail = s.substr(1); return true; }
    for (;;) {
        s = skip_ws(s);
        FV child; if (!parse_value(s, child, s, depth)) return false;
        v.arr.push_back(std::move(child));
        s = skip_ws(s);
        if (s.empty()) return false;
        if (s[0] == ',') { s = s.substr(1); continue; }
        if (s[0] == ']') { tail = s.substr(1); return true; }
        return false;
    }
}
inline bool parse_value(sv s, FV& v, sv& tail, int depth) {
    if (s.empty()) return false;
    if (++depth > MAX_DEPTH) return false;
    switch (s[0]) {
    case '{': return parse_object(s.substr(1), v, tail, depth);
    case '[': return parse_array(s.substr(1), v, tail, depth);
    case '"': v.t = F_STRING; return parse_raw_string(s.substr(1), v.s, tail);
    case 't': if (s.substr(0, 4) != "true") return false; v.t = F_TRUE; tail = s.substr(4); return true;
    case 'f': if (s.substr(0, 5) != "false") return false; v.t = F_FALSE; tail = s.substr(5); return true;
    case 'n':
        if (s.substr(0, 4) != "null") { if (s.size() >= 3 && fold_eq(s.substr(0, 3), "nan")) { v.t = F_NUMBER; v.s = s.substr(0, 3); tail = s.substr(3); return true; } return false; }
        v.t = F_NULL; tail = s.substr(4); return true;
    }
    v.t = F_NUMBER; return parse_raw_number(s, v.s, tail);
}
inline bool fj_parse(sv line, FV& v) {                          // Parser.Parse
    sv tail; sv s = skip_ws(line);
    if (!parse_value(s, v, tail, 0)) return false;
    return skip_ws(tail).empty();
}

inline void utf8_put(std::string& b, uint32_t r) {              // string(rune(r))
    if (r > 0x10FFFF || (r >= 0xD800 && r <= 0xDFFF)) r = 0xFFFD;
    if (r < 0x80) b += (char)r;
    else if (r < 0x800) { b += (char)(0xC0 | (r >> 6)); b += (char)(0x80 | (r & 0x3F)); }
    else if (r < 0x10000) { b += (char)(0xE0 | (r >> 12)); b += (char)(0x80 | ((r >> 6) & 0x3F)); b += (char)(0x80 | (r & 0x3F)); }
    else { b += (char)(0xF0 | (r >> 18)); b += (char)(0x80 | ((r >> 12) & 0x3F)); b += (char)(0x80 | ((r >> 6) & 0x3F)); b += (char)(0x80 | (r & 0x3F)); }
}
inline bool hex4(sv s, uint32_t& x) { if (s.size() < 4) return false; x = 0; for (int i = 0; i < 4; i++) { char c = s[i]; int d; if (c >= '0' && c <= '9') d = c - '0'; else if ((c | 0x20) >= 'a' && (c | 0x20) <= 'f') d = (c | 0x20) - 'a' + 10; else return false; x = x * 16 + d; } return true; }
inline std::string unescape_best_effort(sv s) {                 // fastjson unescapeStringBestEffort
    size_t n = s.find('\\');
    if (n == sv::npos) return std::string(s);
    std::string b(s.substr(0, n)); s = s.substr(n + 1);
    while (!s.empty()) {
        const char ch = s[0]; s = s.substr(1);
        switch (ch) {
        case '"': b += '"'; break; case '\\': b += '\\'; break; case '/': b += '/'; break;
        case 'b': b += '\b'; break; case 'f': b += '\f'; break; case 'n': b += '\n'; break; case 'r': b += '\r'; break; case 't': b += '\t'; break;
        case 'u': {
            uint32_t x;
            if (s.size() < 4 || !hex4(s, x)) { b += "\\u"; break; }
            sv xs = s.substr(0, 4); s = s.substr(4);
            if (!(x >= 0xD800 && x <= 0xDFFF)) { utf8_put(b, x); break; }
            uint32_t x1;
            if (s.size() < 6 || s[0] != '\\' || s[1] != 'u' || !hex4(s.substr(2), x1)) { b += "\\u"; b += xs; break; }
            uint32_t r = 0xFFFD;                                 // utf16.DecodeRune
            if (x >= 0xD800 && x < 0xDC00 && x1 >= 0xDC00 && x1 < 0xE000) r = ((x - 0xD800) << 10 | (x1 - 0xDC00)) + 0x10000;
            utf8_put(b, r); s = s.substr(6); break;
        }
        default: b += '\\'; b += ch;
        }
        n = s.find('\\');
        if (n == sv::npos) { b += s; break; }
        b += s.substr(0, n); s = s.substr(n + 1);
    }
    return b;
}
// Value.MarshalTo on a freshly parsed tree: nested strings and keys are still raw, whitespace is gone
inline void fj_marshal(const FV& v, std::string& d) {
    switch (v.t) {
    case F_NULL: d += "null"; break; case F_TRUE: d += "true"; break; case F_FALSE: d += "false"; break;
    case F_NUMBER: d += v.s; break;
    case F_STRING: d += '"'; d += v.s; d += '"'; break;
    case F_OBJECT: d += '{'; for (size_t i = 0; i < v.kvs.size(); i++) { if (i) d += ','; d += '"'; d += v.kvs[i].first; d += "\":"; fj_marshal(v.kvs[i].second, d); } d += '}'; break;
    case F_ARRAY: d += '['; for (size_t i = 0; i < v.arr.size(); i++) { if (i) d += ','; fj_marshal(v.arr[i], d); } d += ']'; break;
    }
}

// ------------------------------------------------------------------ strconv (Go 1.2x)
inline double go_nan() { const uint64_t b = 0x7FF8000000000001ull; double d; std::memcpy(&d, &b, 8); return d; }   // math.NaN()
inline bool underscore_ok(sv s) {                                // strconv/atoi.go underscoreOK
    char i = '^'; size_t p = 0;
    if (!s.empty() && (s[0] == '-' || s[0] == '+')) p = 1;
    bool hex = false;
    if (s.size() - p >= 2 && s[p] == '0' && ((s[p + 1] | 0x20) == 'b' || (s[p + 1] | 0x20) == 'o' || (s[p + 1] | 0x20) == 'x')) { i = '0'; hex = (s[p + 1] | 0x20) == 'x'; p += 2; }
    for (; p < s.size(); p++) {
        const char c = s[p];
        if ((c >= '0' && c <= '9') || (hex && (c | 0x20) >= 'a' && (c | 0x20) <= 'f')) { i = '0'; continue; }
        if (c == '_') { if (i != '0') return false; i = '_'; continue; }
        if (i == '_') return false;
        i = '!';
    }
    return i != '_';
}
// strconv.ParseUint(s, base, bits): rc 0 ok, 1 syntax, 2 range.  base 0 = by prefix, underscores allowed
inline int go_parse_uint(sv s0, int base, int bits, uint64_t& out) {
    if (s0.empty()) return 1;
    sv s = s0; const bool base0 = base == 0;
    if (base == 0) {
        base = 10;
        if (s[0] == '0') {
            if (s.size() >= 3 && (s[1] | 0x20) == 'b') { base = 2; s = s.substr(2); }
            else if (s.size() >= 3 && (s[1] | 0x20) == 'o') { base = 8; s = s.substr(2); }
            else if (s.size() >= 3 && (s[1] | 0x20) == 'x') { base = 16; s = s.substr(2); }
            else { base = 8; s = s.substr(1); }
        }
    }
    const uint64_t maxv = bits == 64 ? ~0ull : ((1ull << bits) - 1);
    bool underscores = false; unsigned __int128 n = 0; bool range = false;
    for (char c : s) {
        int d;
        if (c == '_' && base0) { underscores = true; continue; }
        if (c >= '0' && c <= '9') d = c - '0'; else if ((c | 0x20) >= 'a' && (c | 0x20) <= 'z') d = (c | 0x20) - 'a' + 10; else return 1;
        if (d >= base) return 1;
        if (!range) { n = n * base + d; if (n > maxv) range = true; }
    }
    if (underscores && !underscore_ok(s0)) return 1;
    if (range) { out = maxv; return 2; }
    out = (uint64_t)n; return 0;
}
inline int go_parse_int(sv s, int base, int bits, int64_t& out) {
    if (s.empty()) return 1;
    sv s0 = s; bool neg = false;
    if (s[0] == '+') s = s.substr(1); else if (s[0] == '-') { neg = true; s = s.substr(1); }
    uint64_t un; int rc = go_parse_uint(s, base, 64, un);
    if (rc == 1) return 1;
    if (base == 0 && s.find('_') != sv::npos && !underscore_ok(s0)) return 1;
    const uint64_t cutoff = 1ull << (bits - 1);
    if (rc == 2) return 2;
    if (!neg && un >= cutoff) return 2;
    if (neg && un > cutoff) return 2;
    out = neg ? (int64_t)(0 - un) : (int64_t)un; return 0;
}
// strconv.ParseFloat(s, 64): rc 0 ok, 1 syntax, 2 range (out = +-Inf)
inline int go_parse_float(sv s, double& out) {
    if (s.empty()) return 1;
    {   // special()
        sv t = s; double sign = 1; bool had_sign = false;
        if (t[0] == '+' || t[0] == '-') { sign = t[0] == '-' ? -1 : 1; t = t.substr(1); had_sign = true; }
        if (fold_eq(t, "inf") || fold_eq(t, "infinity")) { out = sign * INFINITY; return 0; }
        if (!had_sign && fold_eq(t, "nan")) { out = go_nan(); return 0; }
    }
    // readFloat syntax
    size_t i = 0; if (s[i] == '+' || s[i] == '-') i++;
    bool hex = false; if (i + 2 < s.size() + 0 && s[i] == '0' && (s[i + 1] | 0x20) == 'x') { hex = true; i += 2; }
    bool sawdot = false, sawdigits = false, underscores = false;
    for (; i < s.size(); i++) {
        const char c = s[i];
        if (c == '_') { underscores = true; continue; }
        if (c == '.') { if (sawdot) break; sawdot = true; continue; }
        if ((c >= '0' && c <= '9') || (hex && (c | 0x20) >= 'a' && (c | 0x20) <= 'f')) { sawdigits = true; continue; }
        break;
    }
    if (!sawdigits) return 1;
    bool sawexp = false;
    if (i < s.size() && (s[i] | 0x20) == (hex ? 'p' : 'e')) {
        i++; if (i >= s.size()) return 1;
        if (s[i] == '+' || s[i] == '-') i++;
        if (i >= s.size() || s[i] < '0' || s[i] > '9') return 1;
        for (; i < s.size() && ((s[i] >= '0' && s[i] <= '9') || s[i] == '_'); i++) if (s[i] == '_') underscores = true;
        sawexp = true;
    }
    if (hex && !sawexp) return 1;
    if (i != s.size()) return 1;
    std::string t(s);
    if (underscores) { if (!underscore_ok(s)) return 1; t.erase(std::remove(t.begin(), t.end(), '_'), t.end()); }
    char* end = nullptr; out = std::strtod(t.c_str(), &end);        // glibc: correctly rounded, like Go
    if (std::isinf(out)) return 2;
    return 0;
}

// ------------------------------------------------------------------ fastjson/fastfloat v1.6.4 parse.go
inline uint64_t ff_uint64_best_effort(sv s) {
    if (s.empty()) return 0;
    size_t i = 0; uint64_t d = 0;
    while (i < s.size() && s[i] >= '0' && s[i] <= '9') {
        d = d * 10 + (uint64_t)(s[i] - '0'); i++;
        if (i > 18) { uint64_t dd; return go_parse_uint(s, 10, 64, dd) == 0 ? dd : 0; }
    }
    if (i == 0 || i < s.size()) return 0;
    return d;
}
inline int64_t ff_int64_best_effort(sv s) {
    if (s.empty()) return 0;
    size_t i = 0; const bool minus = s[0] == '-';
    if (minus) { i++; if (i >= s.size()) return 0; }
    uint64_t d = 0; const size_t j = i;
    while (i < s.size() && s[i] >= '0' && s[i] <= '9') {
        d = d * 10 + (uint64_t)(s[i] - '0'); i++;
        if (i > 18) { int64_t dd; return go_parse_int(s, 10, 64, dd) == 0 ? dd : 0; }
    }
    if (i <= j || i < s.size()) return 0;
    return minus ? -(int64_t)d : (int64_t)d;
}
inline double go_pow10(int n) {                                  // math.Pow10: table product / quotient, one rounding
    static const double tab[32] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22, 1e23, 1e24, 1e25, 1e26, 1e27, 1e28, 1e29, 1e30, 1e31};
    static const double pos32[10] = {1e0, 1e32, 1e64, 1e96, 1e128, 1e160, 1e192, 1e224, 1e256, 1e288};
    static const double neg32[11] = {1e-0, 1e-32, 1e-64, 1e-96, 1e-128, 1e-160, 1e-192, 1e-224, 1e-256, 1e-288, 1e-320};
    if (0 <= n && n <= 308) return pos32[n / 32] * tab[n % 32];
    if (-323 <= n && n <= 0) return neg32[(-n) / 32] / tab[(-n) % 32];
    return n > 0 ? INFINITY : 0.0;
}
inline double ff_slow(sv s) { double f; int rc = go_parse_float(s, f); if (rc == 1) return 0; return f; }   // err != nil && !IsInf -> 0
inline double ff_best_effort(sv s) {
    if (s.empty()) return 0;
    size_t i = 0; const bool minus = s[0] == '-';
    if (minus) { i++; if (i >= s.size()) return 0; }
    if (s[i] == '.' && (i + 1 >= s.size() || s[i + 1] < '0' || s[i + 1] > '9')) return 0;
    uint64_t d = 0; const size_t j = i;
    while (i < s.size() && s[i] >= '0' && s[i] <= '9') { d = d * 10 + (uint64_t)(s[i] - '0'); i++; if (i > 18) return ff_slow(s); }
    if (i <= j && s[i] != '.') {
        sv t = s.substr(i); if (!t.empty() && t[0] == '+') t = t.substr(1);
        if (fold_eq(t, "inf") || fold_eq(t, "infinity")) return minus ? -INFINITY : INFINITY;
        if (fold_eq(t, "nan")) return go_nan();
        return 0;
    }
    double f = (double)d;
    if (i >= s.size()) return minus ? -f : f;
    if (s[i] == '.') {
        i++;
        if (i >= s.size()) return f;                              // (sic) the sign is dropped for "-1."
        const size_t k = i;
        while (i < s.size() && s[i] >= '0' && s[i] <= '9') { d = d * 10 + (uint64_t)(s[i] - '0'); i++; if (i - j >= 17) return ff_slow(s); }
        f = (double)d / go_pow10((int)(i - k));
        if (i >= s.size()) return minus ? -f : f;
    }
    if (s[i] == 'e' || s[i] == 'E') {
        i++; if (i >= s.size()) return 0;
        bool em = false;
        if (s[i] == '+' || s[i] == '-') { em = s[i] == '-'; i++; if (i >= s.size()) return 0; }
        int exp = 0; const size_t j2 = i;
        while (i < s.size() && s[i] >= '0' && s[i] <= '9') { exp = exp * 10 + (s[i] - '0'); i++; if (exp > 300) return ff_slow(s); }
        if (i <= j2) return 0;
        if (em) exp = -exp;
        f *= go_pow10(exp);
        if (i >= s.size()) return minus ? -f : f;
    }
    return 0;
}

// ------------------------------------------------------------------ Go values
struct GV {
    enum K { NIL, STR, BYTES, BOOL, F64, NUM /* json.Number */, I64 /* any signed width */, U64, MAP, ARR, TIME } k = NIL;
    std::string s; double f = 0; int64_t i = 0; uint64_t u = 0; bool b = false;
    std::vector<std::pair<std::string, GV>> m; std::vector<GV> a;
};
inline void map_set(std::vector<std::pair<std::string, GV>>& m, const std::string& k, GV v) { for (auto& kv : m) if (kv.first == k) { kv.second = std::move(v); return; } m.emplace_back(k, std::move(v)); }

inline bool valid_json_number(sv s) {                            // encoding/json isValidNumber
    size_t i = 0; if (s.empty()) return false;
    if (s[i] == '-') { i++; if (i == s.size()) return false; }
    if (s[i] == '0') i++; else if (s[i] >= '1' && s[i] <= '9') { while (i < s.size() && s[i] >= '0' && s[i] <= '9') i++; } else return false;
    if (i + 1 < s.size() + 0 && s[i] == '.' && s[i + 1] >= '0' && s[i + 1] <= '9') { i += 2; while (i < s.size() && s[i] >= '0' && s[i] <= '9') i++; }
    if (i + 1 < s.size() + 0 && (s[i] == 'e' || s[i] == 'E')) {
        i++; if (s[i] == '+' || s[i] == '-') { i++; if (i == s.size()) return false; }
        while (i < s.size() && s[i] >= '0' && s[i] <= '9') i++;
    }
    return i == s.size();
}
// encoding/json Marshal of the value (maps sorted by key, HTML escaping on). false: json.Marshal would fail.
inline bool go_marshal(const GV& v, std::string& d) {
    switch (v.k) {
    case GV::NIL: d += "null"; return true;
    case GV::STR: d += go_quote((const uint8_t*)v.s.data(), v.s.size()); return true;
    case GV::BOOL: d += v.b ? "true" : "false"; return true;
    case GV::F64: if (std::isnan(v.f) || std::isinf(v.f)) return false; d += orc::fmt_f64(v.f, orc::FMT_JSON); return true;
    case GV::NUM: if (!valid_json_number(v.s)) return false; d += v.s; return true;     // "" would marshal as 0; fastjson never yields an empty token
    case GV::I64: d += orc::fmt_i64(v.i); return true;
    case GV::U64: d += orc::fmt_u64(v.u); return true;
    case GV::ARR: d += '['; for (size_t i = 0; i < v.a.size(); i++) { if (i) d += ','; if (!go_marshal(v.a[i], d)) return false; } d += ']'; return true;
    case GV::MAP: {
        std::vector<const std::pair<std::string, GV>*> ks; for (auto& kv : v.m) ks.push_back(&kv);
        std::sort(ks.begin(), ks.end(), [](auto a, auto b) { return a->first < b->first; });
        d += '{';
        for (size_t i = 0; i < ks.size(); i++) { if (i) d += ','; d += go_quote((const uint8_t*)ks[i]->first.data(), ks[i]->first.size()); d += ':'; if (!go_marshal(ks[i]->second, d)) return false; }
        d += '}'; return true;
    }
    default: return false;
    }
}
inline GV wrap(const FV& v, bool use_numbers) {                  // wrapIntoEmptyInterface :603-633
    GV g;
    switch (v.t) {
    case F_OBJECT: g.k = GV::MAP; for (auto& kv : v.kvs) map_set(g.m, unescape_best_effort(kv.first), wrap(kv.second, use_numbers)); break;
    case F_ARRAY: g.k = GV::ARR; for (auto& x : v.arr) g.a.push_back(wrap(x, use_numbers)); break;
    case F_STRING: g.k = GV::STR; g.s = unescape_best_effort(v.s); break;
    case F_TRUE: g.k = GV::BOOL; g.b = true; break; case F_FALSE: g.k = GV::BOOL; g.b = false; break;
    case F_NUMBER: if (use_numbers) { g.k = GV::NUM; g.s = std::string(v.s); } else { g.k = GV::F64; g.f = ff_best_effort(v.s); } break;
    default: break;
    }
    return g;
}

// ------------------------------------------------------------------ parser configuration
struct Field { std::string name; int tf; bool key = false, required = false; };
struct Opts {
    bool add_rest = false, add_dedupe_keys = false, null_keys_allowed = false, use_numbers_in_any = false, unpack_bytes_base64 = false;
    std::string partition;                                        // abstract.Partition.String() of the batch
};
struct Msg { uint64_t end; uint64_t offset; int64_t write_sec; uint32_t write_nsec; };   // end = byte offset one past the message in the buffer

struct ColOut { std::vector<uint8_t> values, heap, aux, valid; std::vector<uint32_t> offs{0}; bool any_null = false; };
struct Result { std::vector<ColOut> cols; std::vector<tf_rowerr> errs; uint64_t rows = 0, lines = 0; };

inline int base64_std_decode(sv s, std::string& out) {           // encoding/base64 StdEncoding.DecodeString; 0 ok
    auto dv = [](unsigned char c) -> int { if (c >= 'A' && c <= 'Z') return c - 'A'; if (c >= 'a' && c <= 'z') return c - 'a' + 26; if (c >= '0' && c <= '9') return c - '0' + 52; if (c == '+') return 62; if (c == '/') return 63; return -1; };
    size_t i = 0; bool end = false;
    while (!end) {
        int db[4]; int j = 0; int dlen = 4;
        while (j < 4) {
            if (i == s.size()) { if (j == 0) return 0; return 1; }
            const unsigned char c = (unsigned char)s[i++];
            const int v = dv(c);
            if (v >= 0) { db[j++] = v; continue; }
            if (c == '\n' || c == '\r') continue;
            if (c != '=') return 1;
            if (j < 2) return 1;
            if (j == 2) { while (i < s.size() && (s[i] == '\n' || s[i] == '\r')) i++; if (i == s.size() || s[i] != '=') return 1; i++; }
            while (i < s.size() && (s[i] == '\n' || s[i] == '\r')) i++;
            if (i < s.size()) return 1;
            dlen = j; end = true; break;
        }
        for (int k = dlen; k < 4; k++) db[k] = 0;
        const uint32_t val = (uint32_t)db[0] << 18 | (uint32_t)db[1] << 12 | (uint32_t)db[2] << 6 | (uint32_t)db[3];
        out += (char)(val >> 16); if (dlen >= 3) out += (char)(val >> 8); if (dlen == 4) out += (char)val;
    }
    return 0;
}
inline int64_t go_f64_to_i64(double f) { if (!(f >= -9223372036854775808.0 && f < 9223372036854775808.0)) return INT64_MIN; return (int64_t)f; }   // amd64 CVTTSD2SQ
inline uint64_t go_f64_to_u64(double f) {                        // amd64 Go: via int64 when < 2^63, else int64(f - 2^63) ^ 1<<63
    if (f < 9223372036854775808.0) return (uint64_t)go_f64_to_i64(f);
    return (uint64_t)go_f64_to_i64(f - 9223372036854775808.0) ^ 0x8000000000000000ull;
}

inline bool is_int_tf(int tf) { return tf == TF_INT8 || tf == TF_INT16 || tf == TF_INT32 || tf == TF_INT64; }
inline bool is_uint_tf(int tf) { return tf == TF_UINT8 || tf == TF_UINT16 || tf == TF_UINT32 || tf == TF_UINT64; }
inline int bits_tf(int tf) { switch (tf) { case TF_INT8: case TF_UINT8: return 8; case TF_INT16: case TF_UINT16: return 16; case TF_INT32: case TF_UINT32: return 32; default: return 64; } }
inline int64_t trunc_i(int64_t v, int bits) { switch (bits) { case 8: return (int8_t)v; case 16: return (int16_t)v; case 32: return (int32_t)v; default: return v; } }
inline uint64_t trunc_u(uint64_t v, int bits) { switch (bits) { case 8: return (uint8_t)v; case 16: return (uint16_t)v; case 32: return (uint32_t)v; default: return v; } }

// Unmarshal's per-key extraction (:690-724) for a non-null, non-string fastjson value and the declared type
inline GV extract(const FV& v, int tf, bool use_numbers) {
    GV g;
    const bool num = v.t == F_NUMBER;
    if (tf == TF_UTF8 || tf == TF_BYTES) { g.k = GV::STR; fj_marshal(v, g.s); return g; }                 // v.String()
    if (tf == TF_DOUBLE) { g.k = GV::F64; g.f = num ? ff_best_effort(v.s) : 0; return g; }
    if (tf == TF_BOOLEAN) { g.k = GV::BOOL; g.b = v.t == F_TRUE; return g; }
    if (is_int_tf(tf)) { g.k = GV::I64; g.i = trunc_i(num ? ff_int64_best_effort(v.s) : 0, bits_tf(tf)); return g; }
    if (is_uint_tf(tf)) { g.k = GV::U64; g.u = trunc_u(num ? ff_uint64_best_effort(v.s) : 0, bits_tf(tf)); return g; }
    return wrap(v, use_numbers);
}

// ParseVal :888-1123.  rc 0 ok (out may be NIL), JSN_PARSEVAL, JSN_HOST
inline int parse_val(const GV& v, int tf, const Opts& o, GV& out) {
    out = GV();
    if (tf == TF_DATETIME) {                                       // :889-898 extractTimeValue
        switch (v.k) {
        case GV::NIL: return 0;
        case GV::STR: return JSN_HOST;
        case GV::NUM: { int64_t n; if (go_parse_int(v.s, 10, 64, n)) return JSN_PARSEVAL; out.k = GV::TIME; out.i = n; return 0; }
        case GV::F64: out.k = GV::TIME; out.i = go_f64_to_i64(std::fabs(v.f)); return 0;
        case GV::I64: out.k = GV::TIME; out.i = v.i; return 0;
        case GV::U64: out.k = GV::TIME; out.i = (int64_t)v.u; return 0;
        default: return JSN_PARSEVAL;
        }
    }
    if (v.k == GV::F64) {                                          // :900-925
        if (tf == TF_DOUBLE) { out = v; return 0; }
        if (is_int_tf(tf)) { out.k = GV::I64; out.i = trunc_i(go_f64_to_i64(v.f), bits_tf(tf)); return 0; }
        if (is_uint_tf(tf)) { out.k = GV::U64; out.u = tf == TF_UINT64 ? go_f64_to_u64(v.f) : trunc_u((uint64_t)go_f64_to_i64(v.f), bits_tf(tf)); return 0; }
        if (tf == TF_UTF8 || tf == TF_BYTES) { out.k = GV::STR; out.s = orc::fmt_f64(v.f, orc::FMT_G_V); return 0; }
        out = v; return 0;
    }
    if (v.k == GV::U64) {                                          // :927-948 (only a uint64 column yields a Go uint64)
        out = v; return 0;
    }
    if (v.k == GV::NUM) {                                          // :950-1011
        if (tf == TF_DOUBLE) { out.k = GV::F64; out.f = ff_best_effort(v.s); return 0; }      // fastfloat.Parse; tokens here always parse
        if (is_int_tf(tf) || (is_uint_tf(tf) && tf != TF_UINT64)) { int64_t n; if (go_parse_int(v.s, 10, 64, n)) return JSN_PARSEVAL; if (is_int_tf(tf)) { out.k = GV::I64; out.i = trunc_i(n, bits_tf(tf)); } else { out.k = GV::U64; out.u = trunc_u((uint64_t)n, bits_tf(tf)); } return 0; }
        if (tf == TF_UINT64) { uint64_t n; if (go_parse_uint(v.s, 10, 64, n)) return JSN_PARSEVAL; out.k = GV::U64; out.u = n; return 0; }
        if (tf == TF_UTF8 || tf == TF_BYTES) { out.k = GV::STR; out.s = v.s; return 0; }
        if (tf == TF_ANY) { out = v; return 0; }
        out.k = GV::F64; out.f = ff_best_effort(v.s); return 0;
    }
    if (v.k == GV::STR) {                                          // :1013-1098
        const sv s = v.s;
        if (tf == TF_DOUBLE) { double f; if (go_parse_float(s, f)) return JSN_PARSEVAL; out.k = GV::F64; out.f = f; return 0; }
        if (tf == TF_BOOLEAN) { bool b; if (orc::go_parse_bool((const uint8_t*)s.data(), s.size(), b)) return JSN_PARSEVAL; out.k = GV::BOOL; out.b = b; return 0; }
        if (is_int_tf(tf)) { int64_t n; if (go_parse_int(s, 0, bits_tf(tf), n)) return JSN_PARSEVAL; out.k = GV::I64; out.i = n; return 0; }
        if (is_uint_tf(tf)) { uint64_t n; if (go_parse_uint(s, 0, bits_tf(tf), n)) return JSN_PARSEVAL; out.k = GV::U64; out.u = n; return 0; }
        if (tf == TF_BYTES) { if (o.unpack_bytes_base64) { out.k = GV::BYTES; if (base64_std_decode(s, out.s)) return JSN_PARSEVAL; return 0; } out = v; return 0; }
        if (tf == TF_ANY) {
            std::string r; for (size_t i = 0; i < s.size();) { if (s[i] == '\\' && i + 1 < s.size() && s[i + 1] == '\\') { r += '\\'; i += 2; } else r += s[i++]; }   // strings.ReplaceAll(vv, `\\`, `\`)
            // json.Unmarshal([]byte(vv), &map[string]interface{}) (goccy/go-json v0.10.5, not restated) can only succeed on an
            // object or on `null`: those lines go to the host parser; every other text stays a string
            { size_t q = 0; while (q < r.size() && (r[q] == ' ' || r[q] == '\t' || r[q] == '\r' || r[q] == '\n')) q++; if (q < r.size() && (r[q] == '{' || r[q] == 'n')) return JSN_HOST; }
            out.k = GV::STR; out.s = r; return 0;
        }
        out = v; return 0;
    }
    out = v; return 0;                                             // :1100-1122 (timestamp / interval need Go int64 inputs that JSON never produces)
}

inline int width_tf(int tf) { switch (tf) { case TF_INT8: case TF_UINT8: case TF_BOOLEAN: return 1; case TF_INT16: case TF_UINT16: return 2; case TF_INT32: case TF_UINT32: case TF_FLOAT: return 4; case TF_BYTES: case TF_UTF8: case TF_ANY: return 0; default: return 8; } }

// One line -> cells of the declared fields (+ _rest).  rc 0 row, else JSN_*; err_col = the field that raised it
struct Row { std::vector<GV> cells; std::string rest; };
inline int parse_line(sv line, const std::vector<Field>& all_cols, size_t nfields, const Opts& o, Row& row, int& err_col) {
    err_col = 0;
    FV root; if (!fj_parse(line, root)) return JSN_PARSE;
    std::vector<std::pair<std::string, GV>> item;
    if (root.t == F_OBJECT) for (auto& kv : root.kvs) {              // v.GetObject().Visit :680-726
        std::string k = unescape_best_effort(kv.first);
        const FV& v = kv.second; GV g;
        if (v.t == F_NULL) g.k = GV::NIL;
        else if (v.t == F_STRING) { g.k = GV::STR; g.s = unescape_best_effort(v.s); }
        else { int tf = 0; for (auto& c : all_cols) if (c.name == k) tf = c.tf;      // colTypeMap: the LAST column of that name wins
               g = extract(v, tf, o.use_numbers_in_any); }
        map_set(item, k, std::move(g));
    }
    if (item.empty()) return JSN_SKIP;                              // :536 len(item) > 0
    // a key named like an aux column is extracted with THAT column's type (colTypeMap covers the whole result schema, :1226-1233)
    // and lands in `_rest` as such; rare enough that oracle and device both hand the line to the host parser
    for (auto& kv : item) for (size_t c = nfields; c < all_cols.size(); c++) if (all_cols[c].name == kv.first) { err_col = (int)nfields; return JSN_HOST; }
    row.cells.assign(nfields, GV());
    for (size_t f = 0; f < nfields; f++) {                          // :325-376 (non-nested keys)
        const Field& fd = all_cols[f];
        const GV* raw = nullptr; for (auto& kv : item) if (kv.first == fd.name) raw = &kv.second;
        GV nil; GV out; const int rc = parse_val(raw ? *raw : nil, fd.tf, o, out);
        if (rc == JSN_HOST) { err_col = (int)f; return JSN_HOST; }
        if (rc) { if ((!o.null_keys_allowed && fd.key) || fd.required) { err_col = (int)f; return JSN_PARSEVAL; } continue; }
        if (out.k == GV::NIL && (fd.key || fd.required) && !o.null_keys_allowed) { err_col = (int)f; return JSN_NIL_REQUIRED; }
        if (fd.tf == TF_ANY && out.k != GV::NIL && out.k != GV::STR) { std::string t; if (!go_marshal(out, t)) { err_col = (int)f; return JSN_HOST; } }
        map_set(item, fd.name, out);                                 // :372 item[key.ColumnName] = v
        row.cells[f] = std::move(out);
    }
    if (o.add_rest) {                                               // :377-386
        GV rest; rest.k = GV::MAP;
        for (auto& kv : item) { bool known = false; for (size_t f = 0; f < nfields; f++) if (all_cols[f].name == kv.first) known = true; if (!known) rest.m.push_back(kv); }
        row.rest.clear(); if (!go_marshal(rest, row.rest)) { err_col = (int)nfields; return JSN_HOST; }
    }
    return 0;
}

// DoBatch over a buffer of concatenated messages -> columns of the parser's result schema
//   all_cols = declared fields, then `_rest` (if add_rest), then _timestamp,_partition,_offset,_idx (if add_dedupe_keys)
inline Result parse(const uint8_t* buf, uint64_t len, const std::vector<Msg>& msgs, const std::vector<Field>& all_cols, const Opts& o) {
    Result R; const size_t nc = all_cols.size(); R.cols.resize(nc);
    const size_t naux = (o.add_rest ? 1 : 0) + (o.add_dedupe_keys ? 4 : 0);
    const size_t nf = nc - naux;
    uint64_t mstart = 0; uint64_t lineno = 0;
    auto put_fixed = [&](ColOut& c, int w, uint64_t v) { for (int k = 0; k < w; k++) c.values.push_back((uint8_t)(v >> (8 * k))); };
    auto put_valid = [&](ColOut& c, uint64_t r, bool ok) { if (c.valid.size() < r / 8 + 1) c.valid.resize(r / 8 + 1, 0); if (ok) c.valid[r / 8] |= (uint8_t)(1u << (r % 8)); else c.any_null = true; };
    for (const Msg& m : msgs) {
        const uint64_t mend = m.end <= len ? m.end : len;
        uint64_t p = mstart; uint32_t idx = 0;
        while (p < mend) {
            uint64_t q = p; while (q < mend && buf[q] != '\n') q++;
            uint64_t e = q; if (e > p && buf[e - 1] == '\r') e--;        // bufio.ScanLines dropCR
            const sv line((const char*)buf + p, e - p);
            p = q < mend ? q + 1 : mend;
            if (line.empty()) continue;
            idx++; const uint64_t ln = lineno++;
            Row row; int ecol = 0; const int rc = parse_line(line, all_cols, nf, o, row, ecol);
            if (rc) { R.errs.push_back(tf_rowerr{(uint32_t)ln, (uint16_t)rc, (uint16_t)ecol}); continue; }
            const uint64_t r = R.rows++;
            for (size_t c = 0; c < nc; c++) {
                ColOut& oc = R.cols[c]; const int tf = all_cols[c].tf; const int w = width_tf(tf);
                GV g;
                if (c < nf) g = row.cells[c];
                else {
                    size_t a = c - nf;
                    if (o.add_rest) { if (a == 0) { g.k = GV::MAP; g.s = row.rest; } a--; }
                    if (o.add_dedupe_keys && c >= nf + (o.add_rest ? 1 : 0)) {
                        if (a == 0) { g.k = GV::TIME; g.i = m.write_sec; g.u = m.write_nsec; }
                        else if (a == 1) { g.k = GV::STR; g.s = o.partition; }
                        else if (a == 2) { g.k = GV::U64; g.u = m.offset; }
                        else { g.k = GV::U64; g.u = idx; }
                    }
                }
                const bool null = g.k == GV::NIL;
                put_valid(oc, r, !null);
                if (w) {
                    uint64_t v = 0; uint32_t nsec = 0;
                    switch (g.k) { case GV::I64: v = (uint64_t)g.i; break; case GV::U64: v = g.u; break; case GV::BOOL: v = g.b; break;
                                   case GV::F64: std::memcpy(&v, &g.f, 8); break; case GV::TIME: v = (uint64_t)g.i; nsec = (uint32_t)g.u; break; default: break; }
                    put_fixed(oc, w, v);
                    if (tf == TF_DATE || tf == TF_DATETIME || tf == TF_TIMESTAMP) { for (int k = 0; k < 4; k++) oc.aux.push_back((uint8_t)(nsec >> (8 * k))); }
                } else {
                    std::string text; uint8_t tag = 0;
                    if (tf == TF_ANY) {
                        if (null) text = "";
                        else if (g.k == GV::STR) { text = g.s; tag = 1; }
                        else if (g.k == GV::MAP && c >= nf) text = g.s;                        // _rest, already marshalled
                        else go_marshal(g, text);
                        oc.aux.push_back(tag);
                    } else if (!null) text = g.s;
                    oc.heap.insert(oc.heap.end(), text.begin(), text.end());
                    oc.offs.push_back((uint32_t)oc.heap.size());
                }
            }
        }
        mstart = mend;
    }
    R.lines = lineno;
    return R;
}

}  // namespace jsn
