// ORACLE — test infrastructure only. Never linked into or called from the product path.
//
// CPU restatement of the reference's Debezium parser for messages that carry their schema (or a schema-registry frame):
//   message -> events                pkg/parsers/registry/debezium/engine/parser.go:34-98 DoOne / DoBuf (empty message and a
//                                    wrong magic byte are unparsed rows; 0x00|u32be id|payload frames split at the next 0x00)
//   {"schema":..,"payload":..}       pkg/debezium/unpacker/include_schema.go:13-25 (encoding/json Unmarshal into two RawMessages)
//   payload -> struct                pkg/debezium/common/debezium_schema.go:31-66 (encoding/json Decoder with UseNumber into
//                                    Payload{After,Before map; Op; Source{lsn,ts_ms uint64, txId uint32, ...}; TSMs})
//   op -> kind                       pkg/debezium/kind.go:33-46
//   schema -> columns                pkg/debezium/receiver.go:46-96, receiver_engine.go:104-141 (PrimaryKey = !optional)
//   value -> column value            receiver.go:98-120,186-217; receiver_engine.go:143-330 (extractVal / convertVal);
//                                    default receivers pkg/debezium/common/field_receiver_default.go:15-330
//                                    (Decimal / VariableScaleDecimal through typeutil.Base64ToNumeric helpers.go:972-998, Point "(x,y)")
// encoding/json is the Go standard library: its grammar (RFC 8259, no control characters in strings, valid escapes only,
// invalid UTF-8 replaced by U+FFFD when a string is unquoted), case-insensitive struct key matching and "last duplicate
// wins" are restated here from its documented behaviour.
// PINNED by pkg/parsers/registry/debezium/engine/parser_test.jsonl + gotest/canondata/result.json (59 columns of one pg
// message: every default receiver incl. Bits, Decimal, VariableScaleDecimal, Point) — tests/test_debezium.py.
// Out of scope (both oracle and device return DBZ_HOST so the shim runs the Go parser on that message): database specific
// receivers selected by `__dt_original_type_info` / original types, array / map fields, `__debezium_unavailable_value`
// cells (they make the row ragged), object keys that match a struct field only case-insensitively or need unescaping.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <vector>
#include "json_oracle.hpp"

namespace dbz {
using sv = std::string_view;

enum DbzErr { DBZ_OK = 0, DBZ_UNPARSED = 48 /* the reference emits an `_unparsed` row for the event */, DBZ_HOST = 49 /* see header */,
              DBZ_OTHER_SCHEMA = 50 /* the embedded schema (or registry id) is not the one this plan was built for */,
              DBZ_OTHER_TABLE = 51 /* source.schema / source.table differ from the plan's table */ };

// ------------------------------------------------------------------ encoding/json
struct JV { enum T { NUL, BOOL, NUM, STR, ARR, OBJ } t = NUL; bool b = false; std::string s; std::vector<std::pair<std::string, JV>> kv; std::vector<JV> a; sv raw; bool key_escaped = false; };

struct GoJson {
    sv s; size_t p = 0; bool odd_key = false;       // odd_key: some object key held an escape or a non-ASCII byte
    void ws() { while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\r' || s[p] == '\n')) p++; }
    bool str(std::string& out, bool* odd) {
        if (p >= s.size() || s[p] != '"') return false; p++;
        while (p < s.size()) {
            const unsigned char c = (unsigned char)s[p];
            if (c == '"') { p++; return true; }
            if (c < 0x20) return false;
            if (c >= 0x80) {                                       // unquote: invalid UTF-8 -> U+FFFD, one byte at a time
                if (odd) *odd = true;
                size_t w = 0; uint32_t r = 0xFFFD;
                const size_t n = s.size() - p; const unsigned char* q = (const unsigned char*)s.data() + p;
                if (c >= 0xC2 && c <= 0xDF && n >= 2 && (q[1] & 0xC0) == 0x80) { r = ((c & 0x1Fu) << 6) | (q[1] & 0x3Fu); w = 2; }
                else if (c >= 0xE0 && c <= 0xEF && n >= 3 && (q[1] & 0xC0) == 0x80 && (q[2] & 0xC0) == 0x80) { uint32_t t = ((c & 0x0Fu) << 12) | ((q[1] & 0x3Fu) << 6) | (q[2] & 0x3Fu); if (t >= 0x800 && !(t >= 0xD800 && t <= 0xDFFF)) { r = t; w = 3; } }
                else if (c >= 0xF0 && c <= 0xF4 && n >= 4 && (q[1] & 0xC0) == 0x80 && (q[2] & 0xC0) == 0x80 && (q[3] & 0xC0) == 0x80) { uint32_t t = ((c & 0x07u) << 18) | ((q[1] & 0x3Fu) << 12) | ((q[2] & 0x3Fu) << 6) | (q[3] & 0x3Fu); if (t >= 0x10000 && t <= 0x10FFFF) { r = t; w = 4; } }
                if (!w) { jsn::utf8_put(out, 0xFFFD); p++; } else { out.append(s.substr(p, w)); p += w; }
                continue;
            }
            if (c != '\\') { out += (char)c; p++; continue; }
            if (odd) *odd = true;
            p++; if (p >= s.size()) return false;
            switch (s[p]) {
            case '"': out += '"'; p++; break; case '\\': out += '\\'; p++; break; case '/': out += '/'; p++; break;
            case 'b': out += '\b'; p++; break; case 'f': out += '\f'; p++; break; case 'n': out += '\n'; p++; break; case 'r': out += '\r'; p++; break; case 't': out += '\t'; p++; break;
            case 'u': {
                uint32_t x; if (!jsn::hex4(s.substr(p + 1), x)) return false; p += 5;
                if (x >= 0xD800 && x < 0xDC00) { uint32_t y; if (p + 6 <= s.size() && s[p] == '\\' && s[p + 1] == 'u' && jsn::hex4(s.substr(p + 2), y) && y >= 0xDC00 && y < 0xE000) { x = (((x - 0xD800) << 10) | (y - 0xDC00)) + 0x10000; p += 6; } else x = 0xFFFD; }
                else if (x >= 0xDC00 && x < 0xE000) x = 0xFFFD;
                jsn::utf8_put(out, x); break;
            }
            default: return false;
            }
        }
        return false;
    }
    bool val(JV& g, int depth = 0) {
        ws(); if (p >= s.size() || depth > 10000) return false;
        const size_t start = p; const char c = s[p]; bool ok = false;
        if (c == '{') {
            g.t = JV::OBJ; p++; ws();
            if (p < s.size() && s[p] == '}') { p++; ok = true; }
            else for (;;) {
                ws(); std::string k; bool odd = false; if (!str(k, &odd)) break; if (odd) odd_key = true; ws(); if (p >= s.size() || s[p] != ':') break; p++;
                g.kv.emplace_back(std::move(k), JV()); JV ch; if (!val(ch, depth + 1)) break; g.kv.back().second = std::move(ch); ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == '}') { p++; ok = true; }
                break;
            }
        } else if (c == '[') {
            g.t = JV::ARR; p++; ws();
            if (p < s.size() && s[p] == ']') { p++; ok = true; }
            else for (;;) {
                JV ch; if (!val(ch, depth + 1)) break; g.a.push_back(std::move(ch)); ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == ']') { p++; ok = true; }
                break;
            }
        } else if (c == '"') { g.t = JV::STR; ok = str(g.s, nullptr); }
        else if (s.substr(p, 4) == "true") { g.t = JV::BOOL; g.b = true; p += 4; ok = true; }
        else if (s.substr(p, 5) == "false") { g.t = JV::BOOL; g.b = false; p += 5; ok = true; }
        else if (s.substr(p, 4) == "null") { g.t = JV::NUL; p += 4; ok = true; }
        else {
            size_t q = p; while (q < s.size() && ((s[q] >= '0' && s[q] <= '9') || s[q] == '-' || s[q] == '+' || s[q] == '.' || s[q] == 'e' || s[q] == 'E')) q++;
            sv n = s.substr(p, q - p);
            if (jsn::valid_json_number(n)) { g.t = JV::NUM; g.s = std::string(n); p = q; ok = true; }
        }
        if (ok) g.raw = s.substr(start, p - start);
        return ok;
    }
};
inline const JV* get(const JV& o, const char* k) { const JV* hit = nullptr; if (o.t == JV::OBJ) for (auto& kv : o.kv) if (kv.first == k) hit = &kv.second; return hit; }   // last duplicate wins
inline bool ascii_fold_eq(const std::string& a, const char* b) { size_t n = std::strlen(b); if (a.size() != n) return false; for (size_t i = 0; i < n; i++) { char x = a[i], y = b[i]; if (x >= 'A' && x <= 'Z') x += 32; if (y >= 'A' && y <= 'Z') y += 32; if (x != y) return false; } return true; }
// does the object hold a key that encoding/json would match to `name` although it is not spelled exactly so?
inline bool has_folded_only(const JV& o, std::initializer_list<const char*> names) {
    if (o.t != JV::OBJ) return false;
    for (auto& kv : o.kv) for (const char* n : names) if (kv.first != n && ascii_fold_eq(kv.first, n)) return true;
    return false;
}

// ------------------------------------------------------------------ plan: the `after` / `before` field list
enum Recv { R_INT8 = 1, R_INT16, R_INT32, R_INT64, R_BOOL, R_STRING, R_F64, R_BYTES, R_DECIMAL, R_POINT, R_VSD };
struct Field { std::string name; int recv; int scale = 0; bool key = false; };      // key = !optional (receiver_engine.go:110)

// typeutil.Base64ToNumeric (helpers.go:972-998). rc 0 ok, 1 error, 2 the reference would panic (empty buffer / negative scale)
inline int base64_to_numeric(sv b64, int scale, std::string& out) {
    std::string buf; if (jsn::base64_std_decode(b64, buf)) return 1;
    if (buf.empty() || scale < 0) return 2;
    const bool neg = (unsigned char)buf[0] & 0x80;
    std::vector<uint8_t> mag(buf.begin(), buf.end());
    if (neg) {                                                     // makeNegativeNum: complement the bytes, add one
        for (auto& x : mag) x = (uint8_t)~x;
        for (size_t i = mag.size(); i-- > 0;) { if (++mag[i] != 0) break; }
    }
    std::string digits;                                            // big.Int.String(): repeated division by 10
    std::vector<uint8_t> cur = mag;
    for (;;) {
        size_t first = 0; while (first < cur.size() && cur[first] == 0) first++;
        if (first == cur.size()) break;
        uint32_t rem = 0;
        for (size_t i = first; i < cur.size(); i++) { const uint32_t v = rem * 256 + cur[i]; cur[i] = (uint8_t)(v / 10); rem = v % 10; }
        digits += (char)('0' + rem);
    }
    if (digits.empty()) digits = "0";
    std::string r(digits.rbegin(), digits.rend());
    if (r == "0") { out = r; return 0; }
    if (scale != 0) {
        if ((size_t)scale > r.size()) r = std::string((size_t)scale - r.size() + 1, '0') + r;
        r = r.substr(0, r.size() - (size_t)scale) + "." + r.substr(r.size() - (size_t)scale);
    }
    out = (neg ? "-" : "") + r; return 0;
}

// One cell: rc 0 ok (null -> is_null), DBZ_UNPARSED (receiver error), DBZ_HOST
struct Cell { bool is_null = true; int64_t i = 0; double f = 0; std::string s; };
inline int receive(const Field& fd, const JV& v, Cell& c) {
    c = Cell();
    if (v.t == JV::NUL) return 0;                                                         // receiveField :143-146
    if (v.t == JV::STR && v.s == "__debezium_unavailable_value") return DBZ_HOST;          // :147-151 absent column
    c.is_null = false;
    switch (fd.recv) {
    case R_INT8: case R_INT16: case R_INT32: case R_INT64: {                              // extractVal :213-231 + Do
        if (v.t != JV::NUM) return DBZ_UNPARSED;                                           // "assert no one value extracted"
        int64_t n; if (jsn::go_parse_int(v.s, 10, 64, n)) return DBZ_UNPARSED;
        c.i = fd.recv == R_INT8 ? (int8_t)n : fd.recv == R_INT16 ? (int16_t)n : fd.recv == R_INT32 ? (int32_t)n : n; return 0;
    }
    case R_BOOL: if (v.t != JV::BOOL) return DBZ_UNPARSED; c.i = v.b; return 0;
    case R_STRING: if (v.t != JV::STR && v.t != JV::NUM) return DBZ_UNPARSED; c.s = v.s; return 0;     // :240-251 json.Number -> its text
    case R_F64: { if (v.t != JV::NUM) return DBZ_UNPARSED; double f; if (jsn::go_parse_float(v.s, f)) return DBZ_UNPARSED; c.f = f; return 0; }
    case R_BYTES: if (v.t != JV::STR && v.t != JV::NUM) return DBZ_UNPARSED; if (jsn::base64_std_decode(v.s, c.s)) return DBZ_UNPARSED; return 0;
    case R_DECIMAL: { if (v.t != JV::STR && v.t != JV::NUM) return DBZ_UNPARSED; const int rc = base64_to_numeric(v.s, fd.scale, c.s); return rc == 2 ? DBZ_HOST : rc ? DBZ_UNPARSED : 0; }
    case R_POINT: {                                                                        // Point.Do: "(%v,%v)" of vv["x"], vv["y"]
        if (v.t != JV::OBJ) return DBZ_HOST;                                               // type assertion panic in the reference
        const JV* x = get(v, "x"); const JV* y = get(v, "y"); if (!x || !y) return DBZ_UNPARSED;
        auto pv = [](const JV& a, std::string& o) -> bool { switch (a.t) { case JV::NUM: case JV::STR: o += a.s; return true; case JV::NUL: o += "<nil>"; return true; case JV::BOOL: o += a.b ? "true" : "false"; return true; default: return false; } };
        c.s = "("; if (!pv(*x, c.s)) return DBZ_HOST; c.s += ','; if (!pv(*y, c.s)) return DBZ_HOST; c.s += ')'; return 0;
    }
    case R_VSD: {                                                                          // VariableScaleDecimal.Do
        if (v.t != JV::OBJ) return DBZ_HOST;
        const JV* val = get(v, "value"); if (!val) return DBZ_UNPARSED; if (val->t != JV::STR) return DBZ_HOST;
        int64_t scale = 0; const JV* sc = get(v, "scale");
        if (sc) { if (sc->t != JV::NUM) return DBZ_HOST; if (jsn::go_parse_int(sc->s, 10, 64, scale)) return DBZ_UNPARSED; }
        std::string num; const int rc = base64_to_numeric(val->s, (int)scale, num); if (rc == 2) return DBZ_HOST; if (rc) return DBZ_UNPARSED;
        double f; if (jsn::go_parse_float(num, f)) return DBZ_HOST; c.f = f; return 0;      // json.Number in a double column: carried as the nearest float64
    }
    }
    return DBZ_HOST;
}

struct Plan { std::vector<Field> after, before; std::string schema_text; bool use_sr = false; uint32_t schema_id = 0; std::string table_schema, table_name; bool check_table = false; };
struct Msg { uint64_t end; };
struct Row { int err_col = 0; int kind = 0; uint32_t tx_id = 0; uint64_t lsn = 0, commit_time = 0; std::vector<Cell> cells; };

// strconv.ParseUint of a JSON number literal the way encoding/json fills an unsigned struct field
inline bool lit_uint(const JV& v, int bits, uint64_t& out) { out = 0; if (v.t == JV::NUL) return true; if (v.t != JV::NUM) return false; uint64_t n; if (jsn::go_parse_uint(v.s, 10, 64, n)) return false; if (bits < 64 && (n >> bits)) return false; out = n; return true; }

// Receive() of one event text (schema + payload already separated). rc as above; `row` filled on 0
inline int receive_event(const Plan& pl, sv schema, bool have_payload, const JV& payload, Row& row) {
    if (!have_payload) return DBZ_UNPARSED;                          // UnmarshalPayload(nil): EOF
    bool bad = false; std::string op; JV source; source.t = JV::OBJ; const JV* after = nullptr; const JV* before = nullptr; uint64_t dummy;
    if (payload.t == JV::OBJ) {
        if (has_folded_only(payload, {"after", "before", "op", "source", "transaction", "ts_ms"})) return DBZ_HOST;
        for (auto& kv : payload.kv) {                                 // decode in key order: later duplicates overwrite / merge
            const JV& v = kv.second;
            if (kv.first == "op") { if (v.t == JV::STR) op = v.s; else if (v.t != JV::NUL) bad = true; }
            else if (kv.first == "after") { if (v.t == JV::OBJ) { if (after && after->t == JV::OBJ) return DBZ_HOST; after = &v; } else if (v.t == JV::NUL) after = nullptr; else bad = true; }
            else if (kv.first == "before") { if (v.t == JV::OBJ) { if (before && before->t == JV::OBJ) return DBZ_HOST; before = &v; } else if (v.t == JV::NUL) before = nullptr; else bad = true; }
            else if (kv.first == "ts_ms") { if (!lit_uint(v, 64, dummy)) bad = true; }
            else if (kv.first == "source") {
                if (v.t == JV::OBJ) {
                    if (has_folded_only(v, {"connector", "db", "lsn", "name", "schema", "sequence", "snapshot", "table", "ts_ms", "txId", "version", "xmin"})) return DBZ_HOST;
                    for (auto& f : v.kv) {
                        const JV& x = f.second; const std::string& k = f.first;
                        if (k == "lsn") { uint64_t n; if (!lit_uint(x, 64, n)) bad = true; else if (x.t == JV::NUM) row.lsn = n; }
                        else if (k == "ts_ms") { uint64_t n; if (!lit_uint(x, 64, n)) bad = true; else if (x.t == JV::NUM) row.commit_time = n; }
                        else if (k == "txId") { uint64_t n; if (!lit_uint(x, 32, n)) bad = true; else if (x.t == JV::NUM) row.tx_id = (uint32_t)n; }
                        else if (k == "xmin") { if (x.t == JV::NUM) { int64_t n; if (jsn::go_parse_int(x.s, 10, 64, n)) bad = true; } else if (x.t != JV::NUL) bad = true; }
                        else if (k == "connector" || k == "db" || k == "name" || k == "schema" || k == "sequence" || k == "snapshot" || k == "table" || k == "version") {
                            if (x.t == JV::STR) { if (k == "schema") source.kv.emplace_back("schema", x); if (k == "table") source.kv.emplace_back("table", x); } else if (x.t != JV::NUL) bad = true;
                        }
                    }
                } else if (v.t != JV::NUL) bad = true;
            }
        }
    } else if (payload.t != JV::NUL) bad = true;
    if (bad) return DBZ_UNPARSED;                                   // UnmarshalTypeError
    if (op == "c" || op == "r") row.kind = TF_KIND_INSERT; else if (op == "u") row.kind = TF_KIND_UPDATE; else if (op == "d") row.kind = TF_KIND_DELETE; else return DBZ_UNPARSED;
    row.commit_time *= 1000000;                                      // CommitTime = source.ts_ms * 1e6 (receiver.go:186)
    if (!pl.use_sr && schema != pl.schema_text) return DBZ_OTHER_SCHEMA;
    if (pl.check_table) { const JV* s = get(source, "schema"); const JV* t = get(source, "table"); if ((s ? s->s : "") != pl.table_schema || (t ? t->s : "") != pl.table_name) return DBZ_OTHER_TABLE; }
    const std::vector<Field>& fs = row.kind == TF_KIND_DELETE ? pl.before : pl.after;
    const JV* vals = row.kind == TF_KIND_DELETE ? before : after;
    row.cells.assign(fs.size(), Cell());
    for (size_t i = 0; i < fs.size(); i++) {
        const JV* v = vals ? get(*vals, fs[i].name.c_str()) : nullptr;
        row.err_col = (int)i;
        if (!v) return DBZ_UNPARSED;                                 // "unable to get field %s from 'after'" :214-216
        const int rc = receive(fs[i], *v, row.cells[i]); if (rc) return rc;
    }
    row.err_col = 0;
    return 0;
}

// DoBuf over one message: 0..n events. Calls sink(rc, row) per event.
template <typename F> inline void do_message(const Plan& pl, sv msg, F&& sink) {
    if (msg.empty()) { Row r; sink(DBZ_UNPARSED, r); return; }        // "debezium parser received empty message"
    sv rest = msg;
    while (!rest.empty()) {
        Row row;
        if (pl.use_sr && rest[0] != 0) { sink(DBZ_UNPARSED, row); return; }          // magic byte check: the rest of the buffer is dropped (:38-40 returns nil)
        size_t len = rest.size();
        if (rest[0] == 0) { size_t z = rest.size() > 5 ? rest.find('\0', 5) : sv::npos; if (z != sv::npos) len = z; }
        sv ev = rest.substr(0, len);
        int rc; sv schema; JV payload; bool have_payload = false;
        if (pl.use_sr) {
            if (ev.size() < 5) rc = DBZ_UNPARSED;
            else {
                const uint32_t id = ((uint32_t)(uint8_t)ev[1] << 24) | ((uint32_t)(uint8_t)ev[2] << 16) | ((uint32_t)(uint8_t)ev[3] << 8) | (uint8_t)ev[4];
                GoJson g; g.s = ev.substr(5);
                if (id != pl.schema_id) rc = DBZ_OTHER_SCHEMA;
                else { if (!g.val(payload)) rc = DBZ_UNPARSED; else { if (g.odd_key) rc = DBZ_HOST; else { have_payload = true; rc = receive_event(pl, schema, true, payload, row); } } }   // Decoder: bytes after the first value are ignored
            }
        } else {
            GoJson g; g.s = ev; JV root;
            if (!g.val(root)) rc = DBZ_UNPARSED;
            else { g.ws(); if (g.p != ev.size()) rc = DBZ_UNPARSED;                 // json.Unmarshal: trailing data is a syntax error
                   else if (root.t != JV::OBJ && root.t != JV::NUL) rc = DBZ_UNPARSED;  // UnmarshalTypeError into the struct
                   else if (g.odd_key || has_folded_only(root, {"schema", "payload"})) rc = DBZ_HOST;
                   else { const JV* sc = get(root, "schema"); const JV* pa = get(root, "payload"); if (sc) schema = sc->raw; have_payload = pa != nullptr; rc = receive_event(pl, schema, have_payload, pa ? *pa : payload, row); } }
        }
        sink(rc, row);
        if (rc) return;                                               // DoOne returns no leftover after an unparsed event (:55-66)
        rest = rest.substr(len);
    }
}

}  // namespace dbz
