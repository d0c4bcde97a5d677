// Minimal JSON reader/writer for plan configuration (schema, transformers, sink options).
// Config-time only — never on the per-batch path.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace tfj {

struct Value;
using ValuePtr = std::shared_ptr<Value>;

struct Value {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string num_text;
    std::string str;
    std::vector<ValuePtr> arr;
    std::vector<std::pair<std::string, ValuePtr>> obj;   // insertion order kept

    const Value* get(const std::string& k) const {
        if (kind != Obj) return nullptr;
        for (auto& kv : obj) if (kv.first == k) return kv.second.get();
        return nullptr;
    }
    std::string get_str(const std::string& k, const std::string& def = "") const {
        const Value* v = get(k); return (v && v->kind == Str) ? v->str : def;
    }
    bool get_bool(const std::string& k, bool def = false) const {
        const Value* v = get(k); return (v && v->kind == Bool) ? v->b : def;
    }
    double get_num(const std::string& k, double def = 0) const {
        const Value* v = get(k); return (v && v->kind == Num) ? v->num : def;
    }
    std::vector<std::string> get_str_list(const std::string& k) const {
        std::vector<std::string> out; const Value* v = get(k);
        if (v && v->kind == Arr) for (auto& e : v->arr) if (e->kind == Str) out.push_back(e->str);
        return out;
    }
};

class Parser {
public:
    explicit Parser(const std::string& s) : s_(s) {}
    ValuePtr parse() { ws(); ValuePtr v = value(); ws(); if (p_ != s_.size()) fail("trailing characters"); return v; }
private:
    const std::string& s_; size_t p_ = 0;
    [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json: ") + m + " at " + std::to_string(p_)); }
    void ws() { while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) p_++; }
    ValuePtr value() {
        if (p_ >= s_.size()) fail("unexpected end");
        char c = s_[p_];
        auto v = std::make_shared<Value>();
        if (c == '{') {
            v->kind = Value::Obj; p_++; ws();
            if (p_ < s_.size() && s_[p_] == '}') { p_++; return v; }
            for (;;) {
                ws(); if (p_ >= s_.size() || s_[p_] != '"') fail("expected key");
                std::string k = string(); ws();
                if (p_ >= s_.size() || s_[p_] != ':') fail("expected ':'");
                p_++; ws(); v->obj.emplace_back(k, value()); ws();
                if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
                if (p_ < s_.size() && s_[p_] == '}') { p_++; break; }
                fail("expected ',' or '}'");
            }
            return v;
        }
        if (c == '[') {
            v->kind = Value::Arr; p_++; ws();
            if (p_ < s_.size() && s_[p_] == ']') { p_++; return v; }
            for (;;) {
                ws(); v->arr.push_back(value()); ws();
                if (p_ < s_.size() && s_[p_] == ',') { p_++; continue; }
                if (p_ < s_.size() && s_[p_] == ']') { p_++; break; }
                fail("expected ',' or ']'");
            }
            return v;
        }
        if (c == '"') { v->kind = Value::Str; v->str = string(); return v; }
        if (s_.compare(p_, 4, "true") == 0) { v->kind = Value::Bool; v->b = true; p_ += 4; return v; }
        if (s_.compare(p_, 5, "false") == 0) { v->kind = Value::Bool; v->b = false; p_ += 5; return v; }
        if (s_.compare(p_, 4, "null") == 0) { p_ += 4; return v; }
        size_t st = p_;
        while (p_ < s_.size() && (isdigit((unsigned char)s_[p_]) || s_[p_] == '-' || s_[p_] == '+' || s_[p_] == '.' || s_[p_] == 'e' || s_[p_] == 'E')) p_++;
        if (st == p_) fail("unexpected character");
        v->kind = Value::Num; v->num_text = s_.substr(st, p_ - st); v->num = strtod(v->num_text.c_str(), nullptr);
        return v;
    }
    static void put_utf8(std::string& o, uint32_t cp) {
        if (cp < 0x80) o += (char)cp;
        else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
        else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    }
    std::string string() {
        std::string o; p_++;
        while (p_ < s_.size() && s_[p_] != '"') {
            char c = s_[p_++];
            if (c != '\\') { o += c; continue; }
            if (p_ >= s_.size()) fail("bad escape");
            char e = s_[p_++];
            switch (e) {
            case 'n': o += '\n'; break; case 't': o += '\t'; break; case 'r': o += '\r'; break;
            case 'b': o += '\b'; break; case 'f': o += '\f'; break;
            case 'u': {
                if (p_ + 4 > s_.size()) fail("bad \\u");
                uint32_t cp = (uint32_t)strtoul(s_.substr(p_, 4).c_str(), nullptr, 16); p_ += 4;
                if (cp >= 0xD800 && cp < 0xDC00 && p_ + 6 <= s_.size() && s_[p_] == '\\' && s_[p_ + 1] == 'u') {
                    uint32_t lo = (uint32_t)strtoul(s_.substr(p_ + 2, 4).c_str(), nullptr, 16);
                    if (lo >= 0xDC00 && lo < 0xE000) { cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); p_ += 6; }
                }
                put_utf8(o, cp); break;
            }
            default: o += e;
            }
        }
        if (p_ >= s_.size()) fail("unterminated string");
        p_++;
        return o;
    }
};

inline ValuePtr parse(const std::string& s) { Parser p(s); return p.parse(); }

inline std::string quote(const std::string& s) {
    static const char* hex = "0123456789abcdef";
    std::string o = "\"";
    for (unsigned char c : s) {
        if (c == '"' || c == '\\') { o += '\\'; o += (char)c; }
        else if (c == '\n') o += "\\n";
        else if (c == '\t') o += "\\t";
        else if (c == '\r') o += "\\r";
        else if (c < 0x20) { o += "\\u00"; o += hex[c >> 4]; o += hex[c & 15]; }
        else o += (char)c;
    }
    return o + "\"";
}

}  // namespace tfj
