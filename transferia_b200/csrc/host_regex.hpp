// Go's regexp (RE2 syntax, leftmost-first, rune-wise over UTF-8) for the regex_replace_transformer step of tfgpu_sink_push:
//   regexp.Compile(cfg.RegexMatch) + Regexp.ReplaceAllString / ReplaceAll   pkg/transformer/registry/regex_replace/transformer.go:19,127-142
// Go's standard library is not part of /root/reference; this restates its published design: the parser of regexp/syntax (Perl flags), the
// simplifier's expansion of x{n,m}, the compiler's instruction layout (alternation / quest / star / plus with the preferred branch first, a
// star over a nullable operand compiled as (x+)?, golang.org/issue/46123), the Pike machine of regexp/exec.go (threads in priority order, a
// sparse set per step, the first match cuts the lower-priority threads), the ReplaceAll loop of regexp.go (an empty match right behind a
// match is not replaced; always advance one rune) and Regexp.Expand's template rules ($1, ${1}, $name, $$, the longest name wins, a
// malformed $ stays as text).
// Not taken (compile() throws Unsupported and the step stays with the Go transformer): \pL / \PL Unicode class tables, case folding
// (?i) over runes outside ASCII (ASCII letters fold, with the two runes that fold into them: K U+212A and s U+017F), programs above
// MAX_INST instructions, a non-ASCII rune inside a $name of the replace rule.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace tfre {

struct SyntaxError : std::runtime_error { using std::runtime_error::runtime_error; };     // regexp.Compile fails on it as well
struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };     // valid Go syntax this engine does not take

constexpr int32_t RUNE_MAX = 0x10FFFF, RUNE_ERROR = 0xFFFD, END_OF_TEXT = -1;
constexpr size_t MAX_INST = 10000;

// utf8.DecodeRune: an invalid or truncated sequence is (RuneError, 1)
inline int32_t decode_rune(const uint8_t* p, size_t n, int& width) {
    if (!n) { width = 0; return END_OF_TEXT; }
    const uint8_t b0 = p[0];
    width = 1;
    if (b0 < 0x80) return b0;
    if (b0 < 0xC2 || b0 > 0xF4) return RUNE_ERROR;
    if (b0 < 0xE0) { if (n < 2 || (p[1] & 0xC0) != 0x80) return RUNE_ERROR; width = 2; return ((b0 & 0x1F) << 6) | (p[1] & 0x3F); }
    if (b0 < 0xF0) {
        if (n < 3 || (p[1] & 0xC0) != 0x80 || (p[2] & 0xC0) != 0x80) return RUNE_ERROR;
        if ((b0 == 0xE0 && p[1] < 0xA0) || (b0 == 0xED && p[1] > 0x9F)) return RUNE_ERROR;          // overlong, surrogates
        width = 3; return ((b0 & 0x0F) << 12) | ((p[1] & 0x3F) << 6) | (p[2] & 0x3F);
    }
    if (n < 4 || (p[1] & 0xC0) != 0x80 || (p[2] & 0xC0) != 0x80 || (p[3] & 0xC0) != 0x80) return RUNE_ERROR;
    if ((b0 == 0xF0 && p[1] < 0x90) || (b0 == 0xF4 && p[1] > 0x8F)) return RUNE_ERROR;
    width = 4; return ((b0 & 0x07) << 18) | ((p[1] & 0x3F) << 12) | ((p[2] & 0x3F) << 6) | (p[3] & 0x3F);
}

using Ranges = std::vector<std::pair<int32_t, int32_t>>;          // sorted, merged, inclusive

inline void normalise(Ranges& r) {
    std::sort(r.begin(), r.end());
    Ranges out;
    for (auto& x : r) { if (!out.empty() && x.first <= out.back().second + 1) out.back().second = std::max(out.back().second, x.second); else out.push_back(x); }
    r.swap(out);
}
inline Ranges negate(Ranges r) {
    normalise(r); Ranges out; int32_t next = 0;
    for (auto& x : r) { if (x.first > next) out.push_back({next, x.first - 1}); next = x.second + 1; }
    if (next <= RUNE_MAX) out.push_back({next, RUNE_MAX});
    return out;
}

// ---------------------------------------------------------------- syntax tree (regexp/syntax parse.go, Perl flags: ClassNL | OneLine | PerlX)
enum Op { NO_MATCH, EMPTY, LIT, CLASS, BEGIN_TEXT, END_TEXT, BEGIN_LINE, END_LINE, WORD_B, NO_WORD_B, CAT, ALT, STAR, PLUS, QUEST, REPEAT, CAPTURE };
struct Node {
    Op op = EMPTY; bool lazy = false; int32_t rune = 0; Ranges cls; int min = 0, max = 0, cap = 0;
    std::vector<std::shared_ptr<Node>> sub;                       // shared: the simplifier repeats one operand (x{2,5} = xx(x(x(x)?)?)?)
};
using NodeP = std::shared_ptr<Node>;
inline NodeP mk(Op op) { NodeP n(new Node); n->op = op; return n; }

struct Parser {
    const uint8_t* s; size_t n, at = 0; int ncap = 0; std::vector<std::string> names{""};
    bool dot_nl = false, one_line = true, fold = false, ungreedy = false; int depth = 0;
    explicit Parser(const std::string& p) : s((const uint8_t*)p.data()), n(p.size()) {}

    bool more() const { return at < n; }
    int32_t next() {
        int w; const int32_t r = decode_rune(s + at, n - at, w);
        if (r == RUNE_ERROR && w == 1) throw SyntaxError("invalid UTF-8");      // checkUTF8
        at += w; return r;
    }
    bool eat(char c) { if (at < n && s[at] == (uint8_t)c) { at++; return true; } return false; }
    bool looking(const char* lit) const { const size_t l = std::strlen(lit); return n - at >= l && !std::memcmp(s + at, lit, l); }

    // Case folding (?i): unicode.SimpleFold orbits. Carried for ASCII only — a letter and its other case, and the two runes outside ASCII that
    // fold INTO it: U+212A KELVIN SIGN (k, K) and U+017F LATIN SMALL LETTER LONG S (s, S). A rune >= 0x80 under (?i) would need the tables.
    static void add_folds(Ranges& r) {
        Ranges extra;
        for (auto& x : r) {
            if (x.second >= 0x80) throw Unsupported("case folding (?i) over runes outside ASCII (the Unicode folding tables are not carried)");
            for (int32_t c = x.first; c <= x.second; c++) {
                if (c >= 'a' && c <= 'z') extra.push_back({c - 32, c - 32});
                else if (c >= 'A' && c <= 'Z') extra.push_back({c + 32, c + 32});
                if (c == 'k' || c == 'K') extra.push_back({0x212A, 0x212A});
                if (c == 's' || c == 'S') extra.push_back({0x017F, 0x017F});
            }
        }
        r.insert(r.end(), extra.begin(), extra.end());
    }
    Ranges perl(char c) const {                               // \d \w \s and their negations under the current flags: fold first, negate after (appendGroup)
        Ranges r = perl_class((char)(c | 0x20));
        if (fold) { add_folds(r); normalise(r); }
        return (c >= 'A' && c <= 'Z') ? negate(r) : r;
    }
    static Ranges perl_class(char c) {
        Ranges r;
        switch (c | 0x20) {
        case 'd': r = {{'0', '9'}}; break;
        case 's': r = {{'\t', '\n'}, {'\f', '\r'}, {' ', ' '}}; break;           // no \v (perl_groups)
        default: r = {{'0', '9'}, {'A', 'Z'}, {'_', '_'}, {'a', 'z'}}; break;
        }
        return (c >= 'A' && c <= 'Z') ? negate(r) : r;
    }
    static bool posix_class(const std::string& name, Ranges& r) {
        if (name == "alnum") r = {{'0', '9'}, {'A', 'Z'}, {'a', 'z'}};
        else if (name == "alpha") r = {{'A', 'Z'}, {'a', 'z'}};
        else if (name == "ascii") r = {{0, 0x7F}};
        else if (name == "blank") r = {{'\t', '\t'}, {' ', ' '}};
        else if (name == "cntrl") r = {{0, 0x1F}, {0x7F, 0x7F}};
        else if (name == "digit") r = {{'0', '9'}};
        else if (name == "graph") r = {{'!', '~'}};
        else if (name == "lower") r = {{'a', 'z'}};
        else if (name == "print") r = {{' ', '~'}};
        else if (name == "punct") r = {{'!', '/'}, {':', '@'}, {'[', '`'}, {'{', '~'}};
        else if (name == "space") r = {{'\t', '\r'}, {' ', ' '}};
        else if (name == "upper") r = {{'A', 'Z'}};
        else if (name == "word") r = {{'0', '9'}, {'A', 'Z'}, {'a', 'z'}, {'_', '_'}};
        else if (name == "xdigit") r = {{'0', '9'}, {'A', 'F'}, {'a', 'f'}};
        else return false;
        return true;
    }

    // parseEscape: the rune an escape stands for (the backslash already taken)
    int32_t escape_rune() {
        if (!more()) throw SyntaxError("trailing backslash at end of expression");
        const int32_t c = next();
        auto hex = [](int32_t ch) { return ch >= '0' && ch <= '9' ? ch - '0' : ch >= 'a' && ch <= 'f' ? ch - 'a' + 10 : ch >= 'A' && ch <= 'F' ? ch - 'A' + 10 : -1; };
        switch (c) {
        case '1': case '2': case '3': case '4': case '5': case '6': case '7':
            if (!more() || s[at] < '0' || s[at] > '7') throw SyntaxError("invalid escape sequence");     // a single digit is a backreference
            [[fallthrough]];
        case '0': {
            int32_t r = c - '0';
            for (int i = 1; i < 3 && more() && s[at] >= '0' && s[at] <= '7'; i++) r = r * 8 + (s[at++] - '0');
            return r;
        }
        case 'x': {
            if (!more()) throw SyntaxError("invalid escape sequence");
            if (eat('{')) {
                int32_t r = 0; int nd = 0;
                while (true) {
                    if (!more()) throw SyntaxError("invalid escape sequence");
                    const int32_t ch = next();
                    if (ch == '}') break;
                    const int v = hex(ch); if (v < 0) throw SyntaxError("invalid escape sequence");
                    r = r * 16 + v; if (r > RUNE_MAX) throw SyntaxError("invalid escape sequence"); nd++;
                }
                if (!nd) throw SyntaxError("invalid escape sequence");
                return r;
            }
            const int a = hex(next()); if (a < 0 || !more()) throw SyntaxError("invalid escape sequence");
            const int b = hex(next()); if (b < 0) throw SyntaxError("invalid escape sequence");
            return a * 16 + b;
        }
        case 'a': return 7; case 'f': return '\f'; case 'n': return '\n'; case 'r': return '\r'; case 't': return '\t'; case 'v': return '\v';
        default:
            if (c < 0x80 && !((c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'))) return c;        // punctuation stands for itself
            throw SyntaxError("invalid escape sequence");
        }
    }

    int32_t class_char() {
        if (!more()) throw SyntaxError("missing closing ]");
        if (s[at] == '\\') { at++; return escape_rune(); }
        return next();
    }
    NodeP parse_class() {                                    // after '['; parseClass with PerlX (a '-' is fine anywhere) and ClassNL ([^a] takes \n)
        Ranges r; const bool neg = eat('^'); bool first = true;
        while (!more() || s[at] != ']' || first) {
            first = false;
            if (n - at > 2 && s[at] == '[' && s[at + 1] == ':') {
                const std::string rest((const char*)s + at, n - at); const size_t close = rest.find(":]");
                if (close != std::string::npos) {
                    std::string name = rest.substr(2, close - 2); bool pneg = false;
                    if (!name.empty() && name[0] == '^') { pneg = true; name.erase(0, 1); }
                    Ranges pr; if (!posix_class(name, pr)) throw SyntaxError("invalid character class range");
                    if (fold) add_folds(pr);
                    if (pneg) pr = negate(pr);
                    r.insert(r.end(), pr.begin(), pr.end()); at += close + 2; continue;
                }
            }
            if (n - at >= 2 && s[at] == '\\' && (s[at + 1] == 'p' || s[at + 1] == 'P')) throw Unsupported("Unicode class tables (\\p) are not carried");
            if (n - at >= 2 && s[at] == '\\' && std::strchr("dDsSwW", s[at + 1])) { const Ranges pr = perl((char)s[at + 1]); r.insert(r.end(), pr.begin(), pr.end()); at += 2; continue; }
            const int32_t lo = class_char(); int32_t hi = lo;
            if (n - at >= 2 && s[at] == '-' && s[at + 1] != ']') { at++; hi = class_char(); if (hi < lo) throw SyntaxError("invalid character class range"); }
            if (fold) { Ranges one{{lo, hi}}; add_folds(one); r.insert(r.end(), one.begin(), one.end()); }      // appendFoldedRange
            else r.push_back({lo, hi});
        }
        at++;                                                 // ']'
        normalise(r);
        NodeP node = mk(CLASS); node->cls = neg ? negate(r) : r;
        return node;
    }

    bool parse_int(int& v) {                                  // parseInt: no leading zeros, >= 1e8 reads as -1
        if (!more() || s[at] < '0' || s[at] > '9') return false;
        if (n - at >= 2 && s[at] == '0' && s[at + 1] >= '0' && s[at + 1] <= '9') return false;
        int64_t x = 0; bool big = false;
        while (more() && s[at] >= '0' && s[at] <= '9') { if (x >= 100000000) big = true; else x = x * 10 + (s[at] - '0'); at++; }
        v = big ? -1 : (int)x; return true;
    }
    bool parse_repeat(int& mn, int& mx) {                     // at '{'; leaves `at` untouched when it is not a repeat
        const size_t save = at; at++;
        auto fail = [&] { at = save; return false; };
        if (!parse_int(mn)) return fail();
        if (!more()) return fail();
        if (s[at] != ',') mx = mn;
        else {
            at++;
            if (!more()) return fail();
            if (s[at] == '}') mx = -1;
            else { if (!parse_int(mx)) return fail(); if (mx < 0) mn = -1; }
        }
        if (!more() || s[at] != '}') return fail();
        at++; return true;
    }
    static bool repeat_is_valid(const Node& re, int lim) {
        if (re.op == REPEAT) {
            int m = re.max;
            if (m == 0) return true;
            if (m < 0) m = re.min;
            if (m > lim) return false;
            if (m > 0) lim /= m;
        }
        for (auto& sub : re.sub) if (!repeat_is_valid(*sub, lim)) return false;
        return true;
    }

    NodeP parse_alt() {
        if (++depth > 1000) throw SyntaxError("expression nests too deeply");
        std::vector<NodeP> alts; alts.push_back(parse_cat());
        while (more() && s[at] == '|') { at++; alts.push_back(parse_cat()); }
        depth--;
        if (alts.size() == 1) return alts[0];
        NodeP a = mk(ALT); a->sub = std::move(alts); return a;
    }

    NodeP parse_cat() {
        std::vector<NodeP> items; bool last_repeat = false;
        auto lit = [&](int32_t c) {
            if (fold) {                                       // OpLiteral with FoldCase: the rune's SimpleFold orbit
                Ranges one{{c, c}}; add_folds(one); normalise(one);
                if (one.size() > 1 || one[0].first != one[0].second) { NodeP k = mk(CLASS); k->cls = one; items.push_back(k); return; }
            }
            NodeP l = mk(LIT); l->rune = c; items.push_back(l);
        };
        while (more() && s[at] != '|' && s[at] != ')') {
            bool is_repeat = false;
            const uint8_t c = s[at];
            if (c == '(') {
                const bool sv_dot = dot_nl, sv_one = one_line, sv_fold = fold, sv_ung = ungreedy;
                if (n - at >= 2 && s[at + 1] == '?') {
                    if (looking("(?P<") || (looking("(?<") && !looking("(?<=") && !looking("(?<!"))) {
                        at += looking("(?P<") ? 4 : 3;
                        const std::string rest((const char*)s + at, n - at); const size_t end = rest.find('>');
                        if (end == std::string::npos) throw SyntaxError("invalid named capture");
                        const std::string name = rest.substr(0, end);
                        if (name.empty()) throw SyntaxError("invalid named capture");
                        for (char ch : name) if (!(ch == '_' || (ch >= '0' && ch <= '9') || (ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z'))) throw SyntaxError("invalid named capture");
                        // (a repeated name is legal in Go: Expand takes the first group of that name that took part in the match)
                        at += end + 1;
                        NodeP cp = mk(CAPTURE); cp->cap = ++ncap; names.push_back(name);
                        cp->sub.push_back(parse_alt());
                        if (!eat(')')) throw SyntaxError("missing closing )");
                        dot_nl = sv_dot; one_line = sv_one; fold = sv_fold; ungreedy = sv_ung; items.push_back(cp);
                    } else {
                        at += 2;                              // parsePerlFlags
                        bool neg = false, saw = false, group = false, done = false; bool f_dot = dot_nl, f_one = one_line, f_fold = fold, f_ung = ungreedy;
                        while (more() && !done) {
                            const int32_t f = next();
                            switch (f) {
                            case 'i': f_fold = !neg; saw = true; break;
                            case 'U': f_ung = !neg; saw = true; break;
                            case 'm': f_one = neg; saw = true; break;
                            case 's': f_dot = !neg; saw = true; break;
                            case '-': if (neg) throw SyntaxError("missing closing )"); neg = true; saw = false; break;
                            case ':': case ')':
                                if (neg && !saw) throw SyntaxError("missing closing )");
                                group = f == ':'; done = true; break;
                            default: throw SyntaxError("missing closing )");
                            }
                        }
                        if (!done) throw SyntaxError("missing closing )");
                        dot_nl = f_dot; one_line = f_one; fold = f_fold; ungreedy = f_ung;
                        if (group) {
                            NodeP g = parse_alt();
                            if (!eat(')')) throw SyntaxError("missing closing )");
                            dot_nl = sv_dot; one_line = sv_one; fold = sv_fold; ungreedy = sv_ung; items.push_back(g);
                        } else { last_repeat = false; continue; }      // flags stay until the enclosing group closes; not an operand
                    }
                } else {
                    at++;
                    NodeP cp = mk(CAPTURE); cp->cap = ++ncap; names.push_back("");
                    cp->sub.push_back(parse_alt());
                    if (!eat(')')) throw SyntaxError("missing closing )");
                    dot_nl = sv_dot; one_line = sv_one; fold = sv_fold; ungreedy = sv_ung; items.push_back(cp);
                }
            }
            else if (c == '^') { at++; items.push_back(mk(one_line ? BEGIN_TEXT : BEGIN_LINE)); }
            else if (c == '$') { at++; items.push_back(mk(one_line ? END_TEXT : END_LINE)); }
            else if (c == '.') { at++; NodeP d = mk(CLASS); if (dot_nl) d->cls = {{0, RUNE_MAX}}; else d->cls = {{0, '\n' - 1}, {'\n' + 1, RUNE_MAX}}; items.push_back(d); }
            else if (c == '[') { at++; items.push_back(parse_class()); }
            else if (c == '*' || c == '+' || c == '?' || c == '{') {
                int mn = 0, mx = 0; Op op = c == '*' ? STAR : c == '+' ? PLUS : c == '?' ? QUEST : REPEAT;
                if (c == '{') {
                    if (!parse_repeat(mn, mx)) { at++; lit('{'); last_repeat = false; continue; }
                    if (mn < 0 || mn > 1000 || mx > 1000 || (mx >= 0 && mn > mx)) throw SyntaxError("invalid repeat count");
                } else at++;
                const bool lazy = eat('?') != ungreedy;                  // (?U) swaps the meaning of x* and x*? (flags ^= NonGreedy)
                if (last_repeat) throw SyntaxError("invalid nested repetition operator");
                if (items.empty()) throw SyntaxError("missing argument to repetition operator");
                NodeP r = mk(op); r->lazy = lazy; r->min = mn; r->max = mx; r->sub.push_back(items.back());
                if (op == REPEAT && (mn >= 2 || mx >= 2) && !repeat_is_valid(*r, 1000)) throw SyntaxError("invalid repeat count");
                items.back() = r; is_repeat = true;
            }
            else if (c == '\\') {
                if (n - at >= 2) {
                    const uint8_t e = s[at + 1];
                    if (e == 'A') { at += 2; items.push_back(mk(BEGIN_TEXT)); last_repeat = false; continue; }
                    if (e == 'z') { at += 2; items.push_back(mk(END_TEXT)); last_repeat = false; continue; }
                    if (e == 'b') { at += 2; items.push_back(mk(WORD_B)); last_repeat = false; continue; }
                    if (e == 'B') { at += 2; items.push_back(mk(NO_WORD_B)); last_repeat = false; continue; }
                    if (e == 'C') throw SyntaxError("invalid escape sequence");
                    if (e == 'Q') {
                        at += 2;
                        const std::string rest((const char*)s + at, n - at); const size_t end = rest.find("\\E");
                        const size_t stop = at + (end == std::string::npos ? rest.size() : end);
                        while (at < stop) lit(next());
                        if (end != std::string::npos) at += 2;
                        last_repeat = false; continue;
                    }
                    if (e == 'p' || e == 'P') throw Unsupported("Unicode class tables (\\p) are not carried");
                    if (std::strchr("dDsSwW", e)) { at += 2; NodeP k = mk(CLASS); k->cls = perl((char)e); normalise(k->cls); items.push_back(k); last_repeat = false; continue; }
                }
                at++; lit(escape_rune());
            }
            else lit(next());
            last_repeat = is_repeat;
        }
        if (items.empty()) return mk(EMPTY);
        if (items.size() == 1) return items[0];
        NodeP cat = mk(CAT); cat->sub = std::move(items); return cat;
    }

    NodeP parse() {
        NodeP root = parse_alt();
        if (more()) throw SyntaxError(s[at] == ')' ? "unexpected )" : "unexpected character");
        return root;
    }
};

// ---------------------------------------------------------------- simplify (regexp/syntax simplify.go: only what changes the program's shape)
inline NodeP simplify1(Op op, bool lazy, const NodeP& sub) {
    if (sub->op == EMPTY) return sub;
    if (op == sub->op && lazy == sub->lazy) return sub;
    NodeP r = mk(op); r->lazy = lazy; r->sub.push_back(sub); return r;
}
inline NodeP simplify(const NodeP& re) {
    switch (re->op) {
    case CAPTURE: case CAT: case ALT: {
        NodeP out = mk(re->op); out->cap = re->cap;
        for (auto& s : re->sub) out->sub.push_back(simplify(s));
        return out;
    }
    case STAR: case PLUS: case QUEST: return simplify1(re->op, re->lazy, simplify(re->sub[0]));
    case REPEAT: {
        if (re->min == 0 && re->max == 0) return mk(EMPTY);
        const NodeP sub = simplify(re->sub[0]);
        if (re->max == -1) {
            if (re->min == 0) return simplify1(STAR, re->lazy, sub);
            if (re->min == 1) return simplify1(PLUS, re->lazy, sub);
            NodeP cat = mk(CAT);
            for (int i = 0; i < re->min - 1; i++) cat->sub.push_back(sub);
            cat->sub.push_back(simplify1(PLUS, re->lazy, sub));
            return cat;
        }
        if (re->min == 1 && re->max == 1) return sub;
        NodeP prefix;
        if (re->min > 0) { prefix = mk(CAT); for (int i = 0; i < re->min; i++) prefix->sub.push_back(sub); }
        if (re->max > re->min) {
            NodeP suffix = simplify1(QUEST, re->lazy, sub);
            for (int i = re->min + 1; i < re->max; i++) { NodeP two = mk(CAT); two->sub = {sub, suffix}; suffix = simplify1(QUEST, re->lazy, two); }
            if (!prefix) return suffix;
            prefix->sub.push_back(suffix);
        }
        if (prefix) return prefix;
        return mk(NO_MATCH);
    }
    default: return re;
    }
}

// ---------------------------------------------------------------- program (regexp/syntax compile.go)
enum InstOp : uint8_t { I_FAIL, I_ALT, I_CAP, I_EMPTY, I_MATCH, I_NOP, I_RUNE1, I_CLASS };
enum : uint8_t { E_BEGIN_LINE = 1, E_END_LINE = 2, E_BEGIN_TEXT = 4, E_END_TEXT = 8, E_WORD = 16, E_NO_WORD = 32 };
struct Inst { InstOp op = I_FAIL; uint32_t out = 0, arg = 0; int32_t rune = 0; uint32_t cls = 0; };

struct Prog {
    std::vector<Inst> inst; std::vector<Ranges> classes; uint32_t start = 0; int ncap = 0; std::vector<std::string> names;
    bool anchored = false;            // Prog.StartCond() has EmptyBeginText: a match can only begin at offset 0 (exec.go stops looking elsewhere)
    int32_t first_rune = -1;          // Prog.Prefix(), its first rune: every match begins with this literal (the search jumps to its next occurrence)
    uint8_t first_byte = 0;
    std::vector<std::array<uint64_t, 2>> ascii;   // per class: membership of the runes below 0x80
    bool has_empty = false;           // some assertion instruction: only then the context conditions are worth computing
    bool first_any = true; bool first[256] = {};   // the bytes a match can begin with (a superset; first_any: no restriction, e.g. the empty match)
};

struct Compiler {
    Prog p;
    // a fragment: entry instruction, the list of dangling exits (instruction << 1 | which: 0 = out, 1 = arg), nullable
    struct Frag { uint32_t i = 0; std::vector<uint32_t> out; bool nullable = false; };
    Frag inst(InstOp op) {
        if (p.inst.size() >= MAX_INST) throw Unsupported("the expression compiles to more than MAX_INST instructions");
        Inst in; in.op = op; p.inst.push_back(in);
        Frag f; f.i = (uint32_t)p.inst.size() - 1; f.nullable = true; return f;
    }
    void patch(const std::vector<uint32_t>& l, uint32_t to) { for (uint32_t x : l) { if (x & 1) p.inst[x >> 1].arg = to; else p.inst[x >> 1].out = to; } }
    Frag nop() { Frag f = inst(I_NOP); f.out = {f.i << 1}; return f; }
    Frag fail() { return Frag{}; }
    Frag cap(uint32_t arg) { Frag f = inst(I_CAP); f.out = {f.i << 1}; p.inst[f.i].arg = arg; return f; }
    Frag empty(uint8_t cond) { Frag f = inst(I_EMPTY); p.inst[f.i].arg = cond; f.out = {f.i << 1}; return f; }
    Frag rune(int32_t r) { Frag f = inst(I_RUNE1); f.nullable = false; p.inst[f.i].rune = r; f.out = {f.i << 1}; return f; }
    Frag cls(const Ranges& r) { Frag f = inst(I_CLASS); f.nullable = false; p.inst[f.i].cls = (uint32_t)p.classes.size(); p.classes.push_back(r); f.out = {f.i << 1}; return f; }
    Frag cat(Frag a, Frag b) {
        if (!a.i || !b.i) return Frag{};
        patch(a.out, b.i);
        Frag f; f.i = a.i; f.out = std::move(b.out); f.nullable = a.nullable && b.nullable; return f;
    }
    Frag alt(Frag a, Frag b) {
        if (!a.i) return b;
        if (!b.i) return a;
        Frag f = inst(I_ALT); p.inst[f.i].out = a.i; p.inst[f.i].arg = b.i;
        f.out = std::move(a.out); f.out.insert(f.out.end(), b.out.begin(), b.out.end()); f.nullable = a.nullable || b.nullable; return f;
    }
    Frag quest(Frag a, bool lazy) {
        Frag f = inst(I_ALT);
        if (lazy) { p.inst[f.i].arg = a.i; f.out = {f.i << 1}; } else { p.inst[f.i].out = a.i; f.out = {f.i << 1 | 1}; }
        f.out.insert(f.out.end(), a.out.begin(), a.out.end()); return f;
    }
    Frag loop(Frag a, bool lazy) {
        Frag f = inst(I_ALT);
        if (lazy) { p.inst[f.i].arg = a.i; f.out = {f.i << 1}; } else { p.inst[f.i].out = a.i; f.out = {f.i << 1 | 1}; }
        patch(a.out, f.i); return f;
    }
    Frag plus(Frag a, bool lazy) { const uint32_t entry = a.i; const bool nullable = a.nullable; Frag l = loop(std::move(a), lazy); Frag f; f.i = entry; f.out = std::move(l.out); f.nullable = nullable; return f; }
    Frag star(Frag a, bool lazy) { if (a.nullable) return quest(plus(std::move(a), lazy), lazy); return loop(std::move(a), lazy); }

    Frag compile(const Node& re) {
        switch (re.op) {
        case NO_MATCH: return fail();
        case EMPTY: return nop();
        case LIT: return rune(re.rune);
        case CLASS: return cls(re.cls);
        case BEGIN_TEXT: return empty(E_BEGIN_TEXT); case END_TEXT: return empty(E_END_TEXT);
        case BEGIN_LINE: return empty(E_BEGIN_LINE); case END_LINE: return empty(E_END_LINE);
        case WORD_B: return empty(E_WORD); case NO_WORD_B: return empty(E_NO_WORD);
        case CAPTURE: { Frag bra = cap((uint32_t)re.cap << 1); Frag sub = compile(*re.sub[0]); Frag ket = cap((uint32_t)re.cap << 1 | 1); return cat(cat(std::move(bra), std::move(sub)), std::move(ket)); }
        case STAR: return star(compile(*re.sub[0]), re.lazy);
        case PLUS: return plus(compile(*re.sub[0]), re.lazy);
        case QUEST: return quest(compile(*re.sub[0]), re.lazy);
        case CAT: {
            if (re.sub.empty()) return nop();
            Frag f; bool first = true;
            for (auto& s : re.sub) { if (first) { f = compile(*s); first = false; } else f = cat(std::move(f), compile(*s)); }
            return f;
        }
        case ALT: { Frag f; for (auto& s : re.sub) f = alt(std::move(f), compile(*s)); return f; }
        default: throw std::logic_error("regex: unexpected node");
        }
    }
};

inline Prog compile(const std::string& pattern) {
    Parser ps(pattern);
    const NodeP tree = simplify(ps.parse());
    Compiler c; c.p.inst.emplace_back();                      // instruction 0 is the failure sink (a fragment with i == 0 has failed)
    c.p.ncap = ps.ncap; c.p.names = ps.names;
    Compiler::Frag f = c.compile(*tree);
    Compiler::Frag m = c.inst(I_MATCH);
    if (f.i) { c.patch(f.out, m.i); c.p.start = f.i; } else c.p.start = 0;
    uint8_t cond = 0; uint32_t pc = c.p.start;
    while (pc) {
        const Inst& in = c.p.inst[pc];
        if (in.op == I_EMPTY) cond |= (uint8_t)in.arg; else if (in.op != I_CAP && in.op != I_NOP) break;
        pc = in.out;
    }
    c.p.anchored = cond & E_BEGIN_TEXT;
    pc = c.p.start;
    while (pc && (c.p.inst[pc].op == I_CAP || c.p.inst[pc].op == I_NOP)) pc = c.p.inst[pc].out;
    if (pc && c.p.inst[pc].op == I_RUNE1) {
        const int32_t r = c.p.inst[pc].rune; c.p.first_rune = r;
        c.p.first_byte = r < 0x80 ? (uint8_t)r : r < 0x800 ? (uint8_t)(0xC0 | (r >> 6)) : r < 0x10000 ? (uint8_t)(0xE0 | (r >> 12)) : (uint8_t)(0xF0 | (r >> 18));
        if (r >= 0xD800 && r <= 0xDFFF) c.p.first_rune = -1;                  // no text holds it as a rune
    }
    for (auto& cl : c.p.classes) {
        std::array<uint64_t, 2> bits{0, 0};
        for (auto& r : cl) for (int32_t x = r.first; x <= std::min(r.second, 0x7F); x++) bits[x >> 6] |= 1ull << (x & 63);
        c.p.ascii.push_back(bits);
    }
    for (auto& in : c.p.inst) if (in.op == I_EMPTY) c.p.has_empty = true;
    {   // first-byte set: everything reachable from the start without consuming a rune (assertions taken as passable)
        std::vector<uint8_t> seen(c.p.inst.size(), 0); std::vector<uint32_t> todo{c.p.start}; bool any = c.p.start == 0;
        while (!todo.empty() && !any) {
            const uint32_t at = todo.back(); todo.pop_back();
            if (!at || seen[at]) continue;
            seen[at] = 1; const Inst& in = c.p.inst[at];
            switch (in.op) {
            case I_ALT: todo.push_back(in.out); todo.push_back(in.arg); break;
            case I_CAP: case I_NOP: case I_EMPTY: todo.push_back(in.out); break;
            case I_MATCH: any = true; break;
            case I_RUNE1: case I_CLASS: {
                auto mark = [&](int32_t lo, int32_t hi) {
                    for (int32_t b = lo; b <= std::min(hi, 0x7F); b++) c.p.first[b] = true;
                    if (hi >= 0x80) for (int b = 0x80; b < 0x100; b++) c.p.first[b] = true;      // some multi-byte rune, or U+FFFD standing for an invalid byte
                };
                if (in.op == I_RUNE1) mark(in.rune, in.rune); else for (auto& r : c.p.classes[in.cls]) mark(r.first, r.second);
                break;
            }
            default: break;
            }
        }
        c.p.first_any = any;
    }
    return std::move(c.p);
}

// ---------------------------------------------------------------- Pike machine (regexp/exec.go, first-match mode)
struct Machine {
    const Prog& p; const int nslot;
    struct Queue {
        std::vector<uint32_t> sparse, dense_pc; std::vector<int32_t> slot;    // slot: index into caps (in units of nslot), -1 for a bookkeeping entry
        std::vector<int64_t> caps; uint32_t n = 0, nthreads = 0;              // caps: one row of nslot per thread, sized once for every instruction
        bool has(uint32_t pc) const { const uint32_t j = sparse[pc]; return j < n && dense_pc[j] == pc; }
        void clear() { n = 0; nthreads = 0; }
    } q[2];
    std::vector<int64_t> cur, matchcap; bool matched = false;

    explicit Machine(const Prog& prog) : p(prog), nslot(2 * (prog.ncap + 1)) {
        for (auto& x : q) { x.sparse.assign(p.inst.size(), 0); x.dense_pc.assign(p.inst.size(), 0); x.slot.assign(p.inst.size(), -1); x.caps.assign(p.inst.size() * (size_t)nslot, -1); }
        cur.assign(nslot, -1); matchcap.assign(nslot, -1);
    }
    static bool word(int32_t r) { return (r >= '0' && r <= '9') || (r >= 'A' && r <= 'Z') || (r >= 'a' && r <= 'z') || r == '_'; }
    static uint8_t cond_of(int32_t r0, int32_t r1) {            // syntax.EmptyOpContext
        uint8_t c = 0;
        if (r0 == END_OF_TEXT) c |= E_BEGIN_TEXT | E_BEGIN_LINE; else if (r0 == '\n') c |= E_BEGIN_LINE;
        if (r1 == END_OF_TEXT) c |= E_END_TEXT | E_END_LINE; else if (r1 == '\n') c |= E_END_LINE;
        c |= (word(r0) != word(r1)) ? E_WORD : E_NO_WORD;
        return c;
    }
    void add(Queue& qq, uint32_t pc, int64_t pos, uint8_t cond) {
        while (true) {
            if (pc == 0 || qq.has(pc)) return;
            const uint32_t j = qq.n++; qq.sparse[pc] = j; qq.dense_pc[j] = pc; qq.slot[j] = -1;
            const Inst& in = p.inst[pc];
            switch (in.op) {
            case I_FAIL: return;
            case I_ALT: add(qq, in.out, pos, cond); pc = in.arg; continue;
            case I_EMPTY: if ((cond & in.arg) == in.arg) { pc = in.out; continue; } return;
            case I_NOP: pc = in.out; continue;
            case I_CAP: { const int64_t old = cur[in.arg]; cur[in.arg] = pos; add(qq, in.out, pos, cond); cur[in.arg] = old; return; }
            default:                                           // MATCH, RUNE1, CLASS: a thread
                qq.slot[j] = (int32_t)qq.nthreads; std::copy(cur.begin(), cur.end(), qq.caps.begin() + (size_t)qq.nthreads * nslot); qq.nthreads++; return;
            }
        }
    }
    bool in_class(uint32_t k, int32_t r) const {
        if (r < 0x80) return p.ascii[k][r >> 6] >> (r & 63) & 1;
        const Ranges& rg = p.classes[k];
        size_t lo = 0, hi = rg.size();
        while (lo < hi) { const size_t mid = (lo + hi) / 2; if (r > rg[mid].second) lo = mid + 1; else if (r < rg[mid].first) hi = mid; else return true; }
        return false;
    }
    // the leftmost-first match at or after `pos` (with the text before `pos` as context); slots into `matchcap`
    bool search(const uint8_t* s, size_t n, size_t pos) {
        matched = false; std::fill(matchcap.begin(), matchcap.end(), -1);
        if (!p.start) return false;
        Queue* runq = &q[0]; Queue* nextq = &q[1]; runq->clear(); nextq->clear();
        int32_t r0 = pos == 0 ? END_OF_TEXT : (s[pos - 1] < 0x80 ? (int32_t)s[pos - 1] : RUNE_ERROR);   // only '\n' and ASCII word-ness are ever asked of it
        int w; int32_t r = decode_rune(s + pos, n - pos, w);
        while (true) {
            if (runq->n == 0) {
                if (matched || (p.anchored && pos != 0)) break;
                if (p.first_rune >= 0) {                      // no thread alive: the next match starts at the next occurrence of the leading literal
                    size_t at = pos; bool found = false;
                    while (at < n) {
                        const uint8_t* f = (const uint8_t*)std::memchr(s + at, p.first_byte, n - at);
                        if (!f) break;
                        at = (size_t)(f - s); int fw;
                        if (decode_rune(s + at, n - at, fw) == p.first_rune) { found = true; break; }
                        at++;
                    }
                    if (!found) break;
                    if (at != pos) { pos = at; r0 = s[pos - 1] < 0x80 ? (int32_t)s[pos - 1] : RUNE_ERROR; r = decode_rune(s + pos, n - pos, w); }
                } else if (!p.first_any) {                    // ... or at the next byte some match can begin with
                    size_t at = pos;
                    while (at < n && !p.first[s[at]]) at++;
                    if (at == n) break;
                    if (at != pos) { pos = at; r0 = s[pos - 1] < 0x80 ? (int32_t)s[pos - 1] : RUNE_ERROR; r = decode_rune(s + pos, n - pos, w); }
                }
            }
            const uint8_t cond = p.has_empty ? cond_of(r0, r) : 0;
            if (!matched && (!p.anchored || pos == 0)) { std::fill(cur.begin(), cur.end(), -1); cur[0] = (int64_t)pos; add(*runq, p.start, (int64_t)pos, cond); }
            // step: every thread of runq over the rune r; the followers are added with the context at pos + w
            const size_t npos = pos + w; int w1 = 0; const int32_t r1 = w ? decode_rune(s + npos, n - npos, w1) : END_OF_TEXT;
            const uint8_t ncond = p.has_empty ? cond_of(r, r1) : 0;
            for (uint32_t j = 0; j < runq->n; j++) {
                const int32_t sl = runq->slot[j];
                if (sl < 0) continue;
                const Inst& in = p.inst[runq->dense_pc[j]];
                const int64_t* caps = &runq->caps[(size_t)sl * nslot];
                bool go = false;
                if (in.op == I_MATCH) {
                    std::copy(caps, caps + nslot, matchcap.begin()); matchcap[1] = (int64_t)pos; matched = true;
                    break;                                    // first-match mode: the lower-priority threads are cut off
                }
                else if (in.op == I_RUNE1) go = r == in.rune;
                else go = r >= 0 && in_class(in.cls, r);
                if (go) { std::copy(caps, caps + nslot, cur.begin()); add(*nextq, in.out, (int64_t)npos, ncond); }
            }
            runq->clear();
            if (w == 0) break;
            pos = npos; r0 = r; r = r1; w = w1;
            std::swap(runq, nextq);
        }
        return matched;
    }
};

// Regexp.MatchString: is there a match anywhere in the text
inline bool match_string(const Prog& p, const std::string& text) { Machine m(p); return m.search((const uint8_t*)text.data(), text.size(), 0); }

// ---------------------------------------------------------------- Regexp.Expand templates + ReplaceAll (regexp.go)
struct Template {
    struct Piece { std::string text; bool ref = false; std::vector<int> groups; };   // ref: the first listed group that took part in the match (none listed: nothing)
    std::vector<Piece> pieces;
};
inline Template parse_template(const std::string& t, const Prog& prog) {
    Template out; std::string lit; size_t at = 0;
    auto flush = [&] { if (!lit.empty()) { Template::Piece p; p.text = lit; out.pieces.push_back(p); lit.clear(); } };
    while (at < t.size()) {
        const size_t d = t.find('$', at);
        if (d == std::string::npos) break;
        lit.append(t, at, d - at); at = d + 1;
        if (at < t.size() && t[at] == '$') { lit.push_back('$'); at++; continue; }
        // extract: $name or ${name}
        size_t k = at; bool brace = false;
        if (k < t.size() && t[k] == '{') { brace = true; k++; }
        const size_t name_at = k;
        while (k < t.size()) {
            const uint8_t c = (uint8_t)t[k];
            if (c >= 0x80) throw Unsupported("a non-ASCII rune right inside a $name of the replace rule (unicode.IsLetter tables are not carried)");
            if (!(c == '_' || (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'))) break;
            k++;
        }
        bool ok = k > name_at;
        const std::string name = t.substr(name_at, k - name_at);
        if (ok && brace) { if (k >= t.size() || t[k] != '}') ok = false; else k++; }
        if (!ok) { lit.push_back('$'); continue; }             // malformed: the $ is text, scanning goes on right behind it
        at = k;
        int64_t num = 0;
        for (char c : name) { if (c < '0' || c > '9' || num >= 100000000) { num = -1; break; } num = num * 10 + (c - '0'); }
        if (name[0] == '0' && name.size() > 1) num = -1;
        flush(); Template::Piece p; p.ref = true;
        if (num >= 0) { if (num <= prog.ncap) p.groups.push_back((int)num); }
        else for (size_t i = 1; i < prog.names.size(); i++) if (prog.names[i] == name) p.groups.push_back((int)i);
        out.pieces.push_back(p);
    }
    lit.append(t, at, std::string::npos); flush();
    return out;
}

inline void replace_all(Machine& m, const Template& tpl, const uint8_t* s, size_t n, std::string& dst) {
    dst.clear();
    size_t last_end = 0, search = 0;
    while (search <= n) {
        if (!m.search(s, n, search)) break;
        const size_t a0 = (size_t)m.matchcap[0], a1 = (size_t)m.matchcap[1];
        dst.append((const char*)s + last_end, a0 - last_end);
        if (a1 > last_end || a0 == 0) {
            for (auto& pc : tpl.pieces) {
                if (!pc.ref) { dst += pc.text; continue; }
                for (int g : pc.groups) if (m.matchcap[2 * g] >= 0) { dst.append((const char*)s + m.matchcap[2 * g], (size_t)(m.matchcap[2 * g + 1] - m.matchcap[2 * g])); break; }
            }
        }
        last_end = a1;
        int w; decode_rune(s + search, n - search, w);
        if (search + w > a1) search += w;
        else if (search + 1 > a1) search++;
        else search = a1;
    }
    dst.append((const char*)s + last_end, n - last_end);
}

}  // namespace tfre
