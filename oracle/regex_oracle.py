"""CPU restatement (TEST INFRASTRUCTURE — never imported by the product) of the regex_replace_transformer's arithmetic:

    replace()                pkg/transformer/registry/regex_replace/transformer.go:127-142  (utf8 + Go string -> ReplaceAllString,
                             string + []byte -> ReplaceAll, anything else untouched)
    Apply                    transformer.go:87-123 (column filter by name, type by the value's POSITION in the item, :108)

The matching itself is Go's standard library (regexp, regexp/syntax — not part of /root/reference). Independent of the product's Pike
machine (csrc/host_regex.hpp), this restatement translates the expression into Python's `re` (a backtracking engine with the same
leftmost-first rules on the syntax both share) and re-states around it what differs in Go:
  * \\d \\w \\s and the POSIX classes are ASCII sets (\\s without \\v), `$` without (?m) is the very end of the text (Python's \\Z), \\z;
  * text is a rune sequence where every invalid UTF-8 byte is ONE rune (decoded here with surrogateescape);
  * Regexp.ReplaceAll's loop (regexp.go replaceAll): an empty match right behind the previous match is not replaced, the search always
    advances by one rune;
  * Regexp.Expand's template ($1, ${1}, $name, $$; the longest run of letters / digits / _ is the name; `$1x` is the NAME "1x"; a
    malformed $ is text).
  * (?i) folds ASCII letters by unicode.SimpleFold's orbits (k, K also U+212A; s, S also U+017F), perl / POSIX classes folded before they are
    negated — done here on the translated sets, Python's IGNORECASE is not used (it adds U+0130 / U+0131, which Go does not);
Not translated (NotImplementedError; the product refuses the same with TF_E_FATAL_UNSUPPORTED): \\p{..}, (?i) over non-ASCII runes; and flag groups that
are not at the start of the expression or scoped `(?s:...)` (Python cannot state them; the product takes them).

Parity pinned by: every case of the reference's transformer_test.go (TestTransformer_Apply, TestReplace, TestReplaceMultipleMatches,
TestReplaceComplexRegex) in tests/test_regex_replace.py, and by the results Go's own documentation and test tables publish for
ReplaceAllString / Expand (listed there with their source)."""
from __future__ import annotations

import re
from typing import List, Optional, Tuple

RUNE_MAX = 0x10FFFF
_PERL = {"d": [(48, 57)], "w": [(48, 57), (65, 90), (95, 95), (97, 122)], "s": [(9, 10), (12, 13), (32, 32)]}
_POSIX = {"alnum": [(48, 57), (65, 90), (97, 122)], "alpha": [(65, 90), (97, 122)], "ascii": [(0, 127)], "blank": [(9, 9), (32, 32)],
          "cntrl": [(0, 31), (127, 127)], "digit": [(48, 57)], "graph": [(33, 126)], "lower": [(97, 122)], "print": [(32, 126)],
          "punct": [(33, 47), (58, 64), (91, 96), (123, 126)], "space": [(9, 13), (32, 32)], "upper": [(65, 90)],
          "word": [(48, 57), (65, 90), (97, 122), (95, 95)], "xdigit": [(48, 57), (65, 70), (97, 102)]}


class GoSyntaxError(ValueError):
    pass


def _negate(rs):
    rs = sorted(rs); out, nxt = [], 0
    for lo, hi in rs:
        if lo > nxt: out.append((nxt, lo - 1))
        nxt = max(nxt, hi + 1)
    if nxt <= RUNE_MAX: out.append((nxt, RUNE_MAX))
    return out


def _fold(rs):
    """unicode.SimpleFold orbits, ASCII only (what the product carries): the other case of a letter, U+212A for k / K, U+017F for s / S."""
    out = list(rs)
    for lo, hi in rs:
        if hi >= 128: raise NotImplementedError("(?i) over runes outside ASCII")
        for c in range(lo, hi + 1):
            ch = chr(c)
            if ch.isalpha(): out.append((ord(ch.swapcase()),) * 2)
            if ch in "kK": out.append((0x212A, 0x212A))
            if ch in "sS": out.append((0x017F, 0x017F))
    return out


def _u(cp: int) -> str:
    return "\\U%08x" % cp


def _cls(rs, neg=False) -> str:
    if neg: rs = _negate(rs)
    if not rs: return "[^\\U00000000-\\U0010ffff]"
    return "[" + "".join(_u(lo) if lo == hi else _u(lo) + "-" + _u(hi) for lo, hi in rs) + "]"


def _lit(cp: int, fold: bool) -> str:
    if not fold: return _u(cp)
    rs = sorted(set(_fold([(cp, cp)])))
    return _u(cp) if len(rs) == 1 else _cls(rs)


def _escape(p: str, i: int) -> Tuple[int, int]:
    """parseEscape after the backslash: (rune, next index)."""
    if i >= len(p): raise GoSyntaxError("trailing backslash")
    c = p[i]; i += 1
    if c in "1234567":
        if i >= len(p) or p[i] not in "01234567": raise GoSyntaxError("backreference")
    if c in "01234567":
        v = int(c); k = 1
        while k < 3 and i < len(p) and p[i] in "01234567": v = v * 8 + int(p[i]); i += 1; k += 1
        return v, i
    if c == "x":
        if i >= len(p): raise GoSyntaxError("escape")
        if p[i] == "{":
            j = p.find("}", i)
            if j < 0 or j == i + 1 or not all(ch in "0123456789abcdefABCDEF" for ch in p[i + 1:j]): raise GoSyntaxError("escape")
            v = int(p[i + 1:j], 16)
            if v > RUNE_MAX: raise GoSyntaxError("escape")
            return v, j + 1
        h = p[i:i + 2]
        if len(h) < 2 or not all(ch in "0123456789abcdefABCDEF" for ch in h): raise GoSyntaxError("escape")
        return int(h, 16), i + 2
    simple = {"a": 7, "f": 12, "n": 10, "r": 13, "t": 9, "v": 11}
    if c in simple: return simple[c], i
    if ord(c) < 128 and not c.isalnum(): return ord(c), i
    raise GoSyntaxError("escape")


def translate(p: str) -> Tuple[str, int]:
    """Go expression -> (Python expression, flags)."""
    out: List[str] = []; i = 0; flags = re.ASCII
    multi = [False]                                     # (?m) state per open group
    fold = [False]                                      # (?i) state per open group: folded here, Python's IGNORECASE is not used
    ungreedy = [False]                                  # (?U) state per open group: the trailing ? of a repeat is flipped here
    last_repeat = False; can_repeat = False
    while i < len(p):
        c = p[i]; rep = False
        if c == "(":
            if p.startswith("(?P<", i) or (p.startswith("(?<", i) and not p.startswith("(?<=", i) and not p.startswith("(?<!", i)):
                j = p.find(">", i)
                name = p[i + (4 if p[i + 2] == "P" else 3):j] if j >= 0 else ""
                if not name or not re.fullmatch(r"[A-Za-z0-9_]+", name): raise GoSyntaxError("capture name")
                out.append("(?P<%s>" % name); i = j + 1; multi.append(multi[-1]); fold.append(fold[-1]); ungreedy.append(ungreedy[-1]); can_repeat = False
            elif p.startswith("(?", i):
                j = i + 2; fl = ""
                while j < len(p) and p[j] not in ":)": fl += p[j]; j += 1
                if j >= len(p): raise GoSyntaxError("missing )")
                if not all(ch in "imsU-" for ch in fl) or fl.count("-") > 1 or fl.endswith("-"): raise GoSyntaxError("flags")
                on, _, off = fl.partition("-")
                m, f, u = multi[-1], fold[-1], ungreedy[-1]
                if "U" in on: u = True
                if "U" in off: u = False
                if "m" in on: m = True
                if "m" in off: m = False
                if "i" in on: f = True
                if "i" in off: f = False
                strip = lambda t: t.replace("i", "").replace("U", "")
                pyfl = strip(on) + ("-" + strip(off) if strip(off) else "")
                if p[j] == ":":
                    out.append("(?" + pyfl + ":" if pyfl else "(?:"); multi.append(m); fold.append(f); ungreedy.append(u); can_repeat = False
                else:
                    if not fl: raise GoSyntaxError("flags")
                    if pyfl and (i != 0 or "-" in fl): raise NotImplementedError("flag group inside the expression")      # (?i) alone is folded here: fine anywhere
                    if "m" in on: flags |= re.MULTILINE
                    if "s" in on: flags |= re.DOTALL
                    multi[-1] = m; fold[-1] = f; ungreedy[-1] = u; can_repeat = False
                i = j + 1
            else:
                out.append("("); i += 1; multi.append(multi[-1]); fold.append(fold[-1]); ungreedy.append(ungreedy[-1]); can_repeat = False
            last_repeat = False; continue
        if c == ")":
            if len(multi) == 1: raise GoSyntaxError("unexpected )")
            multi.pop(); fold.pop(); ungreedy.pop(); out.append(")"); i += 1; can_repeat = True; last_repeat = False; continue
        if c == "|":
            out.append("|"); i += 1; can_repeat = False; last_repeat = False; continue
        if c in "*+?" or c == "{":
            if c == "{":
                m = re.match(r"\{(\d+)(,(\d*))?\}", p[i:])
                ok = bool(m) and not (len(m.group(1)) > 1 and m.group(1)[0] == "0") and not (m.group(3) and len(m.group(3)) > 1 and m.group(3)[0] == "0")
                if not ok:
                    out.append("\\{"); i += 1; can_repeat = True; last_repeat = False; continue
                lo = int(m.group(1)); hi = lo if m.group(2) is None else (-1 if m.group(3) == "" else int(m.group(3)))
                if lo > 1000 or hi > 1000 or (hi >= 0 and lo > hi): raise GoSyntaxError("repeat size")
                tok = m.group(0); i += len(tok)
            else:
                tok = c; i += 1
            lazy = False
            if i < len(p) and p[i] == "?": lazy = True; i += 1
            if lazy != ungreedy[-1]: tok += "?"
            if last_repeat: raise GoSyntaxError("nested repetition")
            if not can_repeat: raise GoSyntaxError("missing argument to repetition")
            out.append(tok); last_repeat = True; continue
        if c == "^": out.append("^" if multi[-1] else "\\A")
        elif c == "$": out.append("$" if multi[-1] else "\\Z")
        elif c == ".": out.append(".")
        elif c == "[":
            i += 1; neg = False; rs = []; first = True
            if i < len(p) and p[i] == "^": neg = True; i += 1
            while i >= len(p) or p[i] != "]" or first:
                first = False
                if i >= len(p): raise GoSyntaxError("missing ]")
                if p.startswith("[:", i) and p.find(":]", i) >= 0:
                    j = p.find(":]", i); name = p[i + 2:j]; pn = name.startswith("^"); name = name.lstrip("^") if pn else name
                    if name not in _POSIX: raise GoSyntaxError("class")
                    base = _fold(_POSIX[name]) if fold[-1] else _POSIX[name]
                    rs += _negate(base) if pn else base; i = j + 2; continue
                if p[i] == "\\" and i + 1 < len(p) and p[i + 1] in "pP": raise NotImplementedError("\\p")
                if p[i] == "\\" and i + 1 < len(p) and p[i + 1] in "dDsSwW":
                    base = _PERL[p[i + 1].lower()]
                    if fold[-1]: base = _fold(base)
                    rs += _negate(base) if p[i + 1].isupper() else base; i += 2; continue
                if p[i] == "\\": lo, i = _escape(p, i + 1)
                else: lo = ord(p[i]); i += 1
                hi = lo
                if i + 1 < len(p) and p[i] == "-" and p[i + 1] != "]":
                    i += 1
                    if i >= len(p): raise GoSyntaxError("missing ]")
                    if p[i] == "\\": hi, i = _escape(p, i + 1)
                    else: hi = ord(p[i]); i += 1
                    if hi < lo: raise GoSyntaxError("class range")
                rs += _fold([(lo, hi)]) if fold[-1] else [(lo, hi)]
            out.append(_cls(rs, neg)); i += 1; can_repeat = True; last_repeat = False; continue
        elif c == "\\":
            if i + 1 >= len(p): raise GoSyntaxError("trailing backslash")
            e = p[i + 1]
            if e in "pP": raise NotImplementedError("\\p")
            if e == "C": raise GoSyntaxError("\\C")
            if e == "Q":
                j = p.find("\\E", i + 2); lit = p[i + 2:] if j < 0 else p[i + 2:j]
                out.append("".join(_lit(ord(ch), fold[-1]) for ch in lit)); i = len(p) if j < 0 else j + 2
                can_repeat = bool(lit) or can_repeat; last_repeat = False; continue
            if e in "dDsSwW": out.append(_cls(_fold(_PERL[e.lower()]) if fold[-1] else _PERL[e.lower()], e.isupper())); i += 2; can_repeat = True; last_repeat = False; continue
            if e == "A": out.append("\\A"); i += 2; can_repeat = True; last_repeat = False; continue
            if e == "z": out.append("\\Z"); i += 2; can_repeat = True; last_repeat = False; continue
            if e in "bB": out.append("\\" + e); i += 2; can_repeat = True; last_repeat = False; continue
            r, i = _escape(p, i + 1); out.append(_lit(r, fold[-1])); can_repeat = True; last_repeat = False; continue
        else:
            out.append(_lit(ord(c), fold[-1]))
        i += 1; can_repeat = True; last_repeat = rep
    if len(multi) != 1: raise GoSyntaxError("missing )")
    return "".join(out), flags


def compile_go(pattern: str):
    py, flags = translate(pattern)
    try:
        return re.compile(py, flags)
    except re.error as e:                       # what is left: e.g. a repeat of nothing the checks above let through
        raise GoSyntaxError(str(e))


def parse_template(rule: str):
    """Regexp.expand's view of the template: ("lit", text) | ("num", k) | ("name", n)."""
    out = []; t = rule
    while t:
        k = t.find("$")
        if k < 0: break
        if k: out.append(("lit", t[:k]))
        t = t[k + 1:]
        if t.startswith("$"): out.append(("lit", "$")); t = t[1:]; continue
        s = t; brace = s.startswith("{")
        if brace: s = s[1:]
        i = 0
        while i < len(s) and (s[i] == "_" or s[i].isalpha() or s[i].isdigit()):          # unicode.IsLetter / IsDigit
            if ord(s[i]) >= 128: raise NotImplementedError("non-ASCII rune in a $name")
            i += 1
        if i < len(s) and ord(s[i]) >= 128 and i == 0: raise NotImplementedError("non-ASCII rune behind $")
        name = s[:i]; ok = i > 0
        if ok and brace:
            if i >= len(s) or s[i] != "}": ok = False
            else: i += 1
        if not ok: out.append(("lit", "$")); continue
        t = s[i:]
        num = 0
        for ch in name:
            if not ("0" <= ch <= "9") or num >= 10 ** 8: num = -1; break
            num = num * 10 + int(ch)
        if name[0] == "0" and len(name) > 1: num = -1
        out.append(("num", num) if num >= 0 else ("name", name))
    if t: out.append(("lit", t))
    return out


def replace_all(pattern: str, rule: str, src: bytes) -> bytes:
    """regexp.MustCompile(pattern).ReplaceAll(src, []byte(rule))."""
    rx = compile_go(pattern); tpl = parse_template(rule)
    s = src.decode("utf-8", "surrogateescape")
    dst: List[str] = []; last_end = 0; pos = 0
    while pos <= len(s):
        m = rx.search(s, pos)
        if not m: break
        a0, a1 = m.span()
        dst.append(s[last_end:a0])
        if a1 > last_end or a0 == 0:
            for kind, v in tpl:
                if kind == "lit": dst.append(v)
                elif kind == "num":
                    if v <= rx.groups and m.group(v) is not None: dst.append(m.group(v))
                elif v in rx.groupindex and m.group(v) is not None: dst.append(m.group(v))
        last_end = a1
        w = 1 if pos < len(s) else 0
        if pos + w > a1: pos += w
        elif pos + 1 > a1: pos += 1
        else: pos = a1
    dst.append(s[last_end:])
    return "".join(dst).encode("utf-8", "surrogateescape")


def replace_value(value, typ: str, pattern: str, rule: str):
    """replace() transformer.go:127-142 over the (tag, value) pairs of transferia_b200.rows (12 Go string, 13 []byte)."""
    tag, v = value
    if (typ == "utf8" and tag == 12) or (typ == "string" and tag == 13):
        return (tag, replace_all(pattern, rule, v))
    return value
