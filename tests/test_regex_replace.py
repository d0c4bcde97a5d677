"""regex_replace_transformer: Go's regexp + ReplaceAll as tfgpu_sink_push applies it to the row image (csrc/host_regex.hpp,
a Pike machine) against oracle/regex_oracle.py (a translation into Python's `re` with Go's replace loop and Expand around it), pinned by
every case of the reference's pkg/transformer/registry/regex_replace/transformer_test.go and by the known answers Go publishes for
Regexp.ReplaceAllString / Expand. Host only: no GPU."""
import os
import random
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import regex_oracle as ro
from transferia_b200 import engine, rows, sink
from transferia_b200.rows import ChangeItem, go

K = rows
R = sink.regex_replace_all


def both(pattern, rule, src: bytes) -> bytes:
    got = R(pattern, rule, src)
    assert got == ro.replace_all(pattern, rule, src), (pattern, rule, src)
    return got


def test_reference_replace_cases():
    """transformer_test.go:167-330 (TestReplace, TestReplaceMultipleMatches, TestReplaceComplexRegex)."""
    for src, want in ((b"test123", b"testNUM"), (b"test", b"test"), (b"", b"")):
        assert both(r"\d+", "NUM", src) == want
    assert ro.replace_value((12, b"test123"), "utf8", r"\d+", "NUM") == (12, b"testNUM")
    assert ro.replace_value((13, b"test123"), "string", r"\d+", "NUM") == (13, b"testNUM")
    assert ro.replace_value((5, 123), "int", r"\d+", "NUM") == (5, 123)                       # non-string type
    assert ro.replace_value((0, None), "utf8", r"\d+", "NUM") == (0, None)                    # nil value
    assert ro.replace_value((5, 123), "utf8", r"\d+", "NUM") == (5, 123)                      # wrong type assertion
    assert both("a", "b", b"banana") == b"bbnbnb"
    assert both(r"(\w+)\s(\w+)", "$2, $1", b"John Doe") == b"Doe, John"
    assert both(r".*?/app/(\d+).*", "$1", b"https://yandex.ru/games/app/99348") == b"99348"


# Known answers of Go's regexp package: the ReplaceAllString example of the package documentation (a(x*)b) and the replace table of its
# tests (src/regexp/all_test.go, replaceTests). Every line is also what the documented algorithm gives by hand.
GO_REPLACE = [
    ("a(x*)b", "T", "-ab-axxb-", "-T-T-"), ("a(x*)b", "$1", "-ab-axxb-", "--xx-"), ("a(x*)b", "$1W", "-ab-axxb-", "---"), ("a(x*)b", "${1}W", "-ab-axxb-", "-W-xxW-"),
    ("", "", "", ""), ("", "x", "", "x"), ("", "", "abc", "abc"), ("", "x", "abc", "xaxbxcx"),
    ("b", "", "", ""), ("b", "x", "", ""), ("b", "", "abc", "ac"), ("b", "x", "abc", "axc"),
    ("y", "", "", ""), ("y", "x", "", ""), ("y", "", "abc", "abc"), ("y", "x", "abc", "abc"),
    ("[a-c]*", "x", "日", "x日x"), ("[^日]", "x", "abc日def", "xxx日xxx"),
    ("^[a-c]*", "x", "abcdabc", "xdabc"), ("[a-c]*$", "x", "abcdabc", "abcdx"), ("^[a-c]*$", "x", "abcdabc", "abcdabc"),
    ("^[a-c]*", "x", "abc", "x"), ("[a-c]*$", "x", "abc", "x"), ("^[a-c]*$", "x", "abc", "x"),
    ("^[a-c]*", "x", "dabce", "xdabce"), ("[a-c]*$", "x", "dabce", "dabcex"), ("^[a-c]*$", "x", "dabce", "dabce"),
    ("^[a-c]*", "x", "", "x"), ("[a-c]*$", "x", "", "x"), ("^[a-c]*$", "x", "", "x"),
    ("^[a-c]+", "x", "abcdabc", "xdabc"), ("[a-c]+$", "x", "abcdabc", "abcdx"), ("^[a-c]+$", "x", "abcdabc", "abcdabc"),
    ("^[a-c]+", "x", "abc", "x"), ("[a-c]+$", "x", "abc", "x"), ("^[a-c]+$", "x", "abc", "x"),
    ("^[a-c]+", "x", "dabce", "dabce"), ("[a-c]+$", "x", "dabce", "dabce"), ("^[a-c]+$", "x", "dabce", "dabce"),
    ("^[a-c]+", "x", "", ""), ("[a-c]+$", "x", "", ""), ("^[a-c]+$", "x", "", ""),
    ("abc", "def", "abcdefg", "defdefg"), ("bc", "BC", "abcbcdcdedef", "aBCBCdcdedef"), ("abc", "", "abcdabc", "d"),
    ("x", "xXx", "xxxXxxx", "xXxxXxxXxXxXxxXxxXx"), ("abc", "d", "", ""), ("abc", "d", "abc", "d"), (".+", "x", "abc", "x"),
    ("[a-c]*", "x", "def", "xdxexfx"), ("[a-c]+", "x", "abcbcdcdedef", "xdxdedef"), ("[a-c]*", "x", "abcbcdcdedef", "xdxdxexdxexfx"),
    # substitutions
    ("a+", "($0)", "banana", "b(a)n(a)n(a)"), ("a+", "(${0})", "banana", "b(a)n(a)n(a)"), ("a+", "(${0})$0", "banana", "b(a)an(a)an(a)a"),
    ("hello, (.+)", "goodbye, ${1}", "hello, world", "goodbye, world"), ("hello, (.+)", "goodbye, $1x", "hello, world", "goodbye, "),
    ("hello, (.+)", "goodbye, ${1}x", "hello, world", "goodbye, worldx"), ("hello, (.+)", "<$0><$1><$2><$3>", "hello, world", "<hello, world><world><><>"),
    ("hello, (?P<noun>.+)", "goodbye, $noun!", "hello, world", "goodbye, world!"), ("hello, (?P<noun>.+)", "goodbye, ${noun}", "hello, world", "goodbye, world"),
    ("hello, (?<noun>.+)", "goodbye, $noun!", "hello, world", "goodbye, world!"),
    ("a+", "$$", "aaa", "$"), ("a+", "$", "aaa", "$"),
    # a subexpression that took no part in the match
    ("(x)?", "$1", "123", "123"), ("abc", "$1", "123", "123"),
    # (x){0}
    ("(a)(b){0}c", ".$1|$2.", "xacxacx", "x.a|.x.a|.x"), ("(a)(((b))){0}c", ".$1|$2.", "xacxacx", "x.a|.x.a|.x"),
    ("((a(b){0}){3}){5}(h)", "y caramb$2", "say " + "a" * 16 + "h", "say ay caramba"), ("((a(b){0}){3}){5}h", "y caramb$2", "say " + "a" * 16 + "h", "say ay caramba"),
]
# a name used twice is legal in Go (Python refuses it, so these run on the product alone): Expand takes the group that took part
GO_REPLACE_PRODUCT_ONLY = [
    ("(?P<x>hi)|(?P<x>bye)", "$x$x$x", "hi", "hihihi"), ("(?P<x>hi)|(?P<x>bye)", "$x$x$x", "bye", "byebyebye"),
    ("(?P<x>hi)|(?P<x>bye)", "$xyz", "hi", ""), ("(?P<x>hi)|(?P<x>bye)", "${x}yz", "hi", "hiyz"),
]


def test_go_published_replace_answers():
    for pat, rule, src, want in GO_REPLACE:
        assert both(pat, rule, src.encode()) == want.encode(), (pat, rule, src)
    for pat, rule, src, want in GO_REPLACE_PRODUCT_ONLY:
        assert R(pat, rule, src.encode()) == want.encode(), (pat, rule, src)


def test_go_specifics():
    """Where Go differs from the usual backtracking engines (each stated in Go's regexp/syntax documentation)."""
    assert both(r"\s", "_", b"a\x0bb c") == b"a\x0bb_c"                                   # \s is [\t\n\f\r ]: no vertical tab
    assert both(r"[[:space:]]", "_", b"a\x0bb") == b"a_b"                                  # the POSIX class has it
    assert both(r"a$", "X", b"a\n") == b"a\n" and both(r"(?m)a$", "X", b"a\na") == b"X\nX"  # $ is the end of the TEXT without (?m)
    assert both(r"a\z", "X", b"a\na") == b"a\nX"
    assert both(r".", "x", b"a\r\nb") == b"xx\nx" and both(r"(?s).", "x", b"a\nb") == b"xxx"
    assert both(r"\w+", "W", "naïve café".encode()) == "WïW Wé".encode()              # \w, \b are ASCII
    assert both(r"\bé", "E", "é aé".encode()) == "é aE".encode()
    assert both(r".", "x", b"a\xff\xfeb") == b"xxxx"                                       # an invalid byte is one rune
    assert both(r"[^a]", "x", b"a\xffb\n") == b"axxx"
    assert both(r"x*", "-", "日本".encode()) == "-日-本-".encode()                           # the loop advances by runes
    assert both(r"a{2}{", "X", b"aa{") == b"X" and both(r"a{,2}", "X", b"a{,2}") == b"X"     # a brace that is no repeat is a literal
    assert both(r"\Qa.b\E+", "X", b"a.bbb a.b") == b"X X"
    assert both(r"[a\-z]+", "X", b"a-z b") == b"X b" and both(r"[]a]+", "X", b"]a]b") == b"Xb" and both(r"[a-]+", "X", b"a-b") == b"Xb"
    assert both(r"\x41\x{65e5}\101\0", "X", "A日A\0".encode()) == b"X"
    assert both(r"(?s:.)\n.", "X", b"\n\n\n") == b"\n\n\n" and both(r"(?s:.)\n.", "X", b"\n\na") == b"X"
    assert both(r"(a|ab)(c|bcd)(d*)", "[$1|$2|$3]", b"abcd") == b"[a|bcd|]"                  # leftmost-first, not leftmost-longest
    assert both(r"(a+)(b+)?", "<$2>", b"aab a") == b"<b> <>"
    assert both(r"(?:(a)|b)+", "<$1>", b"ab ba") == b"<a> <a>"                               # a group keeps what it captured in an earlier round
    assert both(r"a*?", "-", b"aa") == b"-a-a-" and both(r"a+?", "-", b"aa") == b"--" and both(r"a{2,3}?", "-", b"aaaaa") == b"--a"
    # golang.org/issue/46123: a star over an operand that can be empty prefers what (x+)? prefers
    assert R(r"(|a)*", "<$1>", b"aa") == b"<>a<>a<>" and R(r"(|a)+", "<$1>", b"aa") == b"<>a<>a<>"


def test_case_folding():
    """(?i): unicode.SimpleFold orbits for ASCII (Go folds literals, ranges, perl and POSIX classes — the classes before they are negated);
    K (U+212A) folds into k and ſ (U+017F) into s, so both match under (?i) — and so does \\w. Python's own IGNORECASE is not involved
    in the oracle (it would add U+0130 / U+0131)."""
    assert both("(?i)hello", "X", b"Hello hELLO hallo") == b"X X hallo"
    assert both("(?i)k+", "X", "kK\u212a k".encode()) == b"X X" and both("(?i)[r-t]+", "X", "sS\u017f!".encode()) == b"X!"
    assert both(r"(?i)\w+", "X", "s\u017fS\u212a!".encode()) == b"X!" and both(r"(?i)\W", "X", "s\u017f-".encode()) == "s\u017fX".encode()
    assert both(r"\w+", "X", "s\u017f".encode()) == "X\u017f".encode()                               # not without (?i)
    assert both("(?i)[^s]", "X", "s\u017fSt".encode()) == "s\u017fSX".encode()
    assert both("(?i)[[:upper:]]+", "X", b"abC-") == b"X-" and both("(?i)[[:^upper:]]", "X", b"abC-") == b"abCX"
    assert both("(?i:a)b", "X", b"Ab AB ab") == b"X AB X" and both("a(?i)b(?-i)c", "X", b"abc aBc aBC Abc") == b"X X aBC Abc"
    assert both("(?i)i", "X", "iI\u0130\u0131".encode()) == "XX\u0130\u0131".encode()                 # the dotted / dotless i are not in i's orbit
    assert both(r"(?i)\Qk.\E", "X", "K.k.k".encode()) == b"XXk" and both(r"(?i)\x4b", "X", b"k") == b"X"
    # (?U) swaps greedy and lazy
    assert both("(?U)a+", "X", b"aaa") == b"XXX" and both("(?U)a+?", "X", b"aaa") == b"X" and both("(?U:a*)b", "X", b"aab") == b"X"
    assert both("(?U)(a{1,3})(a*)", "[$1|$2]", b"aaaa") == b"[a|][a|][a|][a|]" and both("(?iU)A+?b", "X", b"aab") == b"X"
    assert both("(?i)(?P<w>straSSe)", "<$w>", "STRASSE stra\u017fse".encode()) == "<STRASSE> <stra\u017fse>".encode()
    schema = [{"name": n, "type": "int32"} for n in ("Include", "exclude")]
    d = engine.plan_validate("", "T", schema, [{"convert_to_string": {"tables": {"includeTables": ["(?i)^t$"]}, "columns": {"includeColumns": ["(?i)^INCLUDE$"]}}}])
    assert [c["type"] for c in d["result_schema"]] == ["utf8", "int32"]


@pytest.mark.parametrize("pattern", ["(", ")", "a)", "(?P<n>a", "(?P<>a)", "(?P<a b>c)", "[a", "[z-a]", "a**", "a*+", "a??*", "*a", "|*", "(*)", "a{2}{3}",
                                     r"\1", r"\8", "a\\", r"\C", r"\xZ", r"\x{110000}", r"\y", "a{1001}", "a{2,1}", "(a{500}){3}", "[[:bogus:]]",
                                     "(?z)", "(?-)", "(?s-:a)", "(?<=a)", "(?=a)", "(?!a)", "x{99999999999}", "\udcff"])
def test_expressions_go_refuses(pattern):
    """regexp.Compile's errors (regexp/syntax: missing parens / brackets, bad ranges, stacked or argument-less repetition, backreferences,
    lookarounds, repeat counts above 1000 also when nested): the transformer's constructor fails, transformer.go:19-22."""
    with pytest.raises(engine.EngineError) as ei:
        R(pattern, "", b"x")
    assert ei.value.rc == -1, pattern
    if pattern not in ("\udcff", "(?s-:a)", "a??*", "(?-)", "(a{500}){3}"):      # (the oracle does not restate checkUTF8 / the nested-repeat limit)
        with pytest.raises((ro.GoSyntaxError, NotImplementedError)):
            ro.compile_go(pattern)
    with pytest.raises(engine.EngineError) as ei:
        sink.Sink(transformers=[{"regex_replace_transformer": {"regexMatch": pattern, "replaceRule": ""}}])
    assert ei.value.rc == -1


@pytest.mark.parametrize("pattern", [r"\pL", r"[\p{Greek}]", r"\PN", "(?i)é", "(?i:[а-я])b", r"(?i)[\x00-\x{ffff}]", "(a{30}){30}" * 12])
def test_valid_go_the_library_does_not_carry(pattern):
    with pytest.raises(engine.EngineError) as ei:
        R(pattern, "", b"x")
    assert ei.value.rc == -2
    with pytest.raises(engine.EngineError) as ei:
        sink.Sink(transformers=[{"regex_replace_transformer": {"regexMatch": pattern, "replaceRule": ""}}])
    assert ei.value.rc == -2


_names = iter(range(1000, 10 ** 9))          # a name used twice is legal in Go but not in Python: the generator never repeats one


def _gen(rng, depth=0):
    """A random expression of the shared syntax; repeats are only put on operands that cannot match the empty text (Python and Go agree
    on everything else; Go's own rule for the empty case is pinned by the issue-46123 lines above)."""
    def atom():
        k = rng.random()
        if k < 0.30: return rng.choice(["a", "b", "c", "ab", "日", "\\.", "-", "\\n", "\\x61", "K", "s", "(?i:k)", "(?i:aS)", "(?i:[r-t])", "(?i:\\w)", "(?i:[^ab])"]), False
        if k < 0.45: return rng.choice(["[ab]", "[^a]", "[a-c]", "[^\\n]", "\\d", "\\w", "\\W", "\\s", "\\S", ".", "[\\d_]", "[[:alpha:]]", "[^[:^digit:]x]", "[\\Db]"]), False
        if k < 0.55: return rng.choice(["^", "$", "\\b", "\\B", "\\A", "\\z"]), True
        if k < 0.85 and depth < 3:
            inner, nullable = _gen(rng, depth + 1)
            if "日" not in inner and rng.random() < 0.15: return "(?i:%s)" % inner, nullable
            if rng.random() < 0.08: return "(?U:%s)" % inner, nullable
            return rng.choice(["(%s)", "(?:%s)", "(?P<g%d>%%s)" % next(_names), "(?s:%s)", "(?m:%s)"]) % inner, nullable
        return "", True
    alts = []
    any_nullable = False
    for _ in range(rng.choice([1, 1, 1, 2, 3])):
        parts, nullable = [], True
        for _ in range(rng.choice([1, 2, 2, 3, 4])):
            a, an = atom()
            if a and not an and rng.random() < 0.4:
                q = rng.choice(["*", "+", "?", "{2}", "{1,2}", "{0,2}", "{2,}", "*?", "+?", "??", "{1,3}?"])
                if a[-1:] not in ")]" and len(a) > 1 and not a.startswith("\\") and not a.startswith("["): a = "(?:%s)" % a
                a += q; an = q[0] in "*?" or q.startswith("{0")
            parts.append(a); nullable = nullable and an
        alts.append("".join(parts)); any_nullable = any_nullable or nullable
    return "|".join(alts), any_nullable


class _OracleTooSlow(Exception):
    pass


def test_random_expressions_against_the_oracle():
    import signal

    def on_alarm(_sig, _frm):
        raise _OracleTooSlow()
    signal.signal(signal.SIGALRM, on_alarm)
    rng = random.Random(20260923)
    alphabet = ["a", "b", "c", "ab", "1", "_", " ", "\n", ".", "-", "日", "é", b"\xff", b"\xe6\x97", "x", "A", "B", "k", "K", "S", "s", "\u212a", "\u017f"]
    rules = ["", "X", "<$0>", "[$1|$2]", "${1}x$1x", "$$1", "$", "${g}", "$g1234", "a$0b$9"]
    n_checked = n_changed = 0
    for _ in range(4000):
        pat, _n = _gen(rng)
        rule = rng.choice(rules)
        try:
            ro.compile_go(pat)
        except (ro.GoSyntaxError, NotImplementedError):
            with pytest.raises(engine.EngineError):
                R(pat, rule, b"")
            continue
        for _ in range(4):
            src = b"".join(x if isinstance(x, bytes) else x.encode() for x in (rng.choice(alphabet) for _ in range(rng.randrange(0, 12))))
            if not src and "\\B" in pat:
                continue                    # Python before 3.14 never matches \B on the empty text; Go does (EmptyOpContext(-1, -1) is a no-boundary)
            got = R(pat, rule, src)
            signal.alarm(10)                # Python's backtracking engine can go exponential on nested repeats (the product's machine is linear): such a pair is skipped
            try:
                want = ro.replace_all(pat, rule, src)
            except _OracleTooSlow:
                continue
            finally:
                signal.alarm(0)
            assert got == want, (pat, rule, src)
            n_checked += 1; n_changed += got != src
    assert n_checked > 10000 and n_changed > 3000


def _text_col(ev, c):
    return sink.var_cells(ev, c)


def _push(transformers, items, tables):
    s = sink.Sink(transformers=transformers)
    s.push(rows.RowsImage(items, tables))
    ev = [e for e in s.events if e["type"] == sink.EV_ROWS]
    out = [(e["out"], e["items"], e) for e in ev]
    s.close()
    return out


def test_reference_apply_cases():
    """transformer_test.go:14-164 (TestTransformer_Apply): replaced / table_filter / column_filter through Sinker.Push."""
    schema = [{"name": "column1", "type": "utf8", "key": True}, {"name": "column2", "type": "int64"}, {"name": "column3", "type": "int32"},
              {"name": "column4", "type": "string", "key": True}, {"name": "column5", "type": "boolean"}]
    vals = [go.string("value_1"), go.int64(123), go.int32(1234), go.bytes(b"value@2"), go.bool(True)]
    tr = {"regexMatch": "[_@#&]", "replaceRule": "-"}
    ((out, idx, b),) = _push([{"regex_replace_transformer": tr}], [ChangeItem(K.KIND_INSERT, 0, vals)], [("db", "table1", schema)])
    assert out == ("db", "table1") and idx == [0] and _text_col(b, 0) == [b"value-1"] and _text_col(b, 3) == [b"value-2"]
    assert int(b["columns"][1][0]) == 123 and int(b["columns"][2][0]) == 1234
    # table_filter: exclude `bad_.*` -> not Suitable, the values stay
    ((_, _, b),) = _push([{"regex_replace_transformer": dict(tr, tables={"excludeTables": ["bad_.*"]})}], [ChangeItem(K.KIND_INSERT, 0, vals)], [("db", "bad_table", schema)])
    assert _text_col(b, 0) == [b"value_1"] and _text_col(b, 3) == [b"value@2"]
    # column_filter: exclude `.*_private`
    schema2 = [dict(c, name="column4_private") if c["name"] == "column4" else c for c in schema]
    ((_, _, b),) = _push([{"regex_replace_transformer": dict(tr, columns={"excludeColumns": [".*_private"]})}], [ChangeItem(K.KIND_INSERT, 0, vals)], [("db", "table1", schema2)])
    assert _text_col(b, 0) == [b"value-1"] and _text_col(b, 3) == [b"value@2"]


@pytest.mark.parametrize("on_rows", [False, True])
def test_replace_steps_over_mixed_batches_equal_the_oracle(on_rows, monkeypatch):
    """(tfgpu_sink_push runs the steps on the transposed text columns when every row event lists all its columns — on_rows forces the
    row-image form, which it takes for column subsets, mixed text types and in front of table_splitter; same result.)
    Two chained steps with table and column filters over inserts / updates / deletes of two tables, nil values, control items in between,
    OldKeys untouched (Apply rewrites ColumnValues only, transformer.go:98-118)."""
    if on_rows: monkeypatch.setenv("TFGPU_REGEX_ON_ROWS", "1")
    rng = random.Random(7)
    schema_a = [{"name": "id", "type": "int32", "key": True}, {"name": "name", "type": "utf8"}, {"name": "raw", "type": "string"}, {"name": "note_private", "type": "utf8"}, {"name": "doc", "type": "any"}]
    schema_b = [{"name": "k", "type": "utf8", "key": True}, {"name": "v", "type": "double"}]
    steps = [{"regexMatch": r"(\d+)-(\d+)", "replaceRule": "$2:$1", "columns": {"excludeColumns": [".*_private"]}},
             {"regexMatch": r"^\s+|\s+$", "replaceRule": "", "tables": {"includeTables": ["^a$"]}, "columns": {"includeColumns": ["name", "note.*"]}}]
    words = ["12-34", " x ", "日本 7-8", "", "a-b", "  lead", "trail \n", "99-100-101", "\xff1-2"]
    items = []
    for i in range(300):
        t = rng.randrange(2)
        if rng.random() < 0.05:
            items.append(ChangeItem(rng.choice([K.KIND_DDL, K.KIND_TRUNCATE]), t)); continue
        w = lambda: rng.choice(words).encode("utf-8", "surrogateescape")
        if t == 0:
            v = [go.int32(i), go.string(w()) if rng.random() < 0.9 else go.nil, go.bytes(w()), go.string(w()), go.json(b'"12-34"')]
        else:
            v = [go.string(w()), go.float64(i / 8)]
        kind = rng.choice([K.KIND_INSERT, K.KIND_INSERT, K.KIND_UPDATE, K.KIND_DELETE])
        items.append(ChangeItem(kind, t, v, {0: v[0]} if kind != K.KIND_INSERT else None))
    tables = [("public", "a", schema_a), ("public", "b", schema_b)]
    got = _push([{"regex_replace_transformer": s} for s in steps], items, tables)
    import re as _re
    seen = 0
    for out, idx, b in got:
        t = 0 if out[1] == "a" else 1; schema = tables[t][2]
        for c, col in enumerate(schema):
            if col["type"] not in ("utf8", "string", "any"): continue
            want = []
            for i in idx:
                val = items[i].values[c]
                for s in steps:
                    if "tables" in s and not any(_re.search(p, out[1]) for p in s["tables"]["includeTables"]): continue
                    cf = s.get("columns", {})
                    if any(_re.search(p, col["name"]) for p in cf.get("excludeColumns", [])): continue
                    if cf.get("includeColumns") and not any(_re.search(p, col["name"]) for p in cf["includeColumns"]): continue
                    val = ro.replace_value(val, col["type"], s["regexMatch"], s["replaceRule"])
                want.append(None if val[0] == 0 else val[1])
            assert _text_col(b, c) == want, (out, col["name"])
            seen += len(want)
    assert seen > 600
    # the step has to lead the chain (tfgpu_sink_push rewrites the row image before anything else reads it) and needs tfgpu_sink_push
    with pytest.raises(engine.EngineError) as ei:
        engine.plan_validate("public", "a", schema_a, [{"regex_replace_transformer": steps[0]}])
    assert ei.value.rc == -2
    with pytest.raises(engine.EngineError) as ei:
        engine.plan_validate("public", "a", schema_a, [{"rename_tables": {"renameTables": [{"originalName": {"nameSpace": "public", "name": "a"}, "newName": {"nameSpace": "x", "name": "y"}}]}},
                                                          {"regex_replace_transformer@sink": steps[0]}])
    assert ei.value.rc == -2
    d = engine.plan_validate("public", "a", schema_a, [{"regex_replace_transformer@sink": steps[0]}, {"regex_replace_transformer@sink": steps[1]}, {"convert_to_string": {}}])
    assert [c["name"] for c in d["result_schema"]] == [c["name"] for c in schema_a]


def test_type_is_read_at_the_value_position_for_column_subsets():
    """transformer.go:100-109: the column filter sees ColumnNames[i], the type is TableSchema.Columns()[i] — for an item that carries a column
    subset the i-th value is judged by the i-th schema column's type."""
    schema = [{"name": "id", "type": "int32", "key": True}, {"name": "a", "type": "utf8"}, {"name": "b", "type": "utf8"}]
    tr = [{"regex_replace_transformer": {"regexMatch": "x", "replaceRule": "y"}}]
    # values of columns (id, b): position 1 is typed utf8 (column `a`'s slot) -> replaced; values of (b) alone: position 0 is int32 -> untouched
    items = [ChangeItem(K.KIND_UPDATE, 0, {0: go.int32(1), 2: go.string("xx")}, {0: go.int32(1)}), ChangeItem(K.KIND_UPDATE, 0, {2: go.string("xx")}, {0: go.int32(1)})]
    ((_, _, b),) = _push(tr, items, [("public", "t", schema)])
    assert _text_col(b, 2) == [b"yy", b"xx"]


def test_malformed_row_images_are_refused_by_the_replace_step():
    """The replace step reads the row image before the transposer does: truncated / noisy images, offsets past the end and wild value counts
    come back as error codes (run under the sanitizers by scripts/host_asan.sh)."""
    import ctypes as C
    rng = np.random.default_rng(6)
    schema = [{"name": "a", "type": "int32"}, {"name": "s", "type": "utf8"}, {"name": "b", "type": "string"}]
    good = rows.RowsImage([ChangeItem(K.KIND_INSERT, 0, [go.int32(1), go.string("a_b"), go.bytes(b"c_d")]) for _ in range(40)] +
                          [ChangeItem(K.KIND_UPDATE, 0, {1: go.string("x_y")}, {0: go.int32(0)}) for _ in range(10)], [("d", "t", schema)])
    base = good._vals[:good.values_len].copy()
    s = sink.Sink(transformers=[{"regex_replace_transformer": {"regexMatch": "_", "replaceRule": "--"}}])
    outcomes = set()
    for trial in range(250):
        vals = base.copy(); kind = trial % 5
        if kind == 0: vals = vals[: int(rng.integers(0, len(vals)))]
        elif kind == 1: vals[rng.integers(0, len(vals), 6)] = rng.integers(0, 256, 6)
        elif kind == 2: vals = rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8)
        img = rows.RowsImage([], [("d", "t", schema)])
        items = (rows.TfItem * 50)()
        for r in range(50):
            C.memmove(C.byref(items[r]), C.byref(good._items[r]), C.sizeof(rows.TfItem))
            if kind == 3: items[r].values_off = int(rng.integers(0, 2 ** 40))
            if kind == 4: items[r].n_values = int(rng.integers(0, 2 ** 31))
        buf = np.ascontiguousarray(vals)
        img.struct.n_items = 50; img.struct.items = C.cast(items, C.POINTER(rows.TfItem)); img.struct.values = buf.ctypes.data if len(buf) else None; img.struct.values_len = len(buf)
        s.events.clear()
        try:
            s.push(img); outcomes.add(0)
        except engine.EngineError as ex:
            assert ex.rc in (-2, -3), ex.rc; outcomes.add(ex.rc)
    assert -3 in outcomes
    s.close()


def test_mixed_text_types_keep_the_type_assertion():
    """replace() takes a Go string only in a utf8 column and a []byte only in a `string` column (transformer.go:129-139): a []byte that sits in
    a utf8 column stays as it is — the columnar form of the step notices the mix and hands the run to the row-image form."""
    schema = [{"name": "id", "type": "int32", "key": True}, {"name": "u", "type": "utf8"}, {"name": "b", "type": "string"}]
    items = [ChangeItem(K.KIND_INSERT, 0, [go.int32(1), go.string("a_1"), go.bytes(b"b_1")]),
             ChangeItem(K.KIND_INSERT, 0, [go.int32(2), go.bytes(b"a_2"), go.string("b_2")]),
             ChangeItem(K.KIND_INSERT, 0, [go.int32(3), go.nil, go.bytes(b"")])]
    ((_, _, e),) = _push([{"regex_replace_transformer": {"regexMatch": "_", "replaceRule": "-"}}, {"regex_replace_transformer": {"regexMatch": "^", "replaceRule": ">"}}], items, [("d", "t", schema)])
    assert _text_col(e, 1) == [b">a-1", b"a_2", None] and _text_col(e, 2) == [b">b-1", b"b_2", b">"]
    for it in items: it.values[1] = go.string("x_y") if it.values[1][0] else it.values[1]            # all strings: the columnar form
    items[1].values[2] = go.bytes(b"b_2")
    ((_, _, e),) = _push([{"regex_replace_transformer": {"regexMatch": "_", "replaceRule": "-" * 300}}], items, [("d", "t", schema)])
    assert _text_col(e, 1) == [b"x" + b"-" * 300 + b"y"] * 2 + [None] and _text_col(e, 2) == [b"b" + b"-" * 300 + b"1", b"b" + b"-" * 300 + b"2", b""]


def test_replace_steps_under_the_dispatcher():
    """Three sinks with their own replace steps behind tfgpu_dispatcher (a regexp machine belongs to one thread: the row form keeps one per
    sink, the column form makes one per worker task): 30 batches, every delivered text equals the oracle's."""
    import threading
    schema = [{"name": "id", "type": "int32", "key": True}, {"name": "s", "type": "utf8"}]
    got, lock = {}, threading.Lock()
    def downstream(ev):
        if ev["type"] == sink.EV_ROWS:
            with lock:
                for i, text in zip(ev["columns"][0], sink.var_cells(ev, 1)): got[int(i)] = text
        return 0
    tr = [{"regex_replace_transformer": {"regexMatch": r"(\w+)@(\w+)", "replaceRule": "$2 at $1"}}, {"regex_replace_transformer": {"regexMatch": "(?i)x+", "replaceRule": "-"}}]
    sinks = [sink.Sink(transformers=tr, downstream=downstream) for _ in range(3)]
    d = sink.Dispatcher(sinks)
    want = {}
    for b in range(30):
        its = []
        for i in range(b * 100, b * 100 + 40 + b):
            text = ("user%d@hostXx%d " % (i, i % 7)) * (1 + i % 3)
            want[i] = ro.replace_all("(?i)x+", "-", ro.replace_all(r"(\w+)@(\w+)", "$2 at $1", text.encode()))
            its.append(ChangeItem(K.KIND_INSERT, 0, [go.int32(i), go.string(text)]) if (b % 4 or i % 5) else ChangeItem(K.KIND_UPDATE, 0, {0: go.int32(i), 1: go.string(text)}, {0: go.int32(i)}))
        d.submit(rows.RowsImage(its, [("public", "t", schema)]))
    assert d.drain() == 0
    d.close()
    for s in sinks: s.close()
    assert got == want
