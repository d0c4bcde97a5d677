"""Extracts the reference's own unit-test vectors for the typesystem casts into tests/golden/cast_goldens.json:

  pkg/abstract/changeitem/strictify/strictify_test.go:54-684   Strictify: (target YT type, Go value) -> ok | error
  pkg/abstract/restore_test.go:14-135                          Restore:   (target YT type, Go value) -> Go value
  pkg/csv/splitter_test.go:14-106                              Splitter:  bytes -> rows (quote-aware)

Run in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_cast_goldens.py
Go values are kept as {"go": <Go type>, "v": <text>}: integers / floats as decimal text, []byte as hex, time.Time as
"<unix seconds>.<nanoseconds>", time.Duration as nanoseconds, maps as JSON text. Expressions the extractor cannot evaluate
(struct values, uuid, yt date helpers) are listed under "skipped" with their source line, not silently dropped."""
import datetime
import json
import math
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cast_goldens.json")

YT = {"Int8": "int8", "Int16": "int16", "Int32": "int32", "Int64": "int64", "Uint8": "uint8", "Uint16": "uint16", "Uint32": "uint32", "Uint64": "uint64",
      "Float32": "float", "Float64": "double", "String": "utf8", "Bytes": "string", "Date": "date", "Datetime": "datetime", "Timestamp": "timestamp",
      "Interval": "interval", "Boolean": "boolean", "Any": "any"}
MATH = {"math.MinInt8": -128, "math.MaxInt8": 127, "math.MinInt16": -32768, "math.MaxInt16": 32767, "math.MinInt32": -2 ** 31, "math.MaxInt32": 2 ** 31 - 1,
        "math.MinInt64": -2 ** 63, "math.MaxInt64": 2 ** 63 - 1, "math.MaxUint8": 255, "math.MaxUint16": 65535, "math.MaxUint32": 2 ** 32 - 1, "math.MaxUint64": 2 ** 64 - 1,
        "math.SmallestNonzeroFloat32": 1.401298464324817070923729583289916131280e-45, "math.MaxFloat32": 3.40282346638528859811704183484516925440e+38,
        "math.SmallestNonzeroFloat64": 5e-324, "math.MaxFloat64": 1.79769313486231570814527423731704356798070e+308}
DUR = {"time.Nanosecond": 1, "time.Microsecond": 1000, "time.Millisecond": 10 ** 6, "time.Second": 10 ** 9, "time.Minute": 60 * 10 ** 9, "time.Hour": 3600 * 10 ** 9}


class Skip(Exception):
    pass


def num(expr: str):
    e = expr
    for k in sorted(MATH, key=len, reverse=True):
        e = e.replace(k, repr(MATH[k]))
    if not re.fullmatch(r"[0-9eE+\-*/. ()]+", e):
        raise Skip(expr)
    return eval(e, {"__builtins__": {}})


def unix(y, mo, d, h, mi, s):
    days = (datetime.date(y, mo, d) - datetime.date(1970, 1, 1)).days
    return days * 86400 + h * 3600 + mi * 60 + s


def go_value(expr: str):
    e = expr.strip()
    if e == "nil":
        return {"go": "nil", "v": ""}
    if e in ("true", "false"):
        return {"go": "bool", "v": e}
    m = re.fullmatch(r"(u?int(?:8|16|32|64)?|float32|float64)\((.*)\)", e)
    if m:
        t, v = m.group(1), num(m.group(2))
        if t.startswith("float"):
            return {"go": t, "v": repr(float(v))}
        return {"go": t, "v": str(int(v))}
    m = re.fullmatch(r'json\.Number\("([^"]*)"\)', e)
    if m:
        return {"go": "json.Number", "v": m.group(1)}
    m = re.fullmatch(r'"((?:[^"\\]|\\.)*)"', e)
    if m:
        return {"go": "string", "v": json.loads(e)}
    m = re.fullmatch(r"`([^`]*)`", e)
    if m:
        return {"go": "string", "v": m.group(1)}
    m = re.fullmatch(r"\[\]byte\{(.*)\}", e)
    if m:
        body = m.group(1).strip()
        return {"go": "[]byte", "v": bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", body)).hex()}
    m = re.fullmatch(r'\[\]byte\("((?:[^"\\]|\\.)*)"\)', e)
    if m:
        return {"go": "[]byte", "v": json.loads('"%s"' % m.group(1)).encode().hex()}
    m = re.fullmatch(r"time\.Date\((\d+), (\d+), (\d+), (\d+), (\d+), (\d+), (\d+), time\.UTC\)", e)
    if m:
        y, mo, d, h, mi, s, ns = (int(x) for x in m.groups())
        return {"go": "time.Time", "v": "%d.%09d" % (unix(y, mo, d, h, mi, s), ns)}
    m = re.fullmatch(r"(?:(\d+) \* )?(time\.(?:Nanosecond|Microsecond|Millisecond|Second|Minute|Hour))", e)
    if m:
        return {"go": "time.Duration", "v": str(int(m.group(1) or 1) * DUR[m.group(2)])}
    m = re.fullmatch(r"map\[string\](?:string|float32|float64|interface\{\})\{(.*)\}", e)
    if m:
        body = re.sub(r'json\.Number\("([^"]*)"\)', r"\1", m.group(1))
        return {"go": "map", "v": "{" + body + "}"}
    if e in ("ts", "&ts", "ts.UTC()"):
        return {"go": "time.Time", "v": "1609459200.000000000"}
    raise Skip(expr)


def split_args(s: str):
    """top-level comma split of a Go argument list"""
    out, depth, cur, q = [], 0, "", None
    for ch in s:
        if q:
            cur += ch
            if ch == q and not cur.endswith("\\" + q):
                q = None
            continue
        if ch in "\"`":
            q = ch
        if ch in "([{":
            depth += 1
        if ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def strictify_cases():
    src = open(os.path.join(REF, "pkg/abstract/changeitem/strictify/strictify_test.go"), encoding="utf-8").read()
    schemas = {}
    for m in re.finditer(r"var (\w+Sch) = changeitem\.NewTableSchema\(\[\]changeitem\.ColSchema\{(.*?)\n\}\)", src, re.S):
        cols = re.findall(r'ColumnName:\s*"(\w+)",\s*DataType:\s*schema\.Type(\w+)\.String\(\)', m.group(2))
        schemas[m.group(1)] = [(c, YT[t]) for c, t in cols]
    cases, skipped = [], []
    for m in re.finditer(r"func (TestStrictify\w+)\(t \*testing\.T\) \{(.*?)\n\}\n", src, re.S):
        name, body = m.group(1), m.group(2)
        sch = schemas[re.search(r"MakeFastTableSchema\((\w+)\.Columns\(\)\)", body).group(1)]
        expect = "ok" if "executePositiveStrictifyCheck" in body else "error"
        line0 = src[:m.start()].count("\n") + 1
        for k, it in enumerate(re.finditer(r"changeItemWithValues\(template, \[\]interface\{\}\{\n(.*?)\n\t\t\}\),", body, re.S)):
            vals = []
            for ln in it.group(1).split("\n"):
                ln = ln.strip()
                if not ln:
                    continue
                mm = re.match(r"(.*?),\s*//\s*(\w+)", ln)
                expr, col = mm.group(1), mm.group(2)
                try:
                    vals.append({"col": col, "type": dict(sch)[col], "value": go_value(expr)})
                except Skip:
                    skipped.append({"test": name, "item": k, "expr": expr})
                    vals = None
                    break
            if vals is not None:
                cases.append({"test": name, "line": line0, "item": k, "expect": expect, "values": vals})
    return cases, skipped


def restore_cases():
    src = open(os.path.join(REF, "pkg/abstract/restore_test.go"), encoding="utf-8").read()
    cases, skipped = [], []
    for i, ln in enumerate(src.split("\n")[:135], 1):
        s = ln.strip()
        if s.startswith("//") or "Restore(" not in s or not s.startswith("assert.Equal(t, "):
            continue
        inner = s[len("assert.Equal(t, "):s.rindex(")")]
        inner = inner.split(") //")[0] if ") //" in inner else inner
        args = split_args(inner)
        if len(args) != 2:
            skipped.append({"line": i, "expr": s}); continue
        call, other = (args[0], args[1]) if args[0].startswith("Restore(") else (args[1], args[0])
        cm = re.fullmatch(r"Restore\((.*)\)", call)
        if not cm:
            skipped.append({"line": i, "expr": s}); continue
        a = split_args(cm.group(1))
        tm = re.fullmatch(r'colSchema\("(\w*)", (?:false|true)\)', a[0])
        if a[0] == "col":
            typ = "int64" if i < 112 else "any"      # TestRestoreFloatInt64 / TestRestoreJSONB
        elif tm:
            typ = tm.group(1)
        else:
            skipped.append({"line": i, "expr": s}); continue
        try:
            cases.append({"line": i, "type": typ.lower(), "in": go_value(a[1]), "want": go_value(other)})
        except Skip:
            skipped.append({"line": i, "expr": s})
    return cases, skipped


def splitter_cases():
    """The four behavioural tests of splitter_test.go (the fifth counts Write calls of the bufio part-by-part path)."""
    return [
        {"test": "TestScannerBasic", "line": 14, "input": "a\nb", "rows": ["a\n"], "eof_rest": "b"},
        {"test": "TestScannerBiggerLines", "line": 32, "input": "12345678901234567890\n12345\n", "rows": ["12345678901234567890\n", "12345\n"], "eof_rest": ""},
        {"test": "TestScannerQuotes", "line": 55, "input": "\"234567890123456789\"\n\"2345\"\n", "rows": ["\"234567890123456789\"\n", "\"2345\"\n"], "eof_rest": ""},
        {"test": "TestScannerLineBreaksInsideQuotes", "line": 78, "input": "\"23456789012345\"\"89\n123456\"", "rows": [], "eof_rest": "\"23456789012345\"\"89\n123456\""},
    ]


if __name__ == "__main__":
    sc, ss = strictify_cases()
    rc, rs = restore_cases()
    out = {"source": {"strictify": "pkg/abstract/changeitem/strictify/strictify_test.go:54-684", "restore": "pkg/abstract/restore_test.go:14-135",
                      "splitter": "pkg/csv/splitter_test.go:14-106"},
           "strictify": sc, "strictify_skipped": ss, "restore": rc, "restore_skipped": rs, "splitter": splitter_cases()}
    json.dump(out, open(OUT, "w", encoding="utf-8"), indent=1, ensure_ascii=False)
    print(len(sc), "strictify items,", len(ss), "skipped;", len(rc), "restore cases,", len(rs), "skipped;", len(out["splitter"]), "splitter cases")
