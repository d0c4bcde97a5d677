"""Builds tests/golden/reference_goldens.json from the reference's own golden files and test tables.
Run in the build container only (reads /root/reference); the output is committed.

Sources (reference repo paths):
  mask     pkg/transformer/registry/mask/gotest/canondata/result.json  (expected digests)
           pkg/transformer/registry/mask/hmac_hasher_test.go:23,56-92   (salt + inputs, transcribed below)
  filter   pkg/transformer/registry/filter_rows/filter_rows_test.go:47-560 (input/expected tables, transcribed)
"""
import json, os, math

REF = "/root/reference"
canon = json.load(open(f"{REF}/pkg/transformer/registry/mask/gotest/canondata/result.json"))["gotest.gotest.TestHmacHasherTransformer"]
digests = [canon[i]["Transformed"][0]["columnvalues"] for i in range(3)]
# hmac_hasher_test.go:56-92 — (yt column type, Go dynamic type, value)
mask_inputs = [
    [["utf8", "string", "value1"], ["int64", "int64", 123], ["int32", "int32", 1234], ["boolean", "bool", True]],
    [["string", "string", "value1"], ["date", "time", "1703-01-02T00:00:00Z"], ["double", "float64", 123.123], ["float", "float32", 312.321]],
    [["int8", "int8", -3], ["uint32", "uint32", 12345], ["date", "duration", 60_000_000_000]],
]
mask = {"salt": "the-best-tasty-saint-petersburg-salt",
        "cases": [{"type": t, "go": g, "value": v, "digest": d} for ins, ds in zip(mask_inputs, digests) for (t, g, v), d in zip(ins, ds)]}

I8, I16, I32, I64 = (-2**7, 2**7 - 1), (-2**15, 2**15 - 1), (-2**31, 2**31 - 1), (-2**63, 2**63 - 1)
MAXF32 = 3.4028234663852886e+38
MAXF64 = 1.7976931348623157e+308
def ints(lo, hi): return [lo, 10, 11, 14, 15, 16, hi]
filt = []
# TestIntFiltering :47-145 and TestIntFilteringByFloat :176-270 (Go untyped constants -> `int`)
for fname, f in (("int", "column > 10 AND column <= 15 AND column IN (11, 15)"), ("int-by-float", "column > 10.1 AND column <= 15.1 AND column IN (11.0, 15.0)")):
    for t, (lo, hi) in (("int8", I8), ("int16", I16), ("int32", I32), ("int64", I64), ("uint8", (0, 255)), ("uint16", (0, 65535)), ("uint32", (0, 2**32 - 1)), ("uint64", (0, 2**63 - 1))):
        filt.append({"name": f"{fname}/{t}", "filter": f, "type": t, "go": "int", "input": ints(lo, hi), "expected": [11, 15], "errors": 0})
    filt.append({"name": f"{fname}/uint64-int-overflow", "filter": f, "type": "uint64", "go": "mixed", "input": [["int", 0], ["int", 10], ["int", 11], ["int", 14], ["int", 15], ["int", 16], ["uint64", 2**64 - 1]],
                 "expected": [11, 15], "errors": 1, "error_code": 2})
# TestFloatFiltering :147-174
f = "column >= 10.1 AND column < 15.3 AND column NOT IN (15.2, 11.0)"
for t in ("float", "double"):
    filt.append({"name": f"float/{t}", "filter": f, "type": t, "go": "float64", "input": [-1.0, 10.09, 10.1, 11.0, 14.0, 15.0, 15.2, 15.29, 15.3, 16.0, MAXF32], "expected": [10.1, 14.0, 15.0, 15.29], "errors": 0})
# TestFloatFilteringByInt :272-300
f = "column >= 10 AND column < 15 AND column IN (10.0, 11.0, 14.9)"
filt.append({"name": "float-by-int/float", "filter": f, "type": "float", "go": "float64", "input": [-1.0, 10.0, 10.1, 11.0, 14.0, 14.9, 15.0, MAXF32], "expected": [10.0, 11.0, 14.9], "errors": 0})
filt.append({"name": "float-by-int/double", "filter": f, "type": "double", "go": "float64", "input": [-1.0, 10.0, 10.1, 11.0, 14.0, 14.9, 15.0, MAXF64], "expected": [10.0, 11.0, 14.9], "errors": 0})
# TestBoolFiltering :302-334
filt.append({"name": "bool1", "filter": "column = true AND column != false AND column > false AND column >= false AND column <= true", "type": "boolean", "go": "bool", "input": [True, False], "expected": [True], "errors": 0})
filt.append({"name": "bool2", "filter": "column = false AND column != true AND column < true AND column >= false AND column <= false", "type": "boolean", "go": "bool", "input": [True, False], "expected": [False], "errors": 0})
# TestTimeFiltering :367-388 (times as RFC3339Nano text, zones kept)
t1, t2, t3 = "1986-04-26T01:23:47+03:00", "1990-07-22T00:00:00+04:00", "1990-07-22T00:00:00.001+04:00"
t4, t5, t6 = "1991-12-26T00:00:00+03:00", "2003-04-17T10:19:00+03:00", "2003-04-17T10:19:00.001+03:00"
filt.append({"name": "time", "filter": f"column >= {t2} AND column < {t6} AND column NOT IN ({t3})", "type": "timestamp", "go": "time", "input": [t1, t2, t3, t4, t5, t6], "expected": [t2, t4, t5], "errors": 0})
# TestStringFiltering :390-520 (utf8 column, Go string values)
S1 = ["str", "st", "strr", "tr", "", '"', '""']
S2 = ["ab", "abc", "abca", "abcz", "abd", "ac", "", '"', '""']
S3 = ["ab", "bcc", "bccz", "bcd", "bcda", "bce", "", '"', '""']
S4 = ["str", "st", "sstr", "sttr", "strr", "astrb", "rts", "", '"', '""']
for nm, flt, inp, exp in (
    ("=", 'column = "str"', S1, ["str"]), ("!=", 'column != "str"', S1, ["st", "strr", "tr", "", '"', '""']),
    (">", 'column > "abc"', S2, ["abca", "abcz", "abd", "ac"]), (">=", 'column >= "abc"', S2, ["abc", "abca", "abcz", "abd", "ac"]),
    ("<", 'column < "bcd"', S3, ["ab", "bcc", "bccz", "", '"', '""']), ("<=", 'column <= "bcd"', S3, ["ab", "bcc", "bccz", "bcd", "", '"', '""']),
    ("~", 'column ~ "str"', S4, ["str", "sstr", "strr", "astrb"]), ("!~", 'column !~ "str"', S4, ["st", "sttr", "rts", "", '"', '""']),
    ("in", "column IN ('str', '\"')", S1, ["str", '"'])):
    filt.append({"name": f"string/{nm}", "filter": flt, "type": "utf8", "go": "string", "input": inp, "expected": exp, "errors": 0})
    filt.append({"name": f"bytes/{nm}", "filter": flt, "type": "string", "go": "bytes", "input": inp, "expected": exp, "errors": 0})
B = ["☺str", "st", "ss☺tr", "sttr", "strr☺", "ast☺rb", "rts", "", '"', '""']
filt.append({"name": "bytes/~smiley", "filter": 'column ~ "☺"', "type": "utf8", "go": "bytes", "input": B, "expected": ["☺str", "ss☺tr", "strr☺", "ast☺rb"], "errors": 0})
filt.append({"name": "bytes/in-smiley", "filter": "column IN ('str', '\"', '☺')", "type": "utf8", "go": "bytes", "input": S1 + ["☺"], "expected": ["str", '"', "☺"], "errors": 0})

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
json.dump({"mask": mask, "filter_rows": filt}, open(out, "w"), indent=1, ensure_ascii=False)
print("mask cases", len(mask["cases"]), "filter cases", len(filt))
