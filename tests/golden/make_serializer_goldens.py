"""Serializer goldens: the reference's canonised ChangeItems (tests/canon/**/extracted, typed JSON dumps) whose values the
columnar tf_batch carries losslessly, and the lines the reference's batch serializers wrote for them
(pkg/serializer/reference/canondata/*TestBatchSerializer_{json,csv}_default/result).
Run in the build container only:  python tests/golden/make_serializer_goldens.py  ->  tests/golden/serializer_goldens.json
An item is kept when every column converts: Go ints -> intN/uintN, json.Number/float64 -> double only if the literal is what
FormatFloat(f,'f',-1,64) prints (strictify keeps the TEXT, castx/caste.go:36-60), float32, bool, string -> utf8, []uint8 ->
string, time.Time in UTC, Duration, and `any` made of nil/bool/string/integers/maps/lists."""
import base64, glob, json, os, sys
from datetime import datetime, timezone

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import pyoracle as po  # noqa: E402  (float text forms only)

ROOT = "/root/reference/tests/canon"
GOLD = "/root/reference/pkg/serializer/reference/canondata/reference.reference.TestBatchSerializer_%s_default/result"
INTS = {"int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64", "int", "uint"}
RANGE = {"int8": (-128, 127), "int16": (-32768, 32767), "int32": (-2**31, 2**31 - 1), "int64": (-2**63, 2**63 - 1),
         "uint8": (0, 255), "uint16": (0, 65535), "uint32": (0, 2**32 - 1), "uint64": (0, 2**64 - 1)}


KNOWN = set(RANGE) | {"double", "float", "boolean", "utf8", "string", "date", "datetime", "timestamp", "interval", "any"}


class Raw(str):
    pass


class Skip(Exception):
    pass


def go_json(v):
    """json.Marshal text of an `any` value built from typed dump nodes (sorted keys, HTML escaping) -- simple shapes only."""
    t, x = v.get("type"), v.get("value")
    if t == "nil" or x is None and t not in ("string",):
        return "null"
    if t == "string":
        return go_quote(x)
    if t == "bool":
        return "true" if x else "false"
    if t in INTS:
        return str(int(x))
    if t == "json.Number":
        if isinstance(x, Raw) and x.lstrip("-").isdigit():
            return str(x)
        raise Skip("float in any")
    if t == "map[string]interface {}":
        return "{" + ",".join(go_quote(k) + ":" + go_json(x[k]) for k in sorted(x, key=lambda s: s.encode())) + "}"
    if t == "[]interface {}":
        return "[" + ",".join(go_json(e) for e in x) + "]"
    raise Skip("any of " + str(t))


def go_quote(s):
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch in '"\\':
            out.append("\\" + ch)
        elif ch == "\n": out.append("\\n")
        elif ch == "\r": out.append("\\r")
        elif ch == "\t": out.append("\\t")
        elif ch == "\b": out.append("\\b")
        elif ch == "\f": out.append("\\f")
        elif o < 0x20 or ch in "<>&": out.append("\\u%04x" % o)
        elif o in (0x2028, 0x2029): out.append("\\u%04x" % o)
        else: out.append(ch)
    out.append('"')
    return "".join(out)


def conv(node, yt):
    t, x = node.get("type"), node.get("value")
    if t == "nil":
        return None
    if yt in RANGE:
        if t not in INTS: raise Skip(f"{t}->{yt}")
        lo, hi = RANGE[yt]
        if not lo <= int(x) <= hi: raise Skip("range")
        return int(x)
    if yt == "double":
        if t in ("json.Number", "float64"):
            txt = str(x); f = float(txt)
            if po.fmt_float64(f, 1) != txt: raise Skip("double text not canonical: " + txt[:30])
            return {"f64": txt}
        raise Skip(f"{t}->double")
    if yt == "float":
        if t != "float32": raise Skip(f"{t}->float")
        return {"f32": float(str(x))}
    if yt == "boolean":
        if t != "bool": raise Skip(f"{t}->boolean")
        return bool(x)
    if yt == "utf8":
        if t != "string": raise Skip(f"{t}->utf8")
        return str(x)
    if yt == "string":
        if t == "[]uint8": return {"b64": x}
        if t == "string": return {"b64": base64.b64encode(x.encode()).decode()}
        raise Skip(f"{t}->string")
    if yt in ("date", "datetime", "timestamp"):
        if t != "time.Time" or not x.endswith("Z"): raise Skip("time not UTC")
        main, _, frac = x[:-1].partition(".")
        d = datetime.strptime(main, "%Y-%m-%dT%H:%M:%S").replace(tzinfo=timezone.utc)
        return {"t": [int((d - datetime(1970, 1, 1, tzinfo=timezone.utc)).total_seconds()), int((frac + "000000000")[:9]) if frac else 0]}
    if yt == "interval":
        if t != "time.Duration": raise Skip(f"{t}->interval")
        return {"dur": int(x)}
    if yt == "any":
        if t == "string": return {"any": str(x), "tag": 1}
        return {"any": go_json(node), "tag": 0}
    raise Skip("type " + yt)


cases, skipped = [], {}
for f in sorted(glob.glob(ROOT + "/**/extracted", recursive=True)):
    try:
        items = json.load(open(f), parse_float=Raw, parse_int=Raw)
    except Exception:
        continue
    for it in items:
        if not isinstance(it, dict) or "ColumnValues" not in it or it.get("Kind", {}).get("value") != "insert": continue
        sch = it["TableSchema"]["value"] or []
        names = it["ColumnNames"]["value"] or []
        vals = it["ColumnValues"]["value"] or []
        if len(sch) != len(names) or len(vals) != len(names) or [c["name"] for c in sch] != names or len(set(names)) != len(names): continue
        if any(c["type"] not in KNOWN for c in sch): continue
        try:
            row = [conv(v, c["type"]) for v, c in zip(vals, sch)]
        except Skip as e:
            skipped[str(e)] = skipped.get(str(e), 0) + 1; continue
        schema = [{"name": c["name"], "type": c["type"]} for c in sch]
        for case in cases:
            if case["schema"] == schema:
                if row not in case["rows"]: case["rows"].append(row)
                break
        else:
            cases.append({"schema": schema, "rows": [row], "source": os.path.relpath(f, ROOT)})
out = {"cases": cases, "json_lines": open(GOLD % "json", encoding="utf-8").read().split("\n"), "csv_text": open(GOLD % "csv", encoding="utf-8").read()}
# pkg/serializer/queue/json_serializer_test.go:33-126 TestJSONSerializerTopicNameAllTypes: the value the queue JSON serializer wrote for an
# item with every YT type (the item is not strict: its `string` column holds a Go string, its `timestamp` column a Duration)
out["queue_json_all_types"] = open("/root/reference/pkg/serializer/queue/gotest/canondata/gotest.gotest.TestJSONSerializerTopicNameAllTypes/extracted", encoding="utf-8").read()
json.dump(out, open(os.path.join(os.path.dirname(__file__), "serializer_goldens.json"), "w"), ensure_ascii=False, indent=0)
print(len(cases), "schemas", sum(len(c["rows"]) for c in cases), "rows; skipped:", sorted(skipped.items(), key=lambda kv: -kv[1])[:12])
