"""Staged GPU-vs-oracle check with diagnostics (developer tool; the real suite is tests/ -m gpu)."""
import sys, os, time, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from transferia_b200 import abi, engine, workload
from oracle import pyoracle as po

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
batch, schema = workload.make_hits_batch(n)
k = workload.counterid_threshold(batch, schema)
eng = engine.Engine(0)
for label, trs in (("nofilter", []), ("filter", workload.headline_transformers(k))):
    pid = eng.plan("public", "hits", schema, trs, {"type": "clickhouse"})
    plan = po.build_plan("public", "hits", schema, trs)
    ref = po.push_encode(batch, plan, abi.TF_WIRE_CH_NATIVE_LZ4, eng.frame_bytes)
    res = eng.push_encode(pid, batch, abi.TF_WIRE_CH_NATIVE)
    print(label, "rows", res.rows_out, ref.rows_out, "raw", res.raw_len, len(ref.raw), "errors", len(res.errors), len(ref.errors))
    if res.wire != ref.raw:
        a = np.frombuffer(res.wire, dtype=np.uint8); b = np.frombuffer(ref.raw, dtype=np.uint8)
        m = min(len(a), len(b)); d = np.nonzero(a[:m] != b[:m])[0]
        print("  RAW MISMATCH: first diff at", d[:10], "of", m, "ndiff", len(d))
        if len(d):
            i = int(d[0]); print("   got", a[max(0,i-8):i+24].tobytes(), "\n   exp", b[max(0,i-8):i+24].tobytes())
    else:
        print("  raw block identical")
    t = time.time(); res2 = eng.push_encode(pid, batch, abi.TF_WIRE_CH_NATIVE_LZ4); dt = time.time() - t
    raw, nf = po.ch_decode_frames(res2.wire)
    print("  lz4: wire", len(res2.wire), "frames", res2.n_frames, nf, "ratio", res2.raw_len / max(1, len(res2.wire)), "oracle ratio", len(ref.raw) / len(ref.wire), "call s", dt)
    if raw is None:
        # find the first bad frame
        pos = 0; f = 0; w = res2.wire
        while pos < len(w):
            cs, rs = struct.unpack_from("<II", w, pos + 17)
            blk = w[pos + 25: pos + 16 + cs]
            dec = po.lz4_decompress(blk, rs)
            lo, hi = po.cityhash128(w[pos + 16: pos + 16 + cs])
            ok_h = struct.unpack_from("<QQ", w, pos) == (lo, hi)
            exp = ref.raw[f * eng.frame_bytes: f * eng.frame_bytes + rs]
            if dec is None or dec != exp or not ok_h:
                print("   frame", f, "cs", cs, "rs", rs, "decode", None if dec is None else len(dec), "hash ok", ok_h, "match", dec == exp)
                if dec is not None and dec != exp:
                    x = np.frombuffer(dec, dtype=np.uint8); y = np.frombuffer(exp, dtype=np.uint8); mm = min(len(x), len(y))
                    dd = np.nonzero(x[:mm] != y[:mm])[0]; print("    first diff", dd[:5], "len", len(x), len(y))
                break
            pos += 16 + cs; f += 1
    else:
        print("  lz4 decode ok:", raw == ref.raw)
print("launches", eng.launch_count())
