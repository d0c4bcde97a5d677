"""Debug aid (GPU): push seeded batches through the LZ4 path repeatedly; every frame is decoded with stock liblz4 and a frame that
does not decode (or decodes to other bytes) is dumped with its raw bytes under gpurun_out/."""
import ctypes as C, os, struct, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transferia_b200 import abi, engine, workload

lz = C.CDLL("liblz4.so.1"); lz.LZ4_decompress_safe.restype = C.c_int
eng = engine.Engine(0)
os.makedirs("gpurun_out", exist_ok=True)
bad = 0
for seed in (77, 5, 6, 11, 12, 13):
    batch, schema = workload.make_hits_batch(30_000, seed=seed)
    trs = workload.headline_transformers(workload.counterid_threshold(batch, schema))
    pid = eng.plan("public", "hits", schema, trs, {"type": "clickhouse"})
    raw = eng.push_encode(pid, batch, abi.TF_WIRE_CH_NATIVE).wire
    for rep in range(6):
        wire = eng.push_encode(pid, batch, abi.TF_WIRE_CH_NATIVE_LZ4).wire
        pos = 0; nf = 0; F = eng.frame_bytes
        while pos < len(wire):
            cs, rs = struct.unpack_from("<II", wire, pos + 17)
            dst = C.create_string_buffer(max(1, rs))
            n = lz.LZ4_decompress_safe(wire[pos + 25: pos + 16 + cs], dst, cs - 9, rs)
            want = raw[nf * F: nf * F + rs]
            if n != rs or dst.raw[:rs] != want:
                print("BAD seed", seed, "rep", rep, "frame", nf, "rc", n, "cs", cs, "rs", rs)
                if bad < 4:
                    open(f"gpurun_out/badframe_{bad}.raw", "wb").write(want)
                    open(f"gpurun_out/badframe_{bad}.lz4", "wb").write(wire[pos + 25: pos + 16 + cs])
                bad += 1
            pos += 16 + cs; nf += 1
print("bad frames:", bad)
