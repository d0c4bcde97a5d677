"""GPU aid: correctness + timing of a k_lz4_frames build variant (TFGPU_LIB_PATH=... python scripts/lz4_variant_check.py <frame_bytes>):
the frame stream of a 200 k-row headline batch must decode (oracle decoder: checksums + LZ4) to the oracle's block."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from transferia_b200 import abi, engine, workload
from oracle import pyoracle as po
import bench
fb = int(sys.argv[1]) if len(sys.argv) > 1 else 15360
b, s = bench.make_batch(200_000, workload.SEED)
trs = workload.headline_transformers_watchid(workload.headline_threshold(b, s))
eng = engine.Engine(0, fb)
pid = eng.plan("public", "hits", s, trs, {"type": "clickhouse"})
r = eng.push_encode(pid, b, abi.TF_WIRE_CH_NATIVE_LZ4)
want = po.push_encode(b, po.build_plan("public", "hits", s, trs), abi.TF_WIRE_CH_NATIVE)
raw, nf = po.ch_decode_frames(r.wire)
print("variant", os.environ.get("TFGPU_LIB_PATH", "default"), "frame_bytes", fb, "ok" if raw == want.raw else "MISMATCH", "frames", nf, "ratio %.4f" % (len(want.raw) / len(r.wire)))
