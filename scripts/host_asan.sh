#!/bin/bash
# Host-side C++ (row transposer, sink push / dispatcher, ClickHouse writer) under AddressSanitizer + UBSan: the three host translation units are
# built with g++ into a library of their own, the device entry points of tfgpu.h are stubbed (they answer TF_E_FATAL_NODEVICE), and the CPU tests
# that drive the host code (fuzzers included) run against it. One test asks the real plan builder and fails against the stubs by design.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd); OUT=${1:-/tmp/tfhost_asan}; mkdir -p "$OUT"
python - "$ROOT" "$OUT" <<'PY'
import re, sys
root, out = sys.argv[1], sys.argv[2]
hdr = re.sub(r"/\*.*?\*/", "", open(root + "/include/tfgpu.h").read(), flags=re.S)
protos = re.findall(r"^\s*((?:const\s+)?[\w]+(?:\s*\*)?)\s+(tfgpu_\w+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.M | re.S)
lines = ['#include "%s/include/tfgpu.h"' % root, 'extern "C" {']
for ret, name, args in protos:
    ret = ret.strip(); body = "{}" if ret == "void" else ("{ return TF_E_FATAL_NODEVICE; }" if ret == "int" else "{ return 0; }")
    lines.append("%s %s(%s) %s" % (ret, name, " ".join(args.split()), body))
open(out + "/stubs.cpp", "w").write("\n".join(lines) + "\n}\n")
PY
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -shared -I/usr/local/cuda/include \
    -x c++ "$ROOT/transferia_b200/csrc/host_rows.cu" "$ROOT/transferia_b200/csrc/host_sink.cu" "$ROOT/transferia_b200/csrc/host_chwire.cu" "$OUT/stubs.cpp" \
    -o "$OUT/libtfhost_asan.so" -L/usr/local/cuda/lib64 -lcudart
cd "$ROOT"
TFGPU_LIB_PATH="$OUT/libtfhost_asan.so" LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
    python -m pytest tests/test_rows.py tests/test_sink_push.py tests/test_ch_wire.py -q -m "not gpu" -p no:cacheprovider \
    --deselect tests/test_sink_push.py::test_table_splitter_groups_rows_and_renames_control_items
