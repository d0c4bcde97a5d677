"""Times tfgpu_rows_to_batch (the host transposer) over the row image of ClickBench-shaped items on this machine's cores:
the same leg bench.py reports as other_paths.host_transpose_rows_to_columns, usable without a GPU while working on host_rows.cu."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from transferia_b200 import rows as rws, workload

nb_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
hb, hschema = bench.make_batch(nb_rows, workload.SEED)
image, off = rws.batch_to_rows(hb)
items = (rws.TfItem * nb_rows)(); offs = off.astype(np.uint64)
for r_ in range(nb_rows):
    it = items[r_]; it.values_off = int(offs[r_]); it.n_values = len(hschema); it.old_keys_off = rws.NO_OLD_KEYS
img = rws.RowsImage([], [("public", "hits", hschema)])
vals = np.frombuffer(image, dtype=np.uint8).copy()
img.struct.n_items = nb_rows; img.struct.items = C.cast(items, C.POINTER(rws.TfItem)); img.struct.values = vals.ctypes.data; img.struct.values_len = len(image)
pool = rws.Columnar()
for th in (1, 2, 4, 8, 16, 64):
    if th > (os.cpu_count() or 1):
        continue
    pool.rows_to_batch(img, threads=th)
    best = 0.0
    for _ in range(5):
        t0 = time.perf_counter(); pool.rows_to_batch(img, threads=th); best = max(best, nb_rows / (time.perf_counter() - t0))
    print(f"threads {th:3d}: {best / 1e6:7.2f} M rows/s  ({best * len(image) / nb_rows / 1e9:.2f} GB/s of row image)")
pool.close()
