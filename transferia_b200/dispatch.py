"""Multi-GPU dealing of batches: rows are independent, so N GPUs = N engines and a host
dispatcher — no collective on the data path.

  * `partition_rows`  contiguous row ranges per rank (one process per GPU under torchrun: bench.py, tests)
  * `RoundRobinDispatcher`  one host process driving several engines: whole batches are dealt round-robin,
    every batch carries a sequence number and results are re-emitted in submission order, which keeps the
    in-table row order the reference preserves (pkg/transformer/transformation.go:131-141 keeps per-table
    order, not cross-table order).
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, Iterable, Iterator, List, Sequence, Tuple


def partition_rows(nrows: int, world: int) -> List[Tuple[int, int]]:
    """[lo, hi) per rank; sizes differ by at most one row; empty ranges when world > nrows."""
    base, rem = divmod(nrows, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi)); lo = hi
    return out


class RoundRobinDispatcher:
    """workers[i](batch) -> result, each bound to its own engine/device; calls on one worker never overlap
    (pkg/abstract/sink.go:12 — a Sinker is never called concurrently)."""

    def __init__(self, workers: Sequence[Callable]):
        self.workers = list(workers)
        self._in = [queue.Queue(maxsize=2) for _ in self.workers]
        self._out: "queue.Queue" = queue.Queue()
        self._threads = [threading.Thread(target=self._loop, args=(i,), daemon=True) for i in range(len(self.workers))]
        for t in self._threads:
            t.start()

    def _loop(self, i: int):
        while True:
            item = self._in[i].get()
            if item is None:
                return
            seq, batch = item
            try:
                self._out.put((seq, self.workers[i](batch), None))
            except Exception as e:  # delivered to the caller in order, like a failed Push
                self._out.put((seq, None, e))

    def run(self, batches: Iterable) -> Iterator:
        """Yields results in submission order."""
        pending, next_seq, submitted = {}, 0, 0
        it = iter(batches)
        exhausted = False
        while not exhausted or next_seq < submitted:
            while not exhausted and submitted - next_seq < 2 * len(self.workers):
                try:
                    b = next(it)
                except StopIteration:
                    exhausted = True
                    break
                self._in[submitted % len(self.workers)].put((submitted, b)); submitted += 1
            if next_seq < submitted:
                while next_seq not in pending:
                    seq, res, err = self._out.get()
                    pending[seq] = (res, err)
                res, err = pending.pop(next_seq); next_seq += 1
                if err is not None:
                    raise err
                yield res

    def close(self):
        for q in self._in:
            q.put(None)
        for t in self._threads:
            t.join(timeout=5)
