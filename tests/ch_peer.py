"""Test double for the SERVER side of the ClickHouse native protocol (test infrastructure, not product): a small independent Python
implementation that the product's C++ client writer (transferia_b200/csrc/host_chwire.cu) talks to over a socketpair.
It decodes what the client sends field by field (Hello, addendum, Query with client info and settings, Data packets), answers like a
server (Hello, TableColumns, compressed sample block, Log / ProfileEvents / Progress / ProfileInfo, EndOfStream or Exception) and hands the
received frame streams to the test, which decodes them with the oracle and with stock LZ4.
Frames the peer produces itself are compressed with pyarrow's lz4_raw and checksummed with tests/cityhash_independent.py."""
from __future__ import annotations

import struct
import threading
from typing import List, Optional, Tuple

import pyarrow as pa

from cityhash_independent import cityhash128

SERVER_REVISION = 54467      # newer than the client's 54460: the client must negotiate down


class Reader:
    def __init__(self, sock):
        self.s, self.buf, self.pos = sock, b"", 0

    def _need(self, n):
        while len(self.buf) - self.pos < n:
            chunk = self.s.recv(1 << 20)
            if not chunk:
                raise EOFError("client closed the connection")
            self.buf = self.buf[self.pos:] + chunk; self.pos = 0

    def take(self, n) -> bytes:
        self._need(n); out = self.buf[self.pos:self.pos + n]; self.pos += n; return out

    def u8(self): return self.take(1)[0]

    def uvarint(self):
        v = s = 0
        while True:
            b = self.u8(); v |= (b & 0x7f) << s; s += 7
            if not b & 0x80: return v

    def string(self): return self.take(self.uvarint())
    def i32(self): return struct.unpack("<i", self.take(4))[0]
    def i64(self): return struct.unpack("<q", self.take(8))[0]

    def frame(self) -> bytes:
        """One compressed frame, returned whole (checksum + header + body)."""
        head = self.take(25)
        csz = struct.unpack("<I", head[17:21])[0]
        return head + self.take(csz - 9)


def uvarint(v: int) -> bytes:
    out = bytearray()
    while v >= 0x80: out.append((v & 0x7f) | 0x80); v >>= 7
    out.append(v); return bytes(out)


def string(s) -> bytes:
    b = s if isinstance(s, bytes) else s.encode()
    return uvarint(len(b)) + b


def compress_frame(raw: bytes) -> bytes:
    body = pa.compress(raw, codec="lz4_raw", asbytes=True)
    f = b"\x82" + struct.pack("<II", len(body) + 9, len(raw)) + body
    lo, hi = cityhash128(f)
    return struct.pack("<QQ", lo, hi) + f


def frame_payload(frame: bytes) -> bytes:
    """Checksum (independent CityHash128) + LZ4 (pyarrow) of ONE frame."""
    lo, hi = struct.unpack("<QQ", frame[:16])
    assert (lo, hi) == cityhash128(frame[16:]), "frame checksum"
    method = frame[16]; csz, usz = struct.unpack("<II", frame[17:25])
    assert method == 0x82 and csz == len(frame) - 16
    return pa.decompress(frame[25:], decompressed_size=usz, codec="lz4_raw", asbytes=True)


BLOCK_INFO = uvarint(1) + b"\x00" + uvarint(2) + struct.pack("<i", -1) + uvarint(0)
EMPTY_BLOCK = BLOCK_INFO + uvarint(0) + uvarint(0)


def block(cols: List[Tuple[str, str, bytes]], nrows: int) -> bytes:
    out = BLOCK_INFO + uvarint(len(cols)) + uvarint(nrows)
    for name, typ, data in cols:
        out += string(name) + string(typ) + b"\x00" + data
    return out


def log_block() -> bytes:
    """A two-row server Log block (system.text_log shape): the client has to skip it by column type."""
    strs = lambda vals: b"".join(string(v) for v in vals)
    return block([
        ("event_time", "DateTime", struct.pack("<II", 1700000000, 1700000001)),
        ("event_time_microseconds", "UInt32", struct.pack("<II", 1, 2)),
        ("host_name", "String", strs(["ch-1", "ch-1"])),
        ("query_id", "String", strs(["q", "q"])),
        ("thread_id", "UInt64", struct.pack("<QQ", 7, 7)),
        ("priority", "Int8", struct.pack("<bb", 6, 7)),
        ("source", "String", strs(["executeQuery", "MemoryTracker"])),
        ("text", "String", strs(["(from [::1]:1) INSERT", "Peak memory usage: 4.00 MiB. " * 20])),
        ("tags", "Array(String)", struct.pack("<QQ", 1, 3) + strs(["a", "bb", "ccc"])),
        ("maybe", "Nullable(UInt16)", b"\x00\x01" + struct.pack("<HH", 5, 0)),
    ], 2)


class Peer(threading.Thread):
    """Serves ONE connection: handshake, then `n_inserts` INSERT exchanges. Results are left on the object."""

    def __init__(self, sock, sample_cols: List[Tuple[str, str]], expect_raw_len: Optional[List[int]] = None, revision: int = SERVER_REVISION,
                 n_inserts: int = 1, fail_hello: bool = False, fail_insert_with: Optional[Tuple[int, str]] = None, chatter: bool = True):
        super().__init__(daemon=True)
        self.sock, self.sample_cols, self.revision = sock, sample_cols, revision
        self.expect_raw_len = expect_raw_len or []
        self.n_inserts, self.fail_hello, self.fail_insert_with, self.chatter = n_inserts, fail_hello, fail_insert_with, chatter
        self.hello = {}; self.queries = []; self.blocks: List[List[bytes]] = []; self.error: Optional[BaseException] = None

    def run(self):
        try:
            self._serve()
        except BaseException as e:   # noqa: BLE001 - surfaced by the test
            self.error = e
        finally:
            try: self.sock.close()
            except OSError: pass

    def _exception(self, code, msg):
        return uvarint(2) + struct.pack("<i", code) + string("DB::Exception") + string(msg) + string("stack") + b"\x00"

    def _serve(self):
        r = Reader(self.sock)
        assert r.uvarint() == 0, "client Hello expected"
        h = {"name": r.string().decode(), "major": r.uvarint(), "minor": r.uvarint(), "revision": r.uvarint(),
             "database": r.string().decode(), "user": r.string().decode(), "password": r.string().decode()}
        self.hello = h
        if self.fail_hello:
            self.sock.sendall(self._exception(516, "default: Authentication failed")); return
        rev = min(h["revision"], self.revision)
        out = uvarint(0) + string("ClickHouse") + uvarint(24) + uvarint(3) + uvarint(self.revision)
        if self.revision >= 54058: out += string("Europe/Amsterdam")
        if self.revision >= 54372: out += string("ch-test-1")
        if self.revision >= 54401: out += uvarint(7)
        self.sock.sendall(out)
        if rev >= 54458:
            h["quota_key"] = r.string().decode()
        for ins in range(self.n_inserts):
            self._insert(r, rev, ins)

    def _read_query(self, r: Reader, rev: int) -> dict:
        assert r.uvarint() == 1, "Query packet expected"
        q = {"query_id": r.string().decode(), "kind": r.u8(), "initial_user": r.string().decode(), "initial_query_id": r.string().decode(),
             "initial_address": r.string().decode()}
        if rev >= 54449: q["start_time"] = r.i64()
        q["interface"] = r.u8()
        q["os_user"], q["hostname"], q["client_name"] = r.string().decode(), r.string().decode(), r.string().decode()
        q["client_major"], q["client_minor"], q["client_revision"] = r.uvarint(), r.uvarint(), r.uvarint()
        if rev >= 54060: q["quota_key"] = r.string().decode()
        if rev >= 54448: q["distributed_depth"] = r.uvarint()
        if rev >= 54401: q["patch"] = r.uvarint()
        if rev >= 54442: q["otel"] = r.u8(); assert q["otel"] == 0
        if rev >= 54453: q["replicas"] = (r.uvarint(), r.uvarint(), r.uvarint())
        settings = {}
        while True:
            k = r.string()
            if not k: break
            flags = r.uvarint(); settings[k.decode()] = (flags, r.string().decode())
        q["settings"] = settings
        if rev >= 54441: q["secret"] = r.string()
        q["stage"], q["compression"], q["body"] = r.uvarint(), r.uvarint(), r.string().decode()
        if rev >= 54459:
            q["parameters"] = []
            while True:
                k = r.string()
                if not k: break
                q["parameters"].append((k, r.uvarint(), r.string()))
        return q

    def _read_data(self, r: Reader, compressed: bool, expect_len: Optional[int]) -> Tuple[List[bytes], Optional[bytes]]:
        """One Data packet: (frames, raw). With a length hint the frames are collected until that many raw bytes are announced by
        their headers; without one a single frame is read (the empty blocks)."""
        assert r.uvarint() == 2, "Data packet expected"
        assert r.string() == b"", "temporary table name"
        if not compressed:                                   # the raw block itself: its length is known to the test (the empty block's is fixed)
            return [r.take(len(EMPTY_BLOCK) if expect_len is None else expect_len)], None
        frames, total = [], 0
        while True:
            f = r.frame(); frames.append(f); total += struct.unpack("<I", f[21:25])[0]
            if expect_len is None or total >= expect_len:
                break
        return frames, None

    def _insert(self, r: Reader, rev: int, ins: int):
        q = self._read_query(r, rev); self.queries.append(q)
        comp = bool(q["compression"])
        frames, _ = self._read_data(r, comp, None)                       # external tables: one empty block
        assert (frame_payload(frames[0]) if comp else frames[0]) == EMPTY_BLOCK, "external-tables terminator is not the empty block"
        if self.fail_insert_with and ins == 0 and self.fail_insert_with[0] < 0:
            self.sock.sendall(self._exception(-self.fail_insert_with[0], self.fail_insert_with[1])); return
        out = b""
        if self.chatter:
            out += uvarint(11) + string("") + string("columns format version: 1\n")                        # TableColumns
            out += uvarint(3) + uvarint(0) + uvarint(0) + uvarint(0) + uvarint(0) + uvarint(0) + (uvarint(0) if rev >= 54460 else b"")   # Progress
            out += uvarint(10) + string("") + log_block()                                                    # Log, never compressed
        sample = block([(n, t, b"") for n, t in self.sample_cols], 0)
        out += uvarint(1) + string("") + (compress_frame(sample) if comp else sample)
        self.sock.sendall(out)
        got: List[bytes] = []
        k = 0
        while True:
            hint = self.expect_raw_len[k] if k < len(self.expect_raw_len) else None
            frames, _ = self._read_data(r, comp, hint)
            if hint is None:
                assert len(frames) == 1 and (frame_payload(frames[0]) if comp else frames[0]) == EMPTY_BLOCK, "INSERT terminator is not the empty block"
                break
            got.append(b"".join(frames)); k += 1
        self.blocks.append(got)
        if self.fail_insert_with and ins == 0:
            self.sock.sendall(self._exception(*self.fail_insert_with)); return
        out = b""
        if self.chatter:
            out += uvarint(14) + string("") + log_block()                                                    # ProfileEvents (same reader path)
            out += uvarint(6) + uvarint(0) + uvarint(0) + uvarint(0) + b"\x00" + uvarint(0) + b"\x00"        # ProfileInfo
        rows = 1234 + ins
        out += uvarint(3) + uvarint(0) + uvarint(0) + uvarint(0) + uvarint(rows) + uvarint(rows * 10) + (uvarint(5) if rev >= 54460 else b"")
        out += uvarint(5)                                                                                    # EndOfStream
        self.sock.sendall(out)
