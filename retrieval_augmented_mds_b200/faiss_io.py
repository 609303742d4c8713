"""On-disk interchange with the reference's `index.faiss`.

The reference saves its index with `Dataset.save_faiss_index` (sotasum/mips.py:536), i.e.
`faiss.write_index(index, BufferedIOWriter(PyCallbackIOWriter(f.write)))` (datasets/search.py), and
loads it with `faiss.read_index` (mips.py:547). Only flat indexes are in scope (`string_factory =
"Flat"`), so the file is faiss's flat layout. faiss-cpu 1.7.4 is a third-party wheel that is not
vendored in the reference and not installable here: the layout below is RESTATED from its
published `faiss/impl/index_write.cpp` / `index_read.cpp` (`write_index_header`, `WRITEXBVECTOR`)
and is **parity-unpinned** against a real faiss build (no golden file exists in the reference).

    fourcc   uint32   "IxFI" (IndexFlatIP) | "IxF2" (IndexFlatL2) | "IxFl" (IndexFlat, other metric)
    d        int32
    ntotal   int64
    dummy    int64 x 2   (1 << 20, ignored on read)
    trained  uint8
    metric   int32       (0 = inner product, 1 = L2; > 1 would be followed by a float32 metric_arg)
    n_words  uint64      number of 4-byte words that follow = ntotal * d
    xb       float32[ntotal * d]   row-major

This module is host-side glue (bytes <-> rows); rows enter and leave the HBM shard through the
index's own `add` / `reconstruct_n` (K0 / reconstruct kernels)."""
from __future__ import annotations

import io
import struct
from typing import Callable, Optional, Union

import numpy as np

METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1
_FOURCC = {METRIC_INNER_PRODUCT: b"IxFI", METRIC_L2: b"IxF2"}
_HEADER = struct.Struct("<4siqqqBi")   # fourcc, d, ntotal, dummy, dummy, is_trained, metric_type


class PyCallbackIOWriter:
    """faiss.PyCallbackIOWriter: wraps a `write(bytes)` callable."""

    def __init__(self, write: Callable[[bytes], int]):
        self.write = write


class PyCallbackIOReader:
    """faiss.PyCallbackIOReader: wraps a `read(n) -> bytes` callable."""

    def __init__(self, read: Callable[[int], bytes]):
        self.read = read


class BufferedIOWriter:
    def __init__(self, writer: PyCallbackIOWriter, bsz: int = 1 << 20):
        self.writer, self.bsz = writer, bsz

    def write(self, b: bytes) -> None:
        mv = memoryview(b)
        for i in range(0, len(mv), self.bsz):
            self.writer.write(bytes(mv[i:i + self.bsz]))


class BufferedIOReader:
    def __init__(self, reader: PyCallbackIOReader, bsz: int = 1 << 20):
        self.reader, self.bsz = reader, bsz

    def read(self, n: int) -> bytes:
        chunks, got = [], 0
        while got < n:
            b = self.reader.read(min(self.bsz, n - got))
            if not b:
                break
            chunks.append(b)
            got += len(b)
        return b"".join(chunks)


def _sink(f):
    if isinstance(f, (str, bytes)) or hasattr(f, "__fspath__"):
        fh = open(f, "wb")
        return fh.write, fh.close
    if isinstance(f, (BufferedIOWriter, PyCallbackIOWriter)) or hasattr(f, "write"):
        return f.write, (lambda: None)
    raise TypeError(f"write_index: unsupported destination {type(f)}")


def _source(f):
    if isinstance(f, (str, bytes)) or hasattr(f, "__fspath__"):
        fh = open(f, "rb")
        return fh.read, fh.close
    if isinstance(f, (BufferedIOReader, PyCallbackIOReader)) or hasattr(f, "read"):
        return f.read, (lambda: None)
    raise TypeError(f"read_index: unsupported source {type(f)}")


def write_flat(write: Callable[[bytes], object], rows_iter, d: int, ntotal: int, metric_type: int) -> None:
    """Header + rows; `rows_iter` yields float32 [n_i, d] blocks that add up to ntotal rows."""
    fourcc = _FOURCC.get(int(metric_type), b"IxFl")
    write(_HEADER.pack(fourcc, int(d), int(ntotal), 1 << 20, 1 << 20, 1, int(metric_type)))
    if int(metric_type) > 1:
        write(struct.pack("<f", 0.0))
    write(struct.pack("<Q", int(ntotal) * int(d)))
    seen = 0
    for blk in rows_iter:
        blk = np.ascontiguousarray(blk, dtype="<f4")
        if blk.ndim != 2 or blk.shape[1] != d:
            raise ValueError(f"row block of shape {blk.shape}, expected [n, {d}]")
        write(blk.tobytes())
        seen += blk.shape[0]
    if seen != ntotal:
        raise ValueError(f"wrote {seen} rows, header says {ntotal}")


def read_flat_header(read: Callable[[int], bytes]) -> dict:
    raw = read(_HEADER.size)
    if len(raw) != _HEADER.size:
        raise ValueError("truncated faiss index header")
    fourcc, d, ntotal, _, _, trained, metric = _HEADER.unpack(raw)
    if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
        raise ValueError(f"not a flat faiss index (fourcc {fourcc!r}); only string_factory='Flat' is in scope "
                         "(reference sotasum/mips.py:333-340)")
    if metric > 1:
        read(4)
    (n_words,) = struct.unpack("<Q", read(8))
    if n_words != ntotal * d:
        raise ValueError(f"corrupt flat index: {n_words} words for {ntotal} x {d}")
    return {"d": d, "ntotal": ntotal, "metric_type": metric, "is_trained": bool(trained)}


def write_index(index, f) -> None:
    """faiss.write_index for a B200FlatIndex (stored rows, as float32: a bf16 bank writes its rounded values)."""
    write, close = _sink(f)
    try:
        n, step = int(index.ntotal), 262144
        rows = (index.reconstruct_n(i, min(step, n - i)) for i in range(0, n, step))
        write_flat(write, rows, index.d, n, index.metric_type)
    finally:
        close()


def read_index(f, dtype: str = "fp32", device=None, **kw):
    """faiss.read_index for flat indexes -> B200FlatIndex holding the file's rows (fp32 by default: exact)."""
    from .index import B200FlatIndex

    read, close = _source(f)
    try:
        h = read_flat_header(read)
        idx = B200FlatIndex(h["d"], h["metric_type"], dtype=dtype, device=device, capacity=max(h["ntotal"], 1), **kw)
        step, left = 262144, h["ntotal"]
        while left > 0:
            n = min(step, left)
            raw = read(n * h["d"] * 4)
            if len(raw) != n * h["d"] * 4:
                raise ValueError("truncated faiss index payload")
            idx.add(np.frombuffer(raw, dtype="<f4").reshape(n, h["d"]))
            left -= n
        return idx
    finally:
        close()


def serialize_rows(rows: np.ndarray, metric_type: int) -> bytes:
    """The bytes faiss would write for IndexFlat{IP,L2} holding `rows` (tests / fixtures)."""
    buf = io.BytesIO()
    write_flat(buf.write, [rows], rows.shape[1], rows.shape[0], metric_type)
    return buf.getvalue()
