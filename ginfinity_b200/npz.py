"""Compressed .npz archives written with every host core.

The reference's output format is `numpy.savez_compressed` (cli.py:86, :159): one deflated
`<identifier>.npy` member per record.  NumPy compresses the members one after another on one
thread -- for the 5,840 records of tests/rouskin_sample_6k.tsv that was ~5 s of the `embed`
command's 5.4 s, three orders of magnitude more than the encode.  Here the members are deflated
concurrently (zlib releases the GIL) with the same parameters as `zipfile` (raw deflate, level
6), so every member's compressed stream is the one NumPy would have written, and the ZIP
container is written directly: local headers, central directory, and the ZIP64 records when a
member, the archive or the member count needs them.  `numpy.load` and the reference read the
result like any other .npz.
"""
from __future__ import annotations

import io
import os
import struct
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Iterable, Sequence

import numpy as np

_LIMIT32 = 0xFFFFFFFF
_LIMIT16 = 0xFFFF


def _npy_header(array: np.ndarray) -> bytes:
    buffer = io.BytesIO()
    np.lib.format.write_array_header_1_0(buffer, np.lib.format.header_data_from_array_1_0(array))
    return buffer.getvalue()


def _deflate_member(array) -> tuple:
    """(.npy header + data) of one array, raw-deflated -> (crc32, raw size, compressed bytes)."""
    array = np.asanyarray(array)
    if array.dtype.hasobject:
        raise ValueError("object arrays cannot be written without pickle")
    if not array.flags.c_contiguous:       # np.save writes Fortran-ordered data as such; the
        array = np.ascontiguousarray(array)  # embeddings are C-contiguous, so keep one layout
    header = _npy_header(array)
    data = memoryview(array.reshape(-1).view(np.uint8)) if array.size else b""
    deflater = zlib.compressobj(zlib.Z_DEFAULT_COMPRESSION, zlib.DEFLATED, -15)
    blob = deflater.compress(header) + deflater.compress(data) + deflater.flush()
    crc = zlib.crc32(data, zlib.crc32(header))
    return crc & _LIMIT32, len(header) + array.nbytes, blob


def _dos_time(moment) -> tuple:
    year = max(moment.tm_year, 1980)
    return (moment.tm_hour << 11 | moment.tm_min << 5 | moment.tm_sec // 2,
            (year - 1980) << 9 | moment.tm_mon << 5 | moment.tm_mday)


def write_npz_compressed(path, names: Iterable[str], arrays: Sequence, workers: int = None) -> Path:
    """Write `arrays` under `names` as a deflated .npz; returns the path written (".npz" is
    appended when missing, like numpy.savez).  Equivalent to
    `numpy.savez_compressed(path, **dict(zip(names, arrays)))` for distinct names."""
    path = Path(path)
    if path.suffix != ".npz":
        path = path.with_name(path.name + ".npz")
    names = [str(n) for n in names]
    if len(names) != len(arrays):
        raise ValueError("names and arrays differ in length")
    if len(set(names)) != len(names):
        raise ValueError("duplicate member names in archive")
    workers = workers or min(32, os.cpu_count() or 1)
    dos_time, dos_date = _dos_time(time.localtime())
    central = []
    path.parent.mkdir(parents=True, exist_ok=True)
    with ThreadPoolExecutor(max_workers=workers) as pool, open(path, "wb") as out:
        # map() keeps the order; members are written as soon as their turn comes
        for name, (crc, size, blob) in zip(names, pool.map(_deflate_member, arrays, chunksize=1)):
            fname = (name + ".npy").encode("utf-8")
            flags = 0x800 if any(b > 127 for b in fname) else 0
            offset = out.tell()
            big = size >= _LIMIT32 or len(blob) >= _LIMIT32
            extra = struct.pack("<HHQQ", 1, 16, size, len(blob)) if big else b""
            out.write(struct.pack("<IHHHHHIIIHH", 0x04034B50, 45 if big else 20, flags, 8, dos_time,
                                  dos_date, crc, _LIMIT32 if big else len(blob),
                                  _LIMIT32 if big else size, len(fname), len(extra)))
            out.write(fname)
            out.write(extra)
            out.write(blob)
            central.append((fname, flags, crc, len(blob), size, offset))
        start = out.tell()
        for fname, flags, crc, csize, size, offset in central:
            fields = []
            if size >= _LIMIT32 or csize >= _LIMIT32:
                fields += [size, csize]
            if offset >= _LIMIT32:
                fields.append(offset)
            extra = struct.pack("<HH" + "Q" * len(fields), 1, 8 * len(fields), *fields) if fields else b""
            big_size = size >= _LIMIT32 or csize >= _LIMIT32
            out.write(struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 45 | 3 << 8, 45 if fields else 20,
                                  flags, 8, dos_time, dos_date, crc,
                                  _LIMIT32 if big_size else csize, _LIMIT32 if big_size else size,
                                  len(fname), len(extra), 0, 0, 0, 0o600 << 16,
                                  _LIMIT32 if offset >= _LIMIT32 else offset))
            out.write(fname)
            out.write(extra)
        end = out.tell()
        count, cd_size = len(central), end - start
        if count > _LIMIT16 or cd_size >= _LIMIT32 or start >= _LIMIT32:
            out.write(struct.pack("<IQHHIIQQQQ", 0x06064B50, 44, 45, 45, 0, 0, count, count, cd_size, start))
            out.write(struct.pack("<IIQI", 0x07064B50, 0, end, 1))
        out.write(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, min(count, _LIMIT16), min(count, _LIMIT16),
                              min(cd_size, _LIMIT32), min(start, _LIMIT32), 0))
    return path
