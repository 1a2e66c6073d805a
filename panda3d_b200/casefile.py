"""Binary container ("P3CSBLOB") for golden fixtures and recipes.

Layout (little endian): magic "P3CSBLOB", u32 version, u32 count, then per entry
u32 name_len, name (padded to 4), u32 dtype, u32 ndim, u64 shape[ndim], u64 nbytes,
data (padded to 8).  oracle/ref_harness.cxx reads and writes the same format, so the
reference built from /root/reference and this repository exchange cases through it.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

_DTYPES = {0: np.uint8, 1: np.int32, 2: np.uint32, 3: np.float32, 4: np.uint16, 5: np.int64, 6: np.float64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}
MAGIC = b"P3CSBLOB"


def save_case(path: str, blobs: Dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", 1, len(blobs)))
        for name in sorted(blobs):
            arr = blobs[name]
            if isinstance(arr, (bytes, str)):
                raw = arr.encode() if isinstance(arr, str) else arr
                arr = np.frombuffer(raw, dtype=np.uint8)
            arr = np.ascontiguousarray(arr)
            if arr.dtype not in _CODES:
                raise TypeError(f"{name}: unsupported dtype {arr.dtype}")
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(b"\0" * ((4 - (len(nb) & 3)) & 3))
            f.write(struct.pack("<II", _CODES[arr.dtype], arr.ndim))
            for s in arr.shape:
                f.write(struct.pack("<Q", s))
            data = arr.tobytes()
            f.write(struct.pack("<Q", len(data)))
            f.write(data)
            f.write(b"\0" * ((8 - (len(data) & 7)) & 7))


def load_case(path: str) -> Dict[str, np.ndarray]:
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != MAGIC:
        raise ValueError(f"{path}: not a P3CSBLOB file")
    _ver, count = struct.unpack_from("<II", buf, 8)
    pos = 16
    out: Dict[str, np.ndarray] = {}
    for _ in range(count):
        (nl,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        name = buf[pos:pos + nl].decode()
        pos += nl + ((4 - (nl & 3)) & 3)
        dt, nd = struct.unpack_from("<II", buf, pos)
        pos += 8
        shape = struct.unpack_from("<%dQ" % nd, buf, pos) if nd else ()
        pos += 8 * nd
        (nbytes,) = struct.unpack_from("<Q", buf, pos)
        pos += 8
        itemsize = np.dtype(_DTYPES[dt]).itemsize
        arr = np.frombuffer(buf, dtype=_DTYPES[dt], count=nbytes // itemsize, offset=pos)
        out[name] = arr.reshape(shape).copy()
        pos += nbytes + ((8 - (nbytes & 7)) & 7)
    return out


def text_lines(arr: np.ndarray):
    s = np.asarray(arr, dtype=np.uint8).tobytes().decode()
    return [ln for ln in s.split("\n") if ln != ""]
