"""Flat binary "model blob": the compiled T-rex model handed across the C-ABI.

Layout (little endian)::

    char     magic[8]  = "TREXMDL1"
    uint32   version
    uint32   n_sections
    repeated n_sections times (40 bytes each):
        char     name[24]   (NUL padded)
        uint32   dtype      (0 = float64, 1 = int32)
        uint32   count      (number of elements)
        uint64   offset     (byte offset of the data from the start of the blob)
    data, each section 8-byte aligned

Both the CUDA library (``trex_gym_b200/csrc/trex_capi.cu``) and the CPU oracle
(``oracle/trex_oracle.c``) parse this format by section name, so the same bytes
drive the product and the checker.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

MAGIC = b"TREXMDL1"
VERSION = 3
_F64, _I32 = 0, 1


def pack(sections: "OrderedDict[str, np.ndarray]") -> bytes:
    names = list(sections.keys())
    head = 16 + 40 * len(names)
    off = (head + 7) // 8 * 8
    table = []
    payload = bytearray()
    for name in names:
        arr = np.asarray(sections[name])
        if arr.dtype.kind == "f":
            arr = np.ascontiguousarray(arr, dtype="<f8")
            dt = _F64
        elif arr.dtype.kind in "iub":
            arr = np.ascontiguousarray(arr, dtype="<i4")
            dt = _I32
        else:
            raise TypeError("section %s: unsupported dtype %s" % (name, arr.dtype))
        raw = arr.tobytes()
        bname = name.encode("ascii")
        if len(bname) > 23:
            raise ValueError("section name too long: %s" % name)
        table.append((bname, dt, arr.size, off + len(payload)))
        payload += raw
        payload += b"\0" * ((-len(payload)) % 8)
    out = bytearray()
    out += MAGIC
    out += struct.pack("<II", VERSION, len(names))
    for bname, dt, count, offset in table:
        out += bname.ljust(24, b"\0")
        out += struct.pack("<IIQ", dt, count, offset)
    out += b"\0" * (off - len(out))
    out += payload
    return bytes(out)


def unpack(blob: bytes) -> "OrderedDict[str, np.ndarray]":
    if blob[:8] != MAGIC:
        raise ValueError("not a trex model blob")
    version, n = struct.unpack_from("<II", blob, 8)
    if version != VERSION:
        raise ValueError("model blob version %d, expected %d" % (version, VERSION))
    out = OrderedDict()
    for i in range(n):
        base = 16 + 40 * i
        name = blob[base : base + 24].rstrip(b"\0").decode("ascii")
        dt, count, offset = struct.unpack_from("<IIQ", blob, base + 24)
        if dt == _F64:
            arr = np.frombuffer(blob, dtype="<f8", count=count, offset=offset)
        else:
            arr = np.frombuffer(blob, dtype="<i4", count=count, offset=offset)
        out[name] = arr.copy()
    return out
