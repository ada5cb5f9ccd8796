"""Minimal reader for IT++ `it_file` v3 containers (the format of the reference's
test/*.it fixtures and of `capbuf_NNNN.it`, written by capbuf.cpp:187-197 and
Matlab's itsave).  Layout verified by hexdump:

  magic "IT++", u8 version(3), then per variable
  { u64 hdr_bytes, u64 data_bytes, u64 block_bytes, name\\0, type\\0, desc\\0, payload }
  vectors: u64 n + n items; matrices: u64 rows, u64 cols + column-major items.
All little-endian.
"""
import struct
import numpy as np

_DT = {
    "dvec": ("<f8", 1), "ivec": ("<i4", 1), "bvec": ("u1", 1), "dcvec": ("<c16", 1),
    "dmat": ("<f8", 2), "imat": ("<i4", 2), "bmat": ("u1", 2), "dcmat": ("<c16", 2),
    "fvec": ("<f4", 1), "fcvec": ("<c8", 1),
}


def read_it(path):
    d = open(path, "rb").read()
    if d[:4] != b"IT++" or d[4] != 3:
        raise ValueError("%s: not an IT++ v3 file" % path)
    pos, out = 5, {}
    while pos < len(d):
        hdr, data, block = struct.unpack_from("<QQQ", d, pos)
        p = pos + 24
        strs = []
        for _ in range(3):
            e = d.index(b"\0", p)
            strs.append(d[p:e].decode())
            p = e + 1
        name, typ = strs[0], strs[1]
        q = pos + hdr
        if typ not in _DT:
            raise ValueError("unsupported it_file type %r" % typ)
        dt, nd = _DT[typ]
        if nd == 1:
            (n,) = struct.unpack_from("<Q", d, q)
            arr = np.frombuffer(d, dtype=dt, count=n, offset=q + 8).copy()
        else:
            r, c = struct.unpack_from("<QQ", d, q)
            arr = np.frombuffer(d, dtype=dt, count=r * c, offset=q + 16).reshape((c, r)).T.copy()
        out[name] = arr
        pos += block
    return out


def write_it(path, variables):
    """Write {name: ndarray} as an it_file v3 (dcvec / ivec / dvec only) - the
    `capbuf_NNNN.it` record format of capbuf.cpp:187-197."""
    blob = bytearray(b"IT++" + bytes([3]))
    for name, arr in variables.items():
        arr = np.asarray(arr)
        if np.iscomplexobj(arr):
            typ, payload = "dcvec", arr.astype("<c16").tobytes()
        elif arr.dtype.kind in "iu":
            typ, payload = "ivec", arr.astype("<i4").tobytes()
        else:
            typ, payload = "dvec", arr.astype("<f8").tobytes()
        payload = struct.pack("<Q", arr.size) + payload
        strs = name.encode() + b"\0" + typ.encode() + b"\0" + b"\0"
        hdr = 24 + len(strs)
        blob += struct.pack("<QQQ", hdr, len(payload), hdr + len(payload)) + strs + payload
    open(path, "wb").write(bytes(blob))
