"""Readers for the reference's committed artefacts (test infrastructure).

Formats: ``results/*_out.csv`` written by
``np.savetxt(fmt='%1.10f', header=..., delimiter=',')`` (``dolfin/bench1.py:210-217``),
and DOLFIN VTU snapshots, either ``format="ascii"`` (bench2, bench6) or
``format="binary"`` with ``vtkZLibDataCompressor`` (bench1).
"""
import base64
import re
import struct
import zlib

import numpy as np

_DTYPES = {"Float64": np.float64, "Float32": np.float32, "UInt32": np.uint32,
           "Int32": np.int32, "UInt8": np.uint8, "Int64": np.int64, "UInt64": np.uint64}


def _decode_zlib_b64(text, dtype):
    """VTK inline compressed array: base64(header uint32[3+nblocks]) + base64(zlib blocks)."""
    text = text.strip()
    # first 3 uint32 -> nblocks; header length (3+nblocks)*4 bytes, base64-padded on its own
    first = base64.b64decode(text[:16])
    nblocks = struct.unpack("<I", first[:4])[0]
    hbytes = (3 + nblocks) * 4
    hchars = ((hbytes + 2) // 3) * 4
    header = np.frombuffer(base64.b64decode(text[:hchars])[:hbytes], dtype="<u4")
    csizes = header[3:]
    payload = base64.b64decode(text[hchars:])
    out, off = [], 0
    for cs in csizes:
        out.append(zlib.decompress(payload[off:off + int(cs)]))
        off += int(cs)
    return np.frombuffer(b"".join(out), dtype=dtype)


def read_vtu(path):
    """Return dict(points=(n,3), cells=(m,k), data={name: (n,)}) from a DOLFIN VTU file."""
    with open(path, "r") as fh:
        txt = fh.read()
    compressed = "vtkZLibDataCompressor" in txt[:400]
    npts = int(re.search(r'NumberOfPoints="(\d+)"', txt).group(1))
    ncells = int(re.search(r'NumberOfCells="(\d+)"', txt).group(1))
    arrays = []
    for m in re.finditer(r"<DataArray([^>]*)>(.*?)</DataArray>", txt, flags=re.S):
        attrs = dict(re.findall(r'(\w+)="([^"]*)"', m.group(1)))
        dtype = _DTYPES[attrs["type"]]
        body = m.group(2)
        if attrs.get("format") == "binary":
            if not compressed:
                raise ValueError("uncompressed binary VTU not supported")
            arr = _decode_zlib_b64(body, dtype)
        else:
            arr = np.array(body.split(), dtype=np.float64).astype(dtype)
        arrays.append((attrs, arr))
    out = {"data": {}}
    pos = txt.find("<PointData")
    for attrs, arr in arrays:
        name = attrs.get("Name", "")
        if name == "connectivity":
            out["cells"] = arr.reshape(ncells, -1).astype(np.int64)
        elif name in ("offsets", "types"):
            continue
        elif int(attrs.get("NumberOfComponents", "1")) == 3 and "points" not in out:
            out["points"] = arr.reshape(npts, 3).astype(np.float64)
        else:
            out["data"][name] = arr.astype(np.float64)
    assert pos > 0 and out["points"].shape[0] == npts
    return out


def read_vtu_scalar(path):
    d = read_vtu(path)
    assert len(d["data"]) == 1, list(d["data"])
    return next(iter(d["data"].values()))


def read_pvd(path):
    """[(time, filename), ...] from a PVD collection file."""
    with open(path) as fh:
        txt = fh.read()
    return [(float(t), f) for t, f in re.findall(r'timestep="([^"]+)"[^>]*file="([^"]+)"', txt)]


def read_csv(path):
    """(header list, rows ndarray) of a reference ``*_out.csv``."""
    with open(path) as fh:
        header = fh.readline().strip().split(",")
    return header, np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2)
