"""Minimal NIfTI-1 single-file reader / writer (``.nii`` / ``.nii.gz``) and scheme-file reader.

The reference goes through nibabel (``amico/core.py:135-141, 180-182, 541-648``), which is not available here; this
covers what that flow needs: the image array in its on-disk dtype, the 348-byte header (kept as raw bytes so that it can
be handed on to the outputs like ``core.py:541-544`` does with ``hdr``), ``scl_slope`` / ``scl_inter``, ``pixdim``, the
affine, and writing float32 volumes with ``cal_min`` / ``cal_max`` / ``descrip`` set (``core.py:546-648``).
"""
from __future__ import annotations

import gzip
import re
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


class NiftiImage:
    """``header`` (bytes, 348), ``data`` (numpy array in Fortran memory order, logical shape = NIfTI dim), ``endian``."""

    def __init__(self, header, data, endian="<"):
        self.header, self.data, self.endian = bytes(header), data, endian

    def _get(self, fmt, off):
        return struct.unpack_from(self.endian + fmt, self.header, off)

    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def zooms(self):
        return tuple(float(v) for v in self._get("8f", 76)[1:1 + self.data.ndim])

    @property
    def scl_slope(self):
        return float(self._get("f", 112)[0])

    @property
    def scl_inter(self):
        return float(self._get("f", 116)[0])

    @property
    def affine(self):
        """sform when set, else the qform-less scaling matrix (enough to carry geometry from input to output)."""
        A = np.eye(4)
        if self._get("h", 254)[0] > 0:
            A[0], A[1], A[2] = self._get("4f", 280), self._get("4f", 296), self._get("4f", 312)
        else:
            z = self._get("8f", 76)
            A[0, 0], A[1, 1], A[2, 2] = z[1], z[2], z[3]
            A[:3, 3] = self._get("3f", 268)
        return A

    def get_fdata(self):
        """Like nibabel: float64 array with the header's scaling applied when it is meaningful."""
        a = self.data.astype(np.float64)
        s, i = self.scl_slope, self.scl_inter
        if np.isfinite(s) and np.isfinite(i) and s != 0 and (s != 1 or i != 0):
            a = a * s + i
        return a


def load(path):
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        raw = f.read()
    if len(raw) < 352:
        raise ValueError(f"{path}: not a NIfTI-1 file")
    endian = "<"
    if struct.unpack_from("<i", raw, 0)[0] != 348:
        endian = ">"
        if struct.unpack_from(">i", raw, 0)[0] != 348:
            raise ValueError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    if raw[344:347] not in (b"n+1",):
        raise ValueError(f"{path}: only single-file NIfTI-1 (magic 'n+1') is supported")
    dim = struct.unpack_from(endian + "8h", raw, 40)
    nd = dim[0]
    if not 1 <= nd <= 7:
        raise ValueError(f"{path}: bad dim[0]={nd}")
    shape = tuple(int(d) for d in dim[1:1 + nd])
    code = struct.unpack_from(endian + "h", raw, 70)[0]
    if code not in _DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype {code}")
    dt = np.dtype(_DTYPES[code]).newbyteorder(endian)
    off = int(struct.unpack_from(endian + "f", raw, 108)[0])
    n = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=n, offset=max(off, 352)).reshape(shape, order="F")
    if endian == ">":
        data = data.astype(dt.newbyteorder("<"))
    return NiftiImage(raw[:348], data, endian)


def save(path, array, like=None, affine=None, descrip=None, cal_min=None, cal_max=None):
    """Write ``array`` as float32 (datatype 16, bitpix 32: what ``core.py:543-544`` sets), geometry taken from ``like``."""
    a = np.asarray(array, dtype=np.float32)
    hdr = bytearray(like.header if like is not None else bytes(348))
    if like is not None and like.endian == ">":
        raise ValueError("writing from a big-endian template is not supported")
    struct.pack_into("<i", hdr, 0, 348)
    dim = [a.ndim] + list(a.shape) + [1] * (7 - a.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<h", hdr, 70, 16)
    struct.pack_into("<h", hdr, 72, 32)
    if like is None:
        struct.pack_into("<8f", hdr, 76, 1.0, *([1.0] * 7))
        A = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
        struct.pack_into("<h", hdr, 254, 1)
        for r, o in enumerate((280, 296, 312)):
            struct.pack_into("<4f", hdr, o, *A[r])
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<f", hdr, 112, 1.0)   # scl_slope
    struct.pack_into("<f", hdr, 116, 0.0)   # scl_inter
    if cal_max is not None:
        struct.pack_into("<f", hdr, 124, float(cal_max))
    if cal_min is not None:
        struct.pack_into("<f", hdr, 128, float(cal_min))
    if descrip is not None:
        d = descrip.encode("ascii", "replace")[:79]
        hdr[148:228] = d + bytes(80 - len(d))
    hdr[344:348] = b"n+1\0"
    payload = bytes(hdr) + bytes(4) + np.asfortranarray(a).tobytes(order="F")
    opener = gzip.open if str(path).endswith(".gz") else open
    kw = {"compresslevel": 1} if str(path).endswith(".gz") else {}
    with opener(path, "wb", **kw) as f:
        f.write(payload)


def load_scheme_table(path):
    """The numeric table of a scheme file, skipping header lines exactly like ``amico/scheme.py:28-39``."""
    n = 0
    with open(path) as fid:
        for line in fid:
            if re.match(r"[+-]?(\d+(\.\d*)?|\.\d+)([eE][+-]?\d+)?", line.strip()):
                break
            n += 1
    return np.loadtxt(path, skiprows=n)
