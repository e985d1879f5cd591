"""Host I/O either side of the flow: minimal NIfTI-1 reader / writer and scheme-file reader (amico_b200/nifti.py)."""
import gzip
import struct

import numpy as np
import pytest

from amico_b200 import nifti


def test_roundtrip_float32_gz_and_plain(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((5, 4, 3, 7)).astype(np.float32)
    A = np.array([[2, 0, 0, -10], [0, 2.5, 0, 5], [0, 0, 3, 1], [0, 0, 0, 1]], dtype=float)
    for name in ("a.nii.gz", "a.nii"):
        nifti.save(tmp_path / name, a, affine=A, descrip="hello", cal_min=-1, cal_max=2)
        img = nifti.load(tmp_path / name)
        assert img.shape == a.shape and img.data.dtype == np.float32
        np.testing.assert_array_equal(img.data, a)
        np.testing.assert_allclose(img.affine, A)
        assert img.header[148:153] == b"hello"
        assert struct.unpack_from("<2f", img.header, 124) == (2.0, -1.0)
    # x is the fastest axis on disk (Fortran order)
    raw = gzip.open(tmp_path / "a.nii.gz").read()
    first = np.frombuffer(raw, dtype="<f4", count=5, offset=352)
    np.testing.assert_array_equal(first, a[:, 0, 0, 0])


def test_int16_with_scaling_and_template_header(tmp_path):
    a = (np.arange(24).reshape(2, 3, 4) - 5).astype(np.int16)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, 2, 3, 4, 1, 1, 1, 1)
    struct.pack_into("<h", hdr, 70, 4)
    struct.pack_into("<h", hdr, 72, 16)
    struct.pack_into("<8f", hdr, 76, 1, 1.25, 1.25, 2.0, 1, 1, 1, 1)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<f", hdr, 112, 0.5)
    struct.pack_into("<f", hdr, 116, 3.0)
    hdr[344:348] = b"n+1\0"
    with open(tmp_path / "m.nii", "wb") as f:
        f.write(bytes(hdr) + bytes(4) + a.tobytes(order="F"))
    img = nifti.load(tmp_path / "m.nii")
    np.testing.assert_array_equal(img.data, a)
    assert img.zooms == (1.25, 1.25, 2.0)
    np.testing.assert_array_equal(img.get_fdata(), a * 0.5 + 3.0)
    # outputs inherit geometry from the input header (core.py:541-544) but are float32 with unit scaling
    nifti.save(tmp_path / "o.nii.gz", np.ones((2, 3, 4)), like=img)
    out = nifti.load(tmp_path / "o.nii.gz")
    assert out.data.dtype == np.float32 and out.zooms == (1.25, 1.25, 2.0) and out.scl_slope == 1.0 and out.scl_inter == 0.0


def test_rejects_non_nifti(tmp_path):
    (tmp_path / "x.nii").write_bytes(b"\0" * 400)
    with pytest.raises(ValueError):
        nifti.load(tmp_path / "x.nii")


def test_scheme_file_with_header_lines(tmp_path):
    p = tmp_path / "DWI.scheme"
    p.write_text("# comment\nVERSION: BVECTOR\n0 0 0 0\n1 0 0 1000\n0 -1 0 1000\n")
    t = nifti.load_scheme_table(p)
    assert t.shape == (3, 4) and t[2, 1] == -1
