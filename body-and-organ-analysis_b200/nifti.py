"""Minimal NIfTI-1 reader / writer (single-file .nii / .nii.gz) - the on-disk format at both ends of the path
(nibabel / SimpleITK are not used; _external/totalsegmentator/nnunet.py:402-414,723-726,779 and
_external/totalsegmentator/nifti_ext_header.py:12-42 for the label-table extension).

Arrays are returned as C-ordered [z, y, x] (= the file's Fortran-ordered [x, y, z] memory, what SimpleITK's
GetArrayFromImage gives) together with the 4x4 voxel->world affine of the (x, y, z) index.
"""
from __future__ import annotations

import gzip
import json
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


@dataclass
class NiftiImage:
    data: np.ndarray          # [z, y, x]
    affine: np.ndarray        # 4x4, maps (i_x, i_y, i_z, 1) to world (RAS+ mm)
    header_bytes: bytes = b""

    @property
    def zooms(self) -> tuple[float, float, float]:
        """Voxel sizes along x, y, z (nibabel's header.get_zooms(); SimpleITK's GetSpacing())."""
        return tuple(float(np.sqrt((self.affine[:3, i] ** 2).sum())) for i in range(3))


def _open(path, mode):
    # compression level 1 is nibabel's default for .nii.gz (Opener.default_compresslevel); Python's default of 9 takes
    # 5.4 s instead of 0.33 s for a 256 x 512 x 512 label map (measured) for a file half the size
    if str(path).endswith(".gz"):
        return gzip.open(path, mode, compresslevel=1) if "w" in mode else gzip.open(path, mode)
    return open(path, mode)


def load(path) -> NiftiImage:
    with _open(path, "rb") as f:
        raw = f.read()
    if struct.unpack("<i", raw[:4])[0] == 348:
        e = "<"
    elif struct.unpack(">i", raw[:4])[0] == 348:
        e = ">"
    else:
        raise ValueError(f"{path}: not a NIfTI-1 file")
    dim = struct.unpack(e + "8h", raw[40:56])
    if dim[0] < 3:
        raise ValueError("TotalSegmentator does not work for 2D images. Use a 3D image.")
    nx, ny, nz = dim[1:4]
    datatype = struct.unpack(e + "h", raw[70:72])[0]
    pixdim = struct.unpack(e + "8f", raw[76:108])
    vox_offset = int(struct.unpack(e + "f", raw[108:112])[0])
    slope, inter = struct.unpack(e + "2f", raw[112:120])
    qform_code, sform_code = struct.unpack(e + "2h", raw[252:256])
    if datatype not in _DTYPES:
        raise TypeError(f"Invalid dtype code {datatype}. Expected a simple dtype, not a structured one.")
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(e)
    n = nx * ny * nz  # only the first volume of a 4-D file (nnunet.py:405-407)
    data = np.frombuffer(raw, dtype=dt, count=n, offset=vox_offset).reshape(nz, ny, nx)
    data = np.ascontiguousarray(data.astype(dt.newbyteorder("="), copy=False))
    # scl_slope / scl_inter as nibabel applies them (nifti1.py get_slope_inter + get_fdata): a zero or non-finite slope
    # means "no scaling"; a valid slope with a non-finite intercept is an error; scaled data are float64 (a uint16 file
    # with inter = -1024 must not wrap, a fractional slope must not be truncated back to the stored dtype)
    if np.isfinite(slope) and slope != 0.0:
        if not np.isfinite(inter):
            raise ValueError(f"{path}: valid scl_slope ({slope}) but invalid scl_inter ({inter})")
        if (slope, inter) != (1.0, 0.0):
            data = data.astype(np.float64) * np.float64(slope) + np.float64(inter)
    if sform_code > 0:
        aff = np.eye(4)
        aff[0] = struct.unpack(e + "4f", raw[280:296])
        aff[1] = struct.unpack(e + "4f", raw[296:312])
        aff[2] = struct.unpack(e + "4f", raw[312:328])
    elif qform_code > 0:
        b, c, d = struct.unpack(e + "3f", raw[256:268])
        off = struct.unpack(e + "3f", raw[268:280])
        a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
        R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                      [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                      [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
        qfac = -1.0 if pixdim[0] < 0 else 1.0
        aff = np.eye(4)
        aff[:3, :3] = R * np.array([pixdim[1], pixdim[2], pixdim[3] * qfac])
        aff[:3, 3] = off
    else:
        aff = np.diag([pixdim[1], pixdim[2], pixdim[3], 1.0])
    return NiftiImage(data, aff.astype(np.float64), raw[:348])


def save(path, data_zyx: np.ndarray, affine: np.ndarray, label_map: dict | None = None) -> None:
    """Write [z, y, x] data; label_map (id -> name) goes into a header extension as TotalSegmentator does."""
    data = np.ascontiguousarray(data_zyx)
    if data.dtype == bool:
        data = data.astype(np.uint8)
    code = _CODES[data.dtype]
    nz, ny, nx = data.shape
    ext = b""
    if label_map is not None:
        payload = json.dumps({"labels": {str(k): v for k, v in label_map.items()}}).encode()
        esize = (len(payload) + 8 + 15) // 16 * 16
        ext = struct.pack("<2i", esize, 0) + payload.ljust(esize - 8, b"\0")
    vox_offset = 352 + len(ext)
    h = bytearray(348)
    struct.pack_into("<i", h, 0, 348)
    struct.pack_into("<8h", h, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into("<h", h, 70, code)
    struct.pack_into("<h", h, 72, data.dtype.itemsize * 8)
    zooms = [float(np.sqrt((affine[:3, i] ** 2).sum())) for i in range(3)]
    struct.pack_into("<8f", h, 76, 1.0, zooms[0], zooms[1], zooms[2], 1.0, 1.0, 1.0, 1.0)
    struct.pack_into("<f", h, 108, float(vox_offset))
    struct.pack_into("<2f", h, 112, 1.0, 0.0)
    h[123] = 2  # xyzt_units: mm
    struct.pack_into("<2h", h, 252, 0, 2)  # sform only (aligned)
    struct.pack_into("<4f", h, 280, *affine[0])
    struct.pack_into("<4f", h, 296, *affine[1])
    struct.pack_into("<4f", h, 312, *affine[2])
    h[344:348] = b"n+1\0"
    head = bytes(h) + struct.pack("4B", 1 if ext else 0, 0, 0, 0) + ext
    threads = int(os.environ.get("BOA_B200_GZIP_THREADS", str(min(8, os.cpu_count() or 1))))
    if str(path).endswith(".gz") and threads > 1 and data.nbytes > _GZIP_CHUNK:
        _write_gzip_members(path, head, memoryview(data).cast("B"), threads)
        return
    with _open(path, "wb") as f:
        f.write(head)
        f.write(data.tobytes())


_GZIP_CHUNK = 4 << 20


def _write_gzip_members(path, head: bytes, body: memoryview, threads: int) -> None:
    """Parallel deflate: the byte stream is cut into 4 MB chunks, every chunk is deflated on its own
    thread (zlib releases the GIL) into a complete gzip MEMBER, and the members are written back to back.  A
    concatenation of members is a valid gzip file (RFC 1952 2.2) that gzip / zlib's gzread / nibabel / ITK read as one
    stream; level 1 as nibabel.  Costs < 0.1 % of file size against a single stream (one 18-byte frame + a cold
    dictionary per 4 MB)."""
    def member(lo: int) -> bytes:
        c = zlib.compressobj(1, zlib.DEFLATED, 31)
        chunk = body[lo:lo + _GZIP_CHUNK]
        return c.compress(head + bytes(chunk) if lo == 0 else chunk) + c.flush()

    with open(path, "wb") as f, ThreadPoolExecutor(max_workers=threads) as ex:
        for blob in ex.map(member, range(0, len(body), _GZIP_CHUNK)):
            f.write(blob)


# ---- orientation (nib.as_closest_canonical / undo, totalsegmentator/alignment.py:8-54)
def canonical_ornt(affine: np.ndarray):
    """For each world axis (R, A, S): which voxel axis (x=0, y=1, z=2) it follows and with which sign."""
    R = affine[:3, :3] / np.maximum(np.sqrt((affine[:3, :3] ** 2).sum(axis=0)), 1e-12)
    order = [None] * 3  # greedy assignment, strongest direction cosine first
    Rabs = np.abs(R).copy()
    for _ in range(3):
        w, v = np.unravel_index(np.argmax(Rabs), Rabs.shape)
        order[w] = (int(v), 1 if R[w, v] > 0 else -1)
        Rabs[w, :] = -1
        Rabs[:, v] = -1
    return order


def to_canonical(data_zyx: np.ndarray, affine: np.ndarray):
    """-> (data [z,y,x] in RAS+ voxel order, spacing (sx, sy, sz) of the canonical axes, undo-info)."""
    order = canonical_ornt(affine)
    xyz = data_zyx.transpose(2, 1, 0)  # [x, y, z] view
    perm = [o[0] for o in order]
    out = xyz.transpose(perm)
    for ax, (_, sgn) in enumerate(order):
        if sgn < 0:
            out = np.flip(out, axis=ax)
    zooms = [float(np.sqrt((affine[:3, v] ** 2).sum())) for v in perm]
    return np.ascontiguousarray(out.transpose(2, 1, 0)), tuple(zooms), order


def from_canonical(data_zyx: np.ndarray, order) -> np.ndarray:
    out = data_zyx.transpose(2, 1, 0)
    for ax, (_, sgn) in enumerate(order):
        if sgn < 0:
            out = np.flip(out, axis=ax)
    inv = np.argsort([o[0] for o in order])
    return np.ascontiguousarray(out.transpose(inv).transpose(2, 1, 0))
