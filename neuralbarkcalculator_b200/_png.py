"""Minimal PNG writer for the two outputs of the hot path (``processed/`` RGB images, models.py:203, and the 0/127/255
dual images, models.py:349-356): 8-bit RGB or grey, one filter type for all rows, one zlib stream in one IDAT chunk.
level 0 = stored blocks (about 2 ms per image); level 1 (default) = the library's own encoder (``nbc_png_idat`` in
libnbc.so, host code: Sub filter + run-length matches + per-segment Huffman codes, about 4 ms for a 1024x624 RGB image
where zlib's Z_RLE takes 27 ms and PIL at its level 1 about 85 ms); levels >= 2 = zlib at that level.  ctypes and zlib
release the GIL, so a thread pool scales."""
import ctypes as C
import struct
import zlib

import numpy as np

from . import _lib

_SIG = b'\x89PNG\r\n\x1a\n'


def _chunk(tag, data):
    return struct.pack('>I', len(data)) + tag + data + struct.pack('>I', zlib.crc32(tag + data) & 0xFFFFFFFF)


def _native_idat(arr, lut=None):
    """-> (IHDR payload, numpy buffer, n): the zlib stream of the image data in buffer[:n] (nbc_png_idat, GIL released)."""
    lib = _lib.load()
    h, w = arr.shape[:2]
    ch = 3 if arr.ndim == 3 else 1
    cap = lib.nbc_png_idat_bound(h, w, ch)
    out = np.empty(cap, dtype=np.uint8)
    n = lib.nbc_png_idat(C.c_void_p(arr.ctypes.data), h, w, ch, 0, C.c_void_p(lut.ctypes.data) if lut is not None else None,
                         C.c_void_p(out.ctypes.data), cap)
    if n < 0:
        raise RuntimeError('nbc_png_idat: ' + lib.nbc_last_error().decode())
    return struct.pack('>IIBBBBB', w, h, 8, 2 if ch == 3 else 0, 0, 0, 0), out, n


def _check(arr, lut):
    arr = np.ascontiguousarray(arr)
    if arr.dtype != np.uint8 or arr.ndim not in (2, 3) or (arr.ndim == 3 and arr.shape[2] != 3):
        raise ValueError('encode_png: uint8 [H,W] or [H,W,3] expected')
    if lut is not None:
        lut = np.ascontiguousarray(lut)
        if arr.ndim != 2 or lut.dtype != np.uint8 or lut.shape != (256,):
            raise ValueError('encode_png: lut is a uint8 [256] table for grey images')
    return arr, lut


def encode_png(arr, level=1, lut=None, _force_zlib=False):
    """arr: uint8 [H,W,3] (RGB) or [H,W] (grey) -> PNG file bytes.  lut (grey only): 256-entry table applied to every
    pixel first (the 0/127/255 dual image straight from a class mask).

    level 0: stored (no compression); level 1: Sub filter (left-pixel prediction; None for grey masks, whose long runs
    need no prediction) + run-length / Huffman deflate by the library's own encoder (nbc_png_idat); levels >= 2: the same
    filter with zlib's default strategy at that level.  ``encode_png_zlib_rle`` is the former level-1 path (zlib Z_RLE),
    kept as the cross-check of the native encoder."""
    arr, lut = _check(arr, lut)
    h, w = arr.shape[:2]
    if level == 1 and not _force_zlib:
        ihdr, out, n = _native_idat(arr, lut)
        return _SIG + _chunk(b'IHDR', ihdr) + _chunk(b'IDAT', out[:n].tobytes()) + _chunk(b'IEND', b'')
    if lut is not None:
        arr = lut[arr]
    ihdr = struct.pack('>IIBBBBB', w, h, 8, 2 if arr.ndim == 3 else 0, 0, 0, 0)
    rows = arr.reshape(h, -1)
    raw = np.empty((h, rows.shape[1] + 1), dtype=np.uint8)
    if arr.ndim == 3 and level > 0 and w > 1:
        raw[:, 0] = 1                  # filter type 1 (Sub): byte - byte of the pixel to the left, modulo 256
        raw[:, 1:4] = rows[:, :3]
        np.subtract(rows[:, 3:], rows[:, :-3], out=raw[:, 4:])
    else:
        raw[:, 0] = 0                  # filter type 0 (None)
        raw[:, 1:] = rows
    strategy = zlib.Z_RLE if level == 1 else zlib.Z_DEFAULT_STRATEGY
    comp = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
    idat = comp.compress(raw) + comp.flush()
    return _SIG + _chunk(b'IHDR', ihdr) + _chunk(b'IDAT', idat) + _chunk(b'IEND', b'')


def encode_png_zlib_rle(arr, lut=None):
    return encode_png(arr, 1, lut, _force_zlib=True)


def write_png(path, arr, level=1, lut=None):
    if level == 1:
        # no megabyte-sized Python copies: the IDAT payload goes from the encoder's buffer to the file; the CRC-32 is a
        # running one over the tag and the payload (zlib.crc32 releases the GIL)
        arr, lut = _check(arr, lut)
        ihdr, out, n = _native_idat(arr, lut)
        body = memoryview(out)[:n]
        crc = zlib.crc32(body, zlib.crc32(b'IDAT')) & 0xFFFFFFFF
        with open(path, 'wb') as f:
            f.write(_SIG + _chunk(b'IHDR', ihdr) + struct.pack('>I', n) + b'IDAT')
            f.write(body)
            f.write(struct.pack('>I', crc) + _chunk(b'IEND', b''))
        return
    data = encode_png(arr, level, lut)
    with open(path, 'wb') as f:
        f.write(data)
