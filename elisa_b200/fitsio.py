"""Minimal FITS reader for the OGIP files of the response-ingest step.

The reference reads PHA / RMF / ARF files with ``astropy.io.fits`` (data/ogip.py:8, 278, 600);
astropy is not a dependency of this path.  This module reads what those files contain and
nothing more (FITS Standard 4.0): the primary HDU is skipped over, ``BINTABLE`` extensions are
decoded -- big-endian fixed-width columns ``L X B I J K A E D``, repeat counts, variable-length
array columns ``P`` / ``Q`` in the heap (the form of RMF ``F_CHAN`` / ``N_CHAN`` / ``MATRIX``
columns), ``TSCALn`` / ``TZEROn`` scaling (unsigned 16-bit channels).  Images, ASCII tables,
compressed tables and complex columns raise ``NotImplementedError``.
"""

from __future__ import annotations

import re

import numpy as np

__all__ = ['HDU', 'read_fits']

_BLOCK = 2880
_FMT = re.compile(r'^\s*(\d*)([LXBIJKAEDCMPQ])(.*)$')
_BASE = {'L': 'u1', 'X': 'u1', 'B': 'u1', 'I': '>i2', 'J': '>i4', 'K': '>i8', 'A': 'S1', 'E': '>f4', 'D': '>f8'}


def _parse_value(text: str):
    text = text.strip()
    if not text:
        return None
    if text[0] == "'":  # string: '' is an escaped quote, trailing blanks are not significant
        out, i = [], 1
        while i < len(text):
            if text[i] == "'":
                if i + 1 < len(text) and text[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(text[i])
            i += 1
        return ''.join(out).rstrip()
    text = text.split('/', 1)[0].strip()
    if text in ('T', 'F'):
        return text == 'T'
    try:
        return int(text)
    except ValueError:
        pass
    try:
        return float(text.replace('D', 'E').replace('d', 'e'))
    except ValueError:
        return text


def _read_header(buf: bytes, pos: int):
    """-> (header dict in card order, position of the data)."""
    header: dict = {}
    while True:
        if pos + _BLOCK > len(buf):
            raise ValueError('truncated FITS header')
        block = buf[pos:pos + _BLOCK].decode('ascii', 'replace')
        pos += _BLOCK
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80]
            key = card[:8].strip()
            if key == 'END':
                return header, pos
            if not key or key in ('COMMENT', 'HISTORY', 'CONTINUE') or card[8:10] != '= ':
                continue
            if key not in header:
                header[key] = _parse_value(card[10:])


class HDU:
    """One header-data unit: ``name`` (EXTNAME), ``ver`` (EXTVER), ``header`` (dict), and for
    binary tables ``names`` (column names in order) and ``columns`` {name: ndarray of shape
    (rows,) / (rows, repeat), or a list of 1-d arrays for variable-length columns}."""

    def __init__(self, header: dict):
        self.header = header
        self.name = str(header.get('EXTNAME', 'PRIMARY')).strip().upper()
        self.ver = int(header.get('EXTVER', 1))
        self.names: list[str] = []
        self.columns: dict = {}
        self.n_rows = 0

    def __getitem__(self, col: str):
        return self.columns[col]

    def __contains__(self, col: str) -> bool:
        return col in self.columns

    def __len__(self) -> int:
        return self.n_rows


def _scaled(arr, header, k):
    scale, zero = header.get(f'TSCAL{k}', 1), header.get(f'TZERO{k}', 0)
    if scale == 1 and zero == 0:
        return arr
    if float(scale) == 1.0 and float(zero) == int(zero) and arr.dtype.kind in 'iu':
        return arr.astype(np.int64) + int(zero)  # the unsigned-integer convention
    return arr.astype(np.float64) * float(scale) + float(zero)


def _bintable(hdu: HDU, buf: bytes, pos: int):
    h = hdu.header
    row_bytes, n_rows, n_fields = int(h['NAXIS1']), int(h['NAXIS2']), int(h['TFIELDS'])
    heap = pos + int(h.get('THEAP', row_bytes * n_rows))
    table = np.frombuffer(buf, dtype=np.uint8, count=row_bytes * n_rows, offset=pos).reshape(n_rows, row_bytes)
    hdu.n_rows = n_rows
    off = 0
    for k in range(1, n_fields + 1):
        name = str(h.get(f'TTYPE{k}', f'COL{k}')).strip().upper()
        m = _FMT.match(str(h[f'TFORM{k}']))
        if not m:
            raise ValueError(f"unreadable TFORM{k} = {h[f'TFORM{k}']!r}")
        repeat = int(m.group(1)) if m.group(1) else 1
        code, rest = m.group(2), m.group(3).strip()
        if code in ('C', 'M'):
            raise NotImplementedError('complex table columns')
        if code in ('P', 'Q'):
            desc = np.dtype('>i4' if code == 'P' else '>i8')
            width = 2 * desc.itemsize * repeat
            et = rest[:1]
            if et not in _BASE or et in ('A', 'X'):
                raise NotImplementedError(f'variable-length column of type {et!r}')
            dt = np.dtype(_BASE[et])
            d = np.ascontiguousarray(table[:, off:off + 2 * desc.itemsize]).view(desc).reshape(n_rows, 2)
            col = []
            for cnt, o in d:
                a = np.frombuffer(buf, dtype=dt, count=int(cnt), offset=heap + int(o))
                col.append(_scaled(a.astype(dt.newbyteorder('=')), h, k))
            hdu.columns[name] = col
        else:
            if code == 'X':
                width = (repeat + 7) // 8
                raw = np.ascontiguousarray(table[:, off:off + width])
                col = np.unpackbits(raw, axis=1)[:, :repeat].astype(bool)
            elif code == 'A':
                width = repeat
                raw = np.ascontiguousarray(table[:, off:off + width])
                col = np.array([bytes(r).decode('ascii', 'replace').rstrip(' \0') for r in raw])
            else:
                dt = np.dtype(_BASE[code])
                width = dt.itemsize * repeat
                raw = np.ascontiguousarray(table[:, off:off + width]).view(dt).reshape(n_rows, repeat)
                col = raw.astype(dt.newbyteorder('=')) if dt.byteorder == '>' else raw.copy()
                if code == 'L':
                    col = col == ord('T')
                else:
                    col = _scaled(col, h, k)
                if repeat == 1:
                    col = col[:, 0]
            hdu.columns[name] = col
        hdu.names.append(name)
        off += width
    if off != row_bytes:
        raise ValueError(f'columns take {off} bytes per row, NAXIS1 says {row_bytes}')


def read_fits(path) -> list[HDU]:
    """All HDUs of a FITS file; binary tables decoded, other data skipped."""
    with open(path, 'rb') as f:
        buf = f.read()
    if buf[:6] != b'SIMPLE':
        raise ValueError(f'{path}: not a FITS file')
    out, pos = [], 0
    while pos < len(buf):
        if not buf[pos:pos + 80].strip(b' \0'):
            break  # trailing padding
        header, pos = _read_header(buf, pos)
        hdu = HDU(header)
        naxis = int(header.get('NAXIS', 0))
        size = 0
        if naxis > 0:
            size = abs(int(header['BITPIX'])) // 8
            for i in range(1, naxis + 1):
                size *= int(header[f'NAXIS{i}'])
            size = (size + int(header.get('PCOUNT', 0))) * int(header.get('GCOUNT', 1))
        xt = str(header.get('XTENSION', '')).strip().upper()
        if xt == 'BINTABLE':
            if str(header.get('ZIMAGE', '')) == 'True' or header.get('ZTABLE'):
                raise NotImplementedError('compressed FITS tables')
            _bintable(hdu, buf, pos)
        elif xt == 'TABLE':
            raise NotImplementedError('ASCII FITS tables')
        out.append(hdu)
        pos += (size + _BLOCK - 1) // _BLOCK * _BLOCK
    return out
