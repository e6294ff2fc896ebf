"""Cross-section tables of the absorption components.

The reference keeps the photo-electric cross-sections of PhAbs / TBAbs / WAbs in
``elisa/models/tables/xsect.hdf5`` and reads them with h5py at import time
(models/mul.py:286-299): dataset ``energy`` plus ``{phabs|tbabs|wabs}/{xsect}/{abund}``,
float32, written by ``tables/generate_xsect_table.py`` with h5py's defaults.  h5py is not a
dependency of this path, so this module carries a minimal reader for exactly that subset of
the HDF5 file format (HDF5 File Format Specification, version 0/1 superblock):

* old-style groups -- symbol-table message, version-1 B-tree (``TREE``), symbol-table nodes
  (``SNOD``), local heap (``HEAP``);
* version-1 object headers with continuation blocks;
* simple dataspaces, fixed-point / IEEE floating-point datatypes of either byte order;
* contiguous or compact data layout (layout message version 3), no filters.

Anything else (new-style groups, chunked or filtered data, version-2 object headers) raises
``NotImplementedError``: read the arrays with h5py and hand them to the component instead.
"""

from __future__ import annotations

import struct

import numpy as np

__all__ = ['read_hdf5', 'load_xsect', 'XSECT_MODELS']

XSECT_MODELS = ('phabs', 'tbabs', 'wabs')
_SIG = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF


class _File:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError('not an HDF5 file (signature at offset 0 expected)')
        ver = buf[8]
        if ver > 1:
            raise NotImplementedError(f'HDF5 superblock version {ver} (only 0 and 1: h5py libver="earliest")')
        if buf[13] != 8 or buf[14] != 8:
            raise NotImplementedError('HDF5 offsets / lengths other than 8 bytes')
        p = 24 + (4 if ver == 1 else 0)  # v1 adds indexed-storage K + 2 reserved bytes
        self.base, _free, _eof, _drv = struct.unpack_from('<4Q', buf, p)
        p += 32
        # root group symbol-table entry
        _name, self.root_header, cache, _res = struct.unpack_from('<QQII', buf, p)
        self.root_cache = struct.unpack_from('<QQ', buf, p + 24) if cache == 1 else None

    def at(self, addr: int) -> int:
        if addr == _UNDEF:
            raise ValueError('undefined address in the file')
        return self.base + addr

    # ---- object headers -------------------------------------------------------------------
    def messages(self, addr: int):
        """(type, data) of every message of the version-1 object header at ``addr``."""
        p = self.at(addr)
        if self.b[p:p + 4] == b'OHDR':
            raise NotImplementedError('version-2 object headers (file written with libver="latest")')
        ver, _r, n_msg, _refs, size = struct.unpack_from('<BBHII', self.b, p)
        if ver != 1:
            raise NotImplementedError(f'object header version {ver}')
        blocks = [(p + 16, size)]  # messages start 8-aligned after the 12-byte prefix
        out = []
        while blocks and len(out) < n_msg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < n_msg:
                mtype, msize, _flags = struct.unpack_from('<HHB', self.b, q)
                data = self.b[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x10:  # continuation: offset, length
                    off, length = struct.unpack_from('<QQ', data, 0)
                    blocks.append((self.at(off), length))
                out.append((mtype, data))
        return out

    # ---- groups ---------------------------------------------------------------------------
    def _heap_data(self, addr: int) -> int:
        p = self.at(addr)
        if self.b[p:p + 4] != b'HEAP':
            raise ValueError('local heap signature expected')
        _size, _free, data = struct.unpack_from('<QQQ', self.b, p + 8)
        return self.at(data)

    def _name(self, heap_data: int, off: int) -> str:
        end = self.b.index(b'\0', heap_data + off)
        return self.b[heap_data + off:end].decode()

    def _walk_btree(self, addr: int, heap_data: int, out: dict):
        p = self.at(addr)
        sig = self.b[p:p + 4]
        if sig == b'SNOD':
            _ver, _r, n = struct.unpack_from('<BBH', self.b, p + 4)
            for i in range(n):
                q = p + 8 + 40 * i
                name_off, header = struct.unpack_from('<QQ', self.b, q)
                out[self._name(heap_data, name_off)] = header
            return
        if sig != b'TREE':
            raise ValueError('B-tree / symbol-table node signature expected')
        ntype, _level, used = struct.unpack_from('<BBH', self.b, p + 4)
        if ntype != 0:
            raise ValueError('group B-tree node expected')
        q = p + 24  # after the two sibling addresses
        for i in range(used):  # key_i, child_i, ..., key_used
            child = struct.unpack_from('<Q', self.b, q + 8 + 16 * i)[0]
            self._walk_btree(child, heap_data, out)

    def group(self, btree: int, heap: int) -> dict:
        out: dict = {}
        self._walk_btree(btree, self._heap_data(heap), out)
        return out

    # ---- datasets -------------------------------------------------------------------------
    @staticmethod
    def _dtype(data: bytes) -> np.dtype:
        cls, ver = data[0] & 0x0F, data[0] >> 4
        bits0 = data[1]
        size = struct.unpack_from('<I', data, 4)[0]
        order = '>' if bits0 & 1 else '<'
        if ver not in (1, 2, 3):
            raise NotImplementedError(f'datatype message version {ver}')
        if cls == 1:  # floating point; IEEE layouts only
            if size not in (2, 4, 8):
                raise NotImplementedError(f'{size}-byte floating point')
            return np.dtype(f'{order}f{size}')
        if cls == 0:  # fixed point
            signed = bool(bits0 & 0x08)
            return np.dtype(f'{order}{"i" if signed else "u"}{size}')
        raise NotImplementedError(f'HDF5 datatype class {cls}')

    @staticmethod
    def _shape(data: bytes) -> tuple:
        ver, rank, flags = data[0], data[1], data[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if data[3] == 2:
                raise NotImplementedError('null dataspace')
            p = 4
        else:
            raise NotImplementedError(f'dataspace message version {ver}')
        del flags
        return struct.unpack_from(f'<{rank}Q', data, p) if rank else ()

    def dataset(self, msgs) -> np.ndarray:
        shape = dtype = raw = None
        for mtype, data in msgs:
            if mtype == 0x01:
                shape = self._shape(data)
            elif mtype == 0x03:
                dtype = self._dtype(data)
            elif mtype == 0x0B:
                raise NotImplementedError('filtered (compressed) dataset')
            elif mtype == 0x08:
                ver, cls = data[0], data[1]
                if ver != 3:
                    raise NotImplementedError(f'data layout message version {ver}')
                if cls == 1:  # contiguous
                    addr, size = struct.unpack_from('<QQ', data, 2)
                    raw = b'' if addr == _UNDEF else self.b[self.at(addr):self.at(addr) + size]
                elif cls == 0:  # compact
                    size = struct.unpack_from('<H', data, 2)[0]
                    raw = data[4:4 + size]
                else:
                    raise NotImplementedError('chunked dataset')
        if shape is None or dtype is None or raw is None:
            raise ValueError('dataset without dataspace / datatype / layout message')
        n = int(np.prod(shape)) if shape else 1
        if len(raw) < n * dtype.itemsize:
            raise ValueError('dataset storage shorter than its dataspace')
        return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).astype(dtype.newbyteorder('='))

    def read(self, header: int, prefix: str, out: dict):
        msgs = self.messages(header)
        for mtype, data in msgs:
            if mtype == 0x11:  # symbol table: a group
                btree, heap = struct.unpack_from('<QQ', data, 0)
                for name, child in sorted(self.group(btree, heap).items()):
                    self.read(child, f'{prefix}{name}/', out)
                return
            if mtype in (0x02, 0x06):
                raise NotImplementedError('new-style (link message) groups')
        out[prefix.rstrip('/')] = self.dataset(msgs)


def read_hdf5(path) -> dict[str, np.ndarray]:
    """Every dataset of a (small, classic-format) HDF5 file: {'group/sub/name': ndarray}."""
    with open(path, 'rb') as f:
        hf = _File(f.read())
    out: dict[str, np.ndarray] = {}
    hf.read(hf.root_header, '', out)
    return out


def load_xsect(path, dtype=None) -> dict:
    """The reference's table file as nested dicts, the shape of `_XSECT_INTERP`'s source
    (models/mul.py:286-299): ``{'energy': e, 'phabs': {xsect: {abund: sigma}}, 'tbabs': ..,
    'wabs': ..}``.  ``dtype=None`` keeps the stored dtype (float32): the reference takes
    ``np.log`` of the arrays as stored, and the components here do the same."""
    flat = read_hdf5(path)
    if dtype is not None:
        flat = {k: v.astype(dtype) for k, v in flat.items()}
    if 'energy' not in flat:
        raise ValueError(f"{path}: no 'energy' dataset")
    out: dict = {'energy': flat['energy']}
    for key, arr in flat.items():
        parts = key.split('/')
        if len(parts) != 3:
            continue
        model, xsect, abund = parts
        if arr.shape != flat['energy'].shape:
            raise ValueError(f'{key}: shape {arr.shape} differs from the energy grid')
        out.setdefault(model, {}).setdefault(xsect, {})[abund] = arr
    return out
