"""``FixedData``: the immutable per-spectrum inputs of the likelihood path, with the field
names of the reference's container (data/base.py:1893-1975, filled by
ObservationData.get_fixed_data :881-915).  Only the fields the path reads are kept."""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def _f64(x, shape=None, name=''):
    if x is None:
        return None
    a = np.ascontiguousarray(x, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f'{name}: expected shape {shape}, got {a.shape}')
    return a


@dataclass
class FixedData:
    """Field semantics follow data/base.py:1893-1975.

    response_matrix is (E, C) C-order like ``FixedData.response_matrix``; alternatively
    ``sparse_matrix`` = (indptr, indices, values) is the CSR-by-channel form of R.T, i.e.
    what ``BCSR.from_scipy_sparse(sparse_matrix.T)`` holds upstream
    (infer/likelihood.py:177-181), used when ``response_sparse``."""

    name: str
    photon_egrid: np.ndarray
    spec_counts: np.ndarray
    area_scale: np.ndarray
    spec_exposure: float
    channel_width: np.ndarray | None = None
    response_matrix: np.ndarray | None = None
    sparse_matrix: tuple | None = None
    response_sparse: bool = False
    net_counts: np.ndarray | None = None
    net_errors: np.ndarray | None = None
    back_counts: np.ndarray | None = None
    back_errors: np.ndarray | None = None
    back_ratio: np.ndarray | None = None
    has_back: bool = field(init=False, default=False)

    def __post_init__(self):
        self.photon_egrid = _f64(self.photon_egrid)
        if self.photon_egrid.ndim != 1 or self.photon_egrid.size < 2:
            raise ValueError('photon_egrid must hold at least two edges')
        E = self.photon_egrid.size - 1
        self.spec_counts = _f64(self.spec_counts)
        C = self.spec_counts.size
        shape = (C,)
        self.spec_counts = _f64(self.spec_counts, shape, 'spec_counts')
        self.area_scale = _f64(np.broadcast_to(np.asarray(self.area_scale, dtype=np.float64), shape), shape, 'area_scale')
        self.spec_exposure = float(self.spec_exposure)
        if self.channel_width is None:
            self.channel_width = np.ones(C)
        self.channel_width = _f64(self.channel_width, shape, 'channel_width')
        for k in ('net_counts', 'net_errors', 'back_counts', 'back_errors'):
            setattr(self, k, _f64(getattr(self, k), shape, k))
        if self.back_ratio is not None:
            self.back_ratio = _f64(np.broadcast_to(np.asarray(self.back_ratio, dtype=np.float64), shape), shape, 'back_ratio')
        self.has_back = self.back_counts is not None
        if self.response_sparse:
            if self.sparse_matrix is None:
                raise ValueError('response_sparse=True needs sparse_matrix')
            indptr, indices, values = self.sparse_matrix
            indptr = np.ascontiguousarray(indptr, dtype=np.int32)
            indices = np.ascontiguousarray(indices, dtype=np.int32)
            values = _f64(values)
            if indptr.shape != (C + 1,) or indices.shape != values.shape:
                raise ValueError('sparse_matrix: expected CSR by channel row (indptr of C + 1 entries)')
            self.sparse_matrix = (indptr, indices, values)
        else:
            if self.response_matrix is None:
                raise ValueError('response_matrix is required unless response_sparse')
            self.response_matrix = _f64(self.response_matrix, (E, C), 'response_matrix')

    @property
    def n_energies(self):
        return self.photon_egrid.size - 1

    @property
    def n_channels(self):
        return self.spec_counts.size

    def dense_response(self) -> np.ndarray:
        """(E, C) dense matrix whatever the storage (host-side helper for tests / oracle)."""
        if not self.response_sparse:
            return self.response_matrix
        indptr, indices, values = self.sparse_matrix
        out = np.zeros((self.n_energies, self.n_channels))
        for c in range(self.n_channels):
            sl = slice(indptr[c], indptr[c + 1])
            np.add.at(out[:, c], indices[sl], values[sl])
        return out

    def as_mapping(self) -> dict:
        """Plain dict with the FixedData field names (the test oracle's input format)."""
        return {
            'photon_egrid': self.photon_egrid, 'response_matrix': self.dense_response(),
            'area_scale': self.area_scale, 'spec_exposure': self.spec_exposure,
            'channel_width': self.channel_width, 'spec_counts': self.spec_counts,
            'net_counts': self.net_counts, 'net_errors': self.net_errors,
            'back_counts': self.back_counts, 'back_errors': self.back_errors, 'back_ratio': self.back_ratio,
        }
