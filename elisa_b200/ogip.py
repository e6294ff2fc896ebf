"""OGIP response / spectrum ingest: PHA, RMF (+ARF) files ->
``FixedData``, the step before the likelihood path.

Mirrors, without astropy, what the reference does between the files and ``get_fixed_data()``:

* ``read_response``  -- ``Response._read_response`` / ``_read_arf`` (data/ogip.py:592-745): the
  compressed ``MATRIX`` column expanded group by group from ``N_GRP / F_CHAN / N_CHAN``
  (``TLMIN`` of ``F_CHAN`` = first channel number), multiplied by the ARF's ``SPECRESP``;
* ``read_spectrum``  -- ``Spectrum.__init__`` (data/ogip.py:267-500): ``COUNTS`` or ``RATE`` x
  exposure, Poisson / ``STAT_ERR`` (+ ``SYS_ERR``) errors, ``QUALITY``, ``GROUPING``,
  ``AREASCAL`` / ``BACKSCAL`` as column or keyword, type-II row selection ``file{N}``;
* ``load_data``      -- ``Data.__init__`` + ``ObservationData.__init__ / set_grouping``
  (data/ogip.py:147-247, data/base.py:58-139, 258-369): quality mask (1, 2, 5 bad; 1 only with
  ``ignore_bad=False``), the file's grouping flags applied to counts (sum), errors (quadrature),
  scales (XSPEC rule, data/grouping.py:314-431: count-weighted harmonic mean, NET spectra the
  arithmetic mean) and to the response as a sparse right-multiply (data/base.py:1528-1538),
  channel groups kept when they lie inside an ``erange`` interval and hold a good channel, the
  background ratio and the net counts.

The adaptive re-grouping methods of the reference (``group='min' | 'opt' | ...``,
data/grouping.py) are not part of this step: group with the file's ``GROUPING`` column (e.g.
written by ``ftgrouppha``) or pass ``grouping=`` flags.
"""

from __future__ import annotations

import os
import re
from dataclasses import dataclass

import numpy as np

from .data import FixedData
from .fitsio import HDU, read_fits

__all__ = ['Spectrum', 'Response', 'read_spectrum', 'read_response', 'load_data']

_UNIT_TO_KEV = {'ev': 1e-3, 'kev': 1.0, 'mev': 1e3, 'gev': 1e6}


def _kev(unit) -> float:
    try:
        return _UNIT_TO_KEV[str(unit).strip().lower()]
    except KeyError:
        raise NotImplementedError(f'energy unit {unit!r}') from None


def _find(hdus: list[HDU], name: str, ver: int | None = None) -> HDU | None:
    for h in hdus:
        if h.name == name and (ver is None or h.ver == ver):
            return h
    return None


def _split_row(path: str):
    m = re.compile(r'(.+){(.+)}').match(path)
    return (m.group(1), int(m.group(2))) if m else (path, None)


# --------------------------------------------------------------------------- response
@dataclass
class Response:
    photon_egrid: np.ndarray   # (E + 1,)
    channel_emin: np.ndarray   # (C,)
    channel_emax: np.ndarray
    channel: tuple             # channel labels, first channel = TLMIN of F_CHAN (default 1)
    channel_type: str
    rows: np.ndarray           # COO of the (E, C) matrix, explicit zeros removed, ARF applied
    cols: np.ndarray
    values: np.ndarray

    @property
    def shape(self):
        return self.photon_egrid.size - 1, len(self.channel)

    def dense(self) -> np.ndarray:
        out = np.zeros(self.shape)
        np.add.at(out, (self.rows, self.cols), self.values)
        return out


def read_response(respfile: str, ancrfile: str | None = None) -> Response:
    file, resp_id = _split_row(respfile)
    resp_id = 1 if resp_id is None else resp_id
    hdus = read_fits(file)
    ext = _find(hdus, 'MATRIX', resp_id) or _find(hdus, 'SPECRESP MATRIX', resp_id)
    if ext is None:  # the first extension with a MATRIX column
        cands = [h for h in hdus if 'MATRIX' in h.names]
        if not cands:
            raise ValueError(f'cannot read response matrix data from {respfile}')
        ext = _find(cands, cands[0].name, resp_id) or cands[0]
    eb = _find(hdus, 'EBOUNDS')
    if eb is None:
        raise ValueError(f'no EBOUNDS extension in {respfile}')
    h = ext.header
    channel_type = str(eb.header.get('CHANTYPE', h.get('CHANTYPE', 'PI')))

    def unit(hdu, col):
        return _kev(hdu.header.get(f'TUNIT{hdu.names.index(col) + 1}', 'keV'))

    # unit conversion in the STORED dtype, as `column * Unit(u).to('keV')` does in the reference
    # (float32 columns stay float32 under NumPy's weak-scalar promotion), float64 afterwards
    channel_emin = (np.asarray(eb['E_MIN']) * unit(eb, 'E_MIN')).astype(np.float64)
    channel_emax = (np.asarray(eb['E_MAX']) * unit(eb, 'E_MAX')).astype(np.float64)
    photon_emin = np.asarray(ext['ENERG_LO']) * unit(ext, 'ENERG_LO')
    photon_emax = np.asarray(ext['ENERG_HI']) * unit(ext, 'ENERG_HI')
    photon_egrid = np.ascontiguousarray(np.append(photon_emin, photon_emax[-1]), dtype=np.float64)
    if photon_egrid[0] <= 0.0:  # data/ogip.py:649-655
        photon_egrid[0] = max(1e-30, 0.001 * photon_egrid[1])

    if h.get('DETCHANS') is None:
        raise ValueError(f'keyword "DETCHANS" is not found in "{respfile}" header')
    n_chan_total = int(h['DETCHANS'])
    first_chan = int(h.get(f"TLMIN{ext.names.index('F_CHAN') + 1}", 1))
    channel = tuple(str(c) for c in range(first_chan, first_chan + n_chan_total))

    n_grp, f_chan, n_chan, matrix = ext['N_GRP'], ext['F_CHAN'], ext['N_CHAN'], ext['MATRIX']
    rows, cols, vals = [], [], []
    for i in range(len(ext)):
        n = int(n_grp[i])
        if n <= 0:
            continue
        f = np.asarray(np.atleast_1d(f_chan[i]), dtype=np.int64).ravel()[:n] - first_chan
        w = np.asarray(np.atleast_1d(n_chan[i]), dtype=np.int64).ravel()[:n]
        if f.size != n or w.size != n:
            raise ValueError(f'cannot parse grouped channel ranges in "{respfile}"')
        n_elem = int(w.sum())
        if n_elem <= 0:
            continue
        v = np.asarray(np.atleast_1d(matrix[i])).ravel()
        if v.size < n_elem:
            raise ValueError(f'cannot parse matrix values in "{respfile}" row {i + 1}')
        # run-length expansion of the channel ranges: start_g, start_g + 1, ..., start_g + w_g - 1
        starts = np.repeat(f - np.concatenate(([0], np.cumsum(w)[:-1])), w)
        cols.append(starts + np.arange(n_elem))
        rows.append(np.full(n_elem, i, dtype=np.int64))
        vals.append(v[:n_elem])
    rows = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    vals = np.concatenate(vals) if vals else np.zeros(0)
    if cols.size and (cols.min() < 0 or cols.max() >= n_chan_total):
        raise ValueError(f'channel range outside DETCHANS in "{respfile}"')

    if ancrfile:
        arf = _find(read_fits(ancrfile), 'SPECRESP')
        if arf is None:
            raise ValueError(f'no SPECRESP extension in {ancrfile}')
        # the reference multiplies the matrix by the ARF in the dtypes the files store (two float32
        # columns give a float32 product: data/ogip.py:581); keep that rounding
        area = np.asarray(arf['SPECRESP'])
        if area.size != len(ext):
            raise ValueError(f'rmf ({respfile}) and arf ({ancrfile}) are not matched')
        keep = vals != 0.0  # coo_array.eliminate_zeros() happens before the ARF is applied
        rows, cols, vals = rows[keep], cols[keep], vals[keep]
        vals = vals * area[rows]
    else:
        keep = vals != 0.0
        rows, cols, vals = rows[keep], cols[keep], vals[keep]
    return Response(photon_egrid, channel_emin, channel_emax, channel, channel_type, rows, cols,
                    vals.astype(np.float64))


# --------------------------------------------------------------------------- spectrum
@dataclass
class Spectrum:
    counts: np.ndarray
    errors: np.ndarray
    poisson: bool
    exposure: float
    quality: np.ndarray
    grouping: np.ndarray
    area_scale: np.ndarray
    back_scale: np.ndarray
    net: bool | None
    name: str
    backfile: str
    respfile: str
    ancrfile: str


def read_spectrum(specfile: str, poisson: bool | None = None) -> Spectrum:
    file, row = _split_row(specfile)
    hdu = _find(read_fits(file), 'SPECTRUM')
    if hdu is None:
        raise ValueError(f'no SPECTRUM extension in {specfile}')
    h, names = hdu.header, hdu.names
    type_ii = row is not None
    if not type_ii:
        msg = f"row id must be provided for type II spectrum, i.e., '{specfile}{{N}}'"
        if int(h.get('DETCHANS', len(hdu))) != len(hdu) or h.get('HDUCLAS4', '') == 'TYPE:II':
            raise ValueError(msg)
    if not any(k in names for k in ('COUNTS', 'RATE', 'RATES')):
        raise ValueError(f'"COUNTS", "RATE" or "RATES" not found in {specfile}')
    poisson = h.get('POISSERR', poisson)
    if poisson is None:
        raise RuntimeError(f'"POISSERR" is undefined in header of {specfile}: pass poisson=True/False')
    poisson = bool(poisson)
    if not poisson and 'STAT_ERR' not in names:
        raise ValueError(f'"STAT_ERR" not found in {specfile}')

    def get_field(name, default=None, excluded=None):
        if name in names:
            v = hdu[name]
            if type_ii:
                v = v[row - 1]
        else:
            v = h.get(name, default)
        if excluded is not None and not isinstance(v, np.ndarray) and v in excluded:
            return default
        return v

    exposure = float(get_field('EXPOSURE'))
    if 'COUNTS' in names:
        counts = np.array(get_field('COUNTS'), dtype=np.float64)
    else:
        counts = np.array(get_field('RATE' if 'RATE' in names else 'RATES'), dtype=np.float64) * exposure
    if poisson:
        stat_err = np.sqrt(counts)
    else:
        stat_err = np.array(get_field('STAT_ERR'), dtype=np.float64)
        if 'RATE' in names or 'RATES' in names:
            stat_err = stat_err * exposure
    n = counts.size

    def per_channel(name, default):  # a column, else a keyword, else the default
        v = get_field(name, default)
        return np.full(n, float(v)) if np.shape(v) == () else np.array(v, dtype=np.float64)

    sys_err = get_field('SYS_ERR', 0)
    sys_err = np.zeros(n) if np.shape(sys_err) == () else np.array(sys_err, dtype=np.float64)
    quality = get_field('QUALITY', 0)
    quality = np.zeros(n, dtype=np.int64) if np.shape(quality) == () else np.array(quality, dtype=np.int64)
    grouping = get_field('GROUPING', 0)
    grouping = np.ones(n, dtype=np.int64) if np.shape(grouping) == () else np.array(grouping, dtype=np.int64)
    if not poisson:
        if np.any(stat_err < 0.0):
            raise ValueError(f'spectrum {specfile} has negative statistical errors')
        if np.any(sys_err < 0.0):
            raise ValueError(f'spectrum {specfile} has systematic errors < 0')
        # Spectrum.__init__ folds SYS_ERR in (ogip.py:423-426) and hands sys_errors to
        # SpectrumData.__init__, which folds it in once more (base.py:1289-1292)
        errors = np.sqrt(np.square(stat_err) + np.square(sys_err * counts))
        errors = np.sqrt(np.square(errors) + np.square(sys_err * counts))
    else:
        errors = stat_err
    name = ''
    for key in ('DETNAM', 'INSTRUME', 'TELESCOP'):
        name = str(h.get(key, ''))
        if name.lower() not in ('', 'none', 'unknown'):
            break
        name = ''
    area_scale = per_channel('AREASCAL', 1.0)
    back_scale = per_channel('BACKSCAL', 1.0)
    if not np.all(np.isin(quality, [0, 1, 2, 5, -1])):
        raise ValueError('quality must be 0, 1, 2, 5, or -1')
    if poisson and np.any(counts < 0):
        raise ValueError('Poisson counts must be non-negative')
    if exposure <= 0:
        raise ValueError('exposure must be positive')
    for s in (area_scale, back_scale):  # base.py:1248-1280
        s[(s == 0.0) & (quality != 1)] = 1.0
        if np.any(s < 0):
            raise ValueError('area_scale / back_scale must be non-negative')
    if grouping[0] == -1:
        grouping[0] = 1
    hc2 = str(h.get('HDUCLAS2', '')).strip().upper()
    net = True if hc2 == 'NET' else (False if hc2 in ('TOTAL', 'BKG') else None)
    none = ('none', 'None', 'NONE')
    return Spectrum(counts, errors, poisson, exposure, quality, grouping, area_scale, back_scale, net, name,
                    get_field('BACKFILE', '', none) or '', get_field('RESPFILE', '', none) or '',
                    get_field('ANCRFILE', '', none) or '')


# --------------------------------------------------------------------------- grouping
def _seg_sum(x, group_idx):
    return np.add.reduceat(x, group_idx)


def _group_scale(counts, scale, group_idx, factor, net: bool):
    """data/grouping.py:314-431."""
    inv = np.divide(counts, scale, out=np.zeros_like(counts), where=scale != 0.0)
    rec = np.divide(1.0, scale, out=np.zeros_like(counts), where=scale != 0.0)
    g_counts = _seg_sum(factor * counts, group_idx)
    g_n = _seg_sum(factor * np.ones_like(counts), group_idx)
    g_inv = _seg_sum(factor * inv, group_idx)
    g_rec = _seg_sum(factor * rec, group_idx)
    g_sum = _seg_sum(factor * scale, group_idx)
    out = np.ones_like(g_counts)
    if net:
        np.divide(g_sum, g_n, out=out, where=g_n != 0.0)
        return out
    mask = g_inv != 0.0
    np.divide(g_counts, g_inv, out=out, where=mask)
    harmonic = np.ones_like(g_counts)
    np.divide(g_n, g_rec, out=harmonic, where=g_rec != 0.0)
    out[~mask] = harmonic[~mask]
    return out


def load_data(erange, specfile: str, backfile: str | None = None, respfile: str | None = None,
              ancrfile: str | None = None, name: str | None = None, spec_poisson: bool | None = None,
              back_poisson: bool | None = None, ignore_bad: bool = True, grouping=None,
              sparse_response: bool = False) -> FixedData:
    """``elisa.data.ogip.Data(erange, specfile, ...).get_fixed_data()`` for the fields the
    likelihood path reads.  File names found in the spectrum header are resolved as the
    reference does (as given, i.e. relative to the working directory), falling back to the
    spectrum's own directory."""
    spec = read_spectrum(specfile, spec_poisson)
    name = str(name) if name else spec.name
    if not name:
        raise ValueError(f"name must be set manually for {specfile} data, i.e., load_data(..., name='NAME')")
    here = os.path.dirname(_split_row(specfile)[0])

    def resolve(path):
        f, row = _split_row(path)
        if not os.path.exists(f) and os.path.exists(os.path.join(here, f)):
            f = os.path.join(here, f)
        return f if row is None else f'{f}{{{row}}}'

    ancrfile = ancrfile or spec.ancrfile
    respfile = respfile or spec.respfile
    if not respfile:
        raise ValueError(f'response file must be set manually for {specfile} data')
    resp = read_response(resolve(respfile), resolve(ancrfile) if ancrfile else None)
    E, C = resp.shape
    if spec.counts.size != C:
        raise ValueError(f'specfile ({specfile}) and respfile ({respfile}) are not matched')
    backfile = backfile or spec.backfile
    back = read_spectrum(resolve(backfile), back_poisson) if backfile else None
    if back is not None and back.counts.size != C:
        raise ValueError(f'specfile ({specfile}) and backfile ({backfile}) are not matched')

    # ObservationData.__init__, data/base.py:58-139
    bad = (1, 2, 5) if ignore_bad else (1,)
    good = ~np.isin(spec.quality, bad)
    if back is not None:
        good &= ~np.isin(back.quality, bad)
    if not np.any(good):
        raise RuntimeError(f'no good channel is found for {name} data')
    er = np.array(erange, dtype=np.float64, ndmin=2)
    if np.any(np.diff(er, axis=1) <= 0.0):
        raise ValueError('erange must be increasing')
    er = er[er[:, 0].argsort()]
    if np.any(np.diff(np.hstack(er)) <= 0.0):
        raise ValueError('erange must not be overlapped')

    # set_grouping, data/base.py:258-369
    grouping = np.array(spec.grouping if grouping is None else grouping, dtype=np.int64)
    if grouping.shape != (C,):
        raise ValueError('grouping must have the same size as the number of channels')
    if grouping[0] == -1:
        grouping[0] = 1
    gidx = np.flatnonzero(grouping != -1)
    factor = good.astype(np.float64)

    def group_spectrum(s: Spectrum):  # SpectrumData.group, data/base.py:1305-1343
        return _seg_sum(factor * s.counts, gidx), np.sqrt(_seg_sum(factor * s.errors * s.errors, gidx))

    spec_counts, spec_errors = group_spectrum(spec)
    source_net = bool(spec.net) and back is None  # _group_scale_arrays, data/base.py:141-190
    spec_area = _group_scale(spec.counts, spec.area_scale, gidx, factor, source_net)
    spec_back = _group_scale(spec.counts, spec.back_scale, gidx, factor, source_net)

    # ResponseData.group / group_energy, data/base.py:1492-1608: R . G with G[c, g] = good[c]
    edges = np.append(gidx, C)
    if gidx.size != C:
        g_emin = np.minimum.reduceat(resp.channel_emin, gidx)
        g_emax = np.maximum.reduceat(resp.channel_emax, gidx)
        group_of = np.repeat(np.arange(gidx.size), np.diff(edges))
        w = factor[resp.cols]
        keep = w != 0.0
        rows, cols, vals = resp.rows[keep], group_of[resp.cols[keep]], resp.values[keep]
    else:
        g_emin, g_emax = resp.channel_emin, resp.channel_emax
        rows, cols, vals = resp.rows, resp.cols, resp.values
    mask = np.any((er[:, :1] <= g_emin) & (g_emax <= er[:, 1:]), axis=0)  # _get_channel_mask
    mask &= _seg_sum(good.astype(np.int64), gidx) > 0
    new_index = np.cumsum(mask) - 1
    sel = mask[cols]
    rows, cols, vals = rows[sel], new_index[cols[sel]], vals[sel]
    Cg = int(mask.sum())
    if Cg == 0:
        raise RuntimeError(f'no channel of {name} data lies inside erange')

    kw = dict(name=name, photon_egrid=resp.photon_egrid, spec_counts=spec_counts[mask], area_scale=spec_area[mask],
              spec_exposure=spec.exposure, channel_width=(g_emax - g_emin)[mask])
    if back is not None:
        back_counts, back_errors = group_spectrum(back)
        back_area = _group_scale(back.counts, back.area_scale, gidx, factor, False)
        back_back = _group_scale(back.counts, back.back_scale, gidx, factor, False)
        ratio = spec.exposure * spec_area * spec_back / (back.exposure * back_area * back_back)  # _group_back_ratio
        ratio = ratio[mask]
        kw.update(back_counts=back_counts[mask], back_errors=back_errors[mask], back_ratio=ratio,
                  net_counts=spec_counts[mask] - ratio * back_counts[mask],
                  net_errors=np.sqrt(np.square(spec_errors[mask]) + np.square(ratio * back_errors[mask])))
    else:
        kw.update(net_counts=spec_counts[mask], net_errors=spec_errors[mask])

    if sparse_response:  # CSR by channel row of R.T (what BCSR.from_scipy_sparse(sparse_matrix.T) holds)
        order = np.lexsort((rows, cols))
        rows, cols, vals = rows[order], cols[order], vals[order]
        # duplicates (several channels of one group hitting the same energy) are summed
        key = cols * np.int64(E) + rows
        first = np.concatenate(([True], key[1:] != key[:-1])) if key.size else np.zeros(0, dtype=bool)
        starts = np.flatnonzero(first)
        vals = np.add.reduceat(vals, starts) if key.size else vals
        rows, cols = rows[first], cols[first]
        indptr = np.zeros(Cg + 1, dtype=np.int32)
        np.add.at(indptr, cols + 1, 1)
        kw.update(sparse_matrix=(np.cumsum(indptr).astype(np.int32), rows.astype(np.int32), vals), response_sparse=True)
    else:
        dense = np.zeros((E, Cg))
        np.add.at(dense, (rows, cols), vals)
        kw.update(response_matrix=dense)
    d = FixedData(**kw)
    d.spec_errors = spec_errors[mask]
    d.channel_emin, d.channel_emax = g_emin[mask], g_emax[mask]
    d.spec_poisson, d.back_poisson = spec.poisson, (back.poisson if back is not None else None)
    d.back_exposure = back.exposure if back is not None else None
    return d
