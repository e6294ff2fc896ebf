"""Synthetic OGIP files (PHA / RMF / ARF) shared by tools/make_golden.py --ogip, which feeds
them to the reference's own ``elisa.data.ogip.Data`` (executed unmodified), and the ingest tests.
The FITS writer below is test infrastructure: big-endian binary tables with fixed-width and
variable-length ('P') columns, just enough to express the layouts found in the wild (HXMT:
fixed-width MATRIX rows, F_CHAN from 0; XMM/Swift-like: variable-length groups)."""

import os

import numpy as np

_BLOCK = 2880


def _card(key, value, quote=None):
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f'{int(value):>20}'
    elif isinstance(value, (float, np.floating)):
        v = f'{float(value):>20.12E}'
    else:
        s = str(value).replace("'", "''")
        v = f"'{s:<8}'"
    return f'{key:<8}= {v}'.ljust(80).encode('ascii')


def _header(cards):
    raw = b''.join(_card(k, v) for k, v in cards) + b'END'.ljust(80)
    return raw + b' ' * (-len(raw) % _BLOCK)


_CODES = {'>i2': 'I', '>i4': 'J', '>i8': 'K', '>f4': 'E', '>f8': 'D'}


def bintable(name, columns, keywords=(), ver=1):
    """columns: list of (name, data[, unit]); data = ndarray (rows,) / (rows, repeat) with a
    big-endian-able dtype, or a list of 1-d arrays (variable length, all of one dtype)."""
    n_rows = len(columns[0][1])
    fields, heap, cards = [], b'', []
    for k, col in enumerate(columns, 1):
        cname, data = col[0], col[1]
        unit = col[2] if len(col) > 2 else None
        if isinstance(data, list):
            dt = np.dtype(data[0].dtype).newbyteorder('>')
            desc = np.zeros((n_rows, 2), dtype='>i4')
            for i, a in enumerate(data):
                desc[i] = (len(a), len(heap))
                heap += np.asarray(a, dtype=dt).tobytes()
            fields.append(desc)
            form = f'1P{_CODES[dt.str]}({max(len(a) for a in data)})'
        else:
            a = np.asarray(data)
            dt = a.dtype.newbyteorder('>')
            a2 = a.reshape(n_rows, -1).astype(dt)
            fields.append(a2)
            form = f'{a2.shape[1]}{_CODES[dt.str]}'
        cards += [(f'TTYPE{k}', cname), (f'TFORM{k}', form)]
        if unit:
            cards.append((f'TUNIT{k}', unit))
    rows = [b''.join(np.ascontiguousarray(f[i]).tobytes() for f in fields) for i in range(n_rows)]
    row_bytes = len(rows[0]) if rows else 0
    data = b''.join(rows) + heap
    head = [('XTENSION', 'BINTABLE'), ('BITPIX', 8), ('NAXIS', 2), ('NAXIS1', row_bytes), ('NAXIS2', n_rows),
            ('PCOUNT', len(heap)), ('GCOUNT', 1), ('TFIELDS', len(columns))] + cards + \
           [('EXTNAME', name), ('EXTVER', ver)] + list(keywords)
    return _header(head) + data + b'\0' * (-len(data) % _BLOCK)


def write_fits(path, tables):
    primary = _header([('SIMPLE', True), ('BITPIX', 8), ('NAXIS', 0), ('EXTEND', True)])
    with open(path, 'wb') as f:
        f.write(primary + b''.join(tables))


# --------------------------------------------------------------------------- the cases
def _rmf_variable(path, rng, E=60, C=48, first_chan=0, arf_path=None):
    """Variable-length groups, up to 3 per energy, some energies without any group."""
    edges = np.geomspace(0.3, 12.0, E + 1).astype(np.float32)
    ch_edges = np.linspace(0.2, 12.5, C + 1).astype(np.float32)
    n_grp, f_chan, n_chan, matrix = [], [], [], []
    for i in range(E):
        if i in (0, 7):  # no response at all
            n_grp.append(0); f_chan.append(np.zeros(0, '>i2')); n_chan.append(np.zeros(0, '>i2'))
            matrix.append(np.zeros(0, '>f4'))
            continue
        centre = int(np.interp(0.5 * (edges[i] + edges[i + 1]), ch_edges, np.arange(C + 1)))
        starts, widths = [], []
        lo = max(0, centre - 6)
        for g in range(int(rng.integers(1, 4))):
            w = int(rng.integers(1, 5))
            if lo + w > C:
                break
            starts.append(lo + first_chan); widths.append(w)
            lo += w + int(rng.integers(1, 3))
        n_grp.append(len(starts))
        f_chan.append(np.array(starts, dtype='>i2')); n_chan.append(np.array(widths, dtype='>i2'))
        vals = rng.random(int(np.sum(widths))).astype('>f4')
        vals[rng.random(vals.size) < 0.1] = 0.0  # explicit zeros inside a group
        matrix.append(vals)
    keys = [('DETCHANS', C), ('CHANTYPE', 'PI'), ('HDUCLAS1', 'RESPONSE'), ('HDUCLAS2', 'RSP_MATRIX')]
    if first_chan != 1:
        keys.append(('TLMIN4', first_chan))
    tables = [bintable('MATRIX', [('ENERG_LO', edges[:-1], 'keV'), ('ENERG_HI', edges[1:], 'keV'),
                                  ('N_GRP', np.array(n_grp, dtype='>i2')), ('F_CHAN', f_chan), ('N_CHAN', n_chan),
                                  ('MATRIX', matrix)], keys),
              bintable('EBOUNDS', [('CHANNEL', np.arange(first_chan, first_chan + C, dtype='>i2')),
                                   ('E_MIN', ch_edges[:-1], 'keV'), ('E_MAX', ch_edges[1:], 'keV')],
                       [('DETCHANS', C), ('CHANTYPE', 'PI')])]
    write_fits(path, tables)
    if arf_path:
        area = (50.0 + 400.0 * rng.random(E)).astype('>f4')
        write_fits(arf_path, [bintable('SPECRESP', [('ENERG_LO', edges[:-1], 'keV'), ('ENERG_HI', edges[1:], 'keV'),
                                                    ('SPECRESP', area, 'cm**2')])])
    return E, C


def _rsp_fixed(path, rng, E=40, C=32, unit='eV'):
    """HXMT-like: one group spanning every channel, fixed-width MATRIX rows, energies in eV,
    F_CHAN counted from 1 (no TLMIN keyword), first photon edge 0."""
    scale = 1000.0 if unit == 'eV' else 1.0
    edges = (np.linspace(0.0, 20.0, E + 1) * scale).astype(np.float32)
    ch_edges = (np.linspace(0.5, 21.0, C + 1) * scale).astype(np.float32)
    mat = np.zeros((E, C), dtype='>f4')
    for i in range(E):
        c = int(i * C / E)
        mat[i, max(0, c - 2):c + 3] = 20.0 * rng.random(min(C, c + 3) - max(0, c - 2))
    tables = [bintable('SPECRESP MATRIX', [('ENERG_LO', edges[:-1], unit), ('ENERG_HI', edges[1:], unit),
                                           ('N_GRP', np.ones(E, dtype='>i2')), ('F_CHAN', np.ones(E, dtype='>i2')),
                                           ('N_CHAN', np.full(E, C, dtype='>i2')), ('MATRIX', mat)],
                       [('DETCHANS', C)]),
              bintable('EBOUNDS', [('CHANNEL', np.arange(1, C + 1, dtype='>i2')), ('E_MIN', ch_edges[:-1], unit),
                                   ('E_MAX', ch_edges[1:], unit)], [('DETCHANS', C), ('CHANTYPE', 'PHA')])]
    write_fits(path, tables)
    return E, C


def _grouping(rng, C):
    g = np.ones(C, dtype='>i2')
    i = 0
    while i < C:
        w = int(rng.integers(1, 5))
        g[i + 1:i + w] = -1
        i += w
    return g


def build(dirname):
    """Write every case's files under dirname; -> {case: kwargs of Data / load_data}."""
    os.makedirs(dirname, exist_ok=True)
    p = lambda f: os.path.join(dirname, f)  # noqa: E731
    rng = np.random.default_rng(20261020)
    cases = {}

    # A: variable-length RMF (channels from 0) x ARF, Poisson source + Poisson background with
    # quality flags, grouping flags, per-channel AREASCAL / BACKSCAL columns, two erange windows
    E, C = _rmf_variable(p('a.rmf'), rng, first_chan=0, arf_path=p('a.arf'))
    qual = np.zeros(C, dtype='>i2'); qual[[3, 4]] = 1; qual[10] = 2; qual[20] = 5; qual[33] = -1
    bqual = np.zeros(C, dtype='>i2'); bqual[[3, 11]] = 1
    grp = _grouping(rng, C)
    area = (0.8 + 0.4 * rng.random(C)).astype('>f4'); area[5] = 0.0  # zero scale in a used channel -> 1
    write_fits(p('a_src.pha'), [bintable('SPECTRUM', [
        ('CHANNEL', np.arange(C, dtype='>i2')), ('COUNTS', rng.poisson(30.0, C).astype('>i4')), ('QUALITY', qual),
        ('GROUPING', grp), ('AREASCAL', area), ('BACKSCAL', (0.02 + 0.01 * rng.random(C)).astype('>f4'))],
        [('DETCHANS', C), ('EXPOSURE', 1234.5), ('POISSERR', True), ('HDUCLAS2', 'TOTAL'), ('INSTRUME', 'CAMA'),
         ('TELESCOP', 'SAT'), ('BACKFILE', 'a_bkg.pha'), ('RESPFILE', 'a.rmf'), ('ANCRFILE', 'a.arf')])])
    write_fits(p('a_bkg.pha'), [bintable('SPECTRUM', [
        ('CHANNEL', np.arange(C, dtype='>i2')), ('COUNTS', rng.poisson(8.0, C).astype('>i4')), ('QUALITY', bqual),
        ('GROUPING', np.ones(C, dtype='>i2'))],
        [('DETCHANS', C), ('EXPOSURE', 4000.0), ('POISSERR', True), ('HDUCLAS2', 'BKG'), ('AREASCAL', 1.0),
         ('BACKSCAL', 0.11), ('BACKFILE', 'NONE'), ('RESPFILE', 'none')])])
    cases['a_varlen_arf_back_grouped'] = dict(erange=[(0.5, 3.0), (4.0, 11.0)], specfile=p('a_src.pha'),
                                              backfile=p('a_bkg.pha'), respfile=p('a.rmf'), ancrfile=p('a.arf'))
    cases['a_keep_dubious'] = dict(erange=[(0.3, 12.0)], specfile=p('a_src.pha'), backfile=p('a_bkg.pha'),
                                   respfile=p('a.rmf'), ancrfile=p('a.arf'), ignore_bad=False)
    cases['a_no_background_sparse'] = dict(erange=[(1.0, 9.0)], specfile=p('a_src.pha'), backfile=None,
                                           respfile=p('a.rmf'), ancrfile=None, sparse_response=True, _no_back=True)

    # B: fixed-width RSP in eV with a zero first edge, RATE + STAT_ERR + SYS_ERR spectrum marked NET
    E, C = _rsp_fixed(p('b.rsp'), rng)
    write_fits(p('b_net.pha'), [bintable('SPECTRUM', [
        ('CHANNEL', np.arange(1, C + 1, dtype='>i4')), ('RATE', (5.0 * rng.random(C)).astype('>f8')),
        ('STAT_ERR', (0.05 + 0.1 * rng.random(C)).astype('>f8')), ('SYS_ERR', np.full(C, 0.02, dtype='>f4')),
        ('GROUPING', _grouping(rng, C)), ('AREASCAL', (0.9 + 0.2 * rng.random(C)).astype('>f4'))],
        [('DETCHANS', C), ('EXPOSURE', 250.0), ('POISSERR', False), ('HDUCLAS2', 'NET'), ('DETNAM', 'none'),
         ('INSTRUME', 'LE'), ('BACKSCAL', 1.0), ('BACKFILE', 'NONE'), ('RESPFILE', 'NONE'), ('ANCRFILE', 'NONE')])])
    cases['b_fixed_rsp_rate_net'] = dict(erange=[(2.0, 18.0)], specfile=p('b_net.pha'), respfile=p('b.rsp'))

    # C: type II spectrum (two rows of vector columns), row {2}; Gaussian background given as RATE
    _rsp_fixed(p('c.rsp'), rng, E=30, C=24, unit='keV')
    C = 24
    write_fits(p('c_ii.pha'), [bintable('SPECTRUM', [
        ('SPEC_NUM', np.array([1, 2], dtype='>i2')), ('COUNTS', rng.poisson(40.0, (2, C)).astype('>i4')),
        ('QUALITY', np.zeros((2, C), dtype='>i2')), ('GROUPING', np.stack([np.ones(C, dtype='>i2'), _grouping(rng, C)])),
        ('EXPOSURE', np.array([100.0, 220.0], dtype='>f8')), ('BACKSCAL', np.array([1.0, 0.5], dtype='>f8'))],
        [('DETCHANS', C), ('POISSERR', True), ('HDUCLAS2', 'TOTAL'), ('HDUCLAS4', 'TYPE:II'), ('TELESCOP', 'OBS'),
         ('AREASCAL', 1.0)])])
    write_fits(p('c_bkg.pha'), [bintable('SPECTRUM', [
        ('CHANNEL', np.arange(1, C + 1, dtype='>i2')), ('RATE', (0.05 + 0.02 * rng.random(C)).astype('>f4')),
        ('STAT_ERR', (0.004 + 0.002 * rng.random(C)).astype('>f4'))],
        [('DETCHANS', C), ('EXPOSURE', 900.0), ('POISSERR', False), ('HDUCLAS2', 'BKG'), ('AREASCAL', 1.0),
         ('BACKSCAL', 2.0)])])
    cases['c_type2_row2_gaussian_back'] = dict(erange=[(1.0, 19.0)], specfile=p('c_ii.pha') + '{2}',
                                               backfile=p('c_bkg.pha'), respfile=p('c.rsp'))
    return cases


FIELDS = ('photon_egrid', 'spec_counts', 'spec_errors', 'area_scale', 'spec_exposure', 'back_counts', 'back_errors',
          'back_ratio', 'back_exposure', 'net_counts', 'net_errors', 'channel_emin', 'channel_emax', 'channel_width',
          'response_matrix')
