"""Response / spectrum ingest against golden FixedData produced by
the reference's own `elisa.data.ogip.Data(...).get_fixed_data()` on the synthetic OGIP files of
tests/ogip_cases.py (tools/make_golden.py --ogip)."""

import os

import numpy as np
import pytest

import ogip_cases
from elisa_b200.fitsio import read_fits
from elisa_b200.ogip import load_data, read_response, read_spectrum

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'ogip.npz')
CASES = ['a_varlen_arf_back_grouped', 'a_keep_dubious', 'a_no_background_sparse', 'b_fixed_rsp_rate_net',
         'c_type2_row2_gaussian_back']


@pytest.fixture(scope='module')
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp('ogip')
    return str(d), ogip_cases.build(str(d))


@pytest.mark.parametrize('case', CASES)
def test_fixed_data_matches_reference(files, case, monkeypatch):
    tmp, cases = files
    monkeypatch.chdir(tmp)
    g = np.load(GOLD)
    kw = {k: v for k, v in cases[case].items() if not k.startswith('_')}
    d = load_data(**kw)
    checked = 0
    for f in ogip_cases.FIELDS:
        key = f'{case}/{f}'
        if key not in g.files:
            assert getattr(d, f, None) is None, f
            continue
        mine = d.dense_response() if f == 'response_matrix' else getattr(d, f)
        mine, ref = np.asarray(mine, dtype=np.float64), g[key]
        assert mine.shape == ref.shape, (f, mine.shape, ref.shape)
        assert np.allclose(mine, ref, rtol=1e-14, atol=0.0), (f, float(np.max(np.abs(mine - ref))))
        checked += 1
    assert checked >= 10
    assert np.allclose(d.dense_response(), g[f'{case}/sparse_dense'], rtol=1e-14, atol=0.0)
    assert d.response_sparse == bool(kw.get('sparse_response', False))


def test_file_names_from_the_header_and_errors(files, monkeypatch):
    tmp, cases = files
    monkeypatch.chdir('/')
    g = np.load(GOLD)
    # BACKFILE / RESPFILE / ANCRFILE keywords resolved next to the spectrum
    d = load_data([(0.5, 3.0), (4.0, 11.0)], os.path.join(tmp, 'a_src.pha'))
    assert np.array_equal(d.net_counts, g['a_varlen_arf_back_grouped/net_counts'])
    with pytest.raises(ValueError):  # type II needs a row
        read_spectrum(os.path.join(tmp, 'c_ii.pha'))
    with pytest.raises(ValueError):  # no RESPFILE in the header, none given
        load_data([(1.0, 5.0)], os.path.join(tmp, 'b_net.pha'))
    with pytest.raises(ValueError):
        load_data([(5.0, 1.0)], os.path.join(tmp, 'a_src.pha'))
    with pytest.raises(ValueError):  # channel numbers do not match
        load_data([(1.0, 5.0)], os.path.join(tmp, 'a_src.pha'), respfile=os.path.join(tmp, 'b.rsp'))
    with pytest.raises(RuntimeError):
        load_data([(100.0, 200.0)], os.path.join(tmp, 'a_src.pha'))


def test_response_decode_details(files):
    tmp, _ = files
    r = read_response(os.path.join(tmp, 'a.rmf'))
    E, C = r.shape
    assert (E, C) == (60, 48) and r.channel[0] == '0' and r.channel_type == 'PI'
    assert not np.any(r.rows == 0) and not np.any(r.rows == 7)      # energies without a group
    assert np.all(r.values != 0.0)                                   # explicit zeros dropped
    hdu = [h for h in read_fits(os.path.join(tmp, 'a.rmf')) if h.name == 'MATRIX'][0]
    assert isinstance(hdu['MATRIX'], list) and hdu['N_GRP'].dtype == np.int16
    b = read_response(os.path.join(tmp, 'b.rsp'))
    assert b.channel[0] == '1' and b.photon_egrid[0] == pytest.approx(0.001 * b.photon_egrid[1])  # zero edge lifted
    assert b.photon_egrid[-1] == pytest.approx(20.0, rel=1e-6) and b.channel_type == 'PHA'         # eV -> keV


HXMT = '/root/reference/docs/notebooks/data'


@pytest.mark.skipif(not os.path.isdir(HXMT), reason='the reference tree is only present in the build container')
def test_hxmt_sample_files():
    """The mission files the reference's notebooks use (Insight-HXMT LE / ME / HE)."""
    for det, C, E in (('HE', 256, 485), ('ME', 1024, 660)):
        d = load_data([(1.0, 300.0)], f'{HXMT}/P011160506306_{det}_TOTAL.fits', backfile=f'{HXMT}/P011160506306_{det}_BKG.fits',
                      respfile=f'{HXMT}/P011160506306_{det}_RSP.fits', name=det)
        assert d.n_energies == E and 0 < d.n_channels <= C
        R = d.dense_response()
        assert np.all(np.isfinite(R)) and R.sum() > 0 and R.min() > -1e-3 * R.max()  # mission RSPs carry small negatives
        assert np.all(np.isfinite(d.back_ratio)) and np.all(d.back_ratio > 0)
        assert np.all(d.spec_counts >= 0) and np.all(d.net_errors >= 0)


def test_grouped_scales_known_answers():
    """The reference's own known-answer tests of the XSPEC scale rules
    (tests/test_data.py:956-1041: count-weighted harmonic mean, arithmetic mean for NET spectra,
    plain harmonic mean when the weighted denominator vanishes, constant scales stay constant)."""
    from elisa_b200.ogip import _group_scale

    gidx, ones = np.array([0]), np.ones(2)

    def g(counts, scale, net):
        return _group_scale(np.asarray(counts, float), np.asarray(scale, float), gidx, ones, net)

    assert np.allclose(g([10.0, 20.0], [2.0, 6.0], False), [3.6])
    assert np.allclose(g([10.0, 20.0], [4.0, 12.0], False), [7.2])
    assert np.allclose(g([10.0, 20.0], [2.0, 6.0], True), [4.0])
    assert np.allclose(g([10.0, 20.0], [3.0, 9.0], True), [6.0])
    assert np.allclose(g([1.0, -3.0], [1.0, 3.0], False), [1.5])   # weighted denominator is zero
    assert np.allclose(g([1.0, -3.0], [2.0, 6.0], False), [3.0])
    # background ratio of a grouped pair (test_background_grouped_ratio_uses_non_net_formula)
    sa, sb = g([10.0, 20.0], [2.0, 6.0], False), g([10.0, 20.0], [3.0, 9.0], False)
    ba, bb = g([4.0, 8.0], [2.0, 6.0], False), g([4.0, 8.0], [2.0, 8.0], False)
    expected = (30.0 / (10.0 / 2.0 + 20.0 / 6.0)) * (30.0 / (10.0 / 3.0 + 20.0 / 9.0)) / (
        (12.0 / (4.0 / 2.0 + 8.0 / 6.0)) * (12.0 / (4.0 / 2.0 + 8.0 / 8.0)))
    assert np.allclose(sa * sb / (ba * bb), [expected])
    assert np.allclose(g([10.0, 20.0], [2.0, 2.0], False), [2.0]) and np.allclose(g([4.0, 8.0], [5.0, 5.0], True), [5.0])


def test_fits_reader_conventions(tmp_path):
    """Column conventions found in mission files beyond the synthetic OGIP cases: unsigned 16-bit
    integers (TZERO = 32768 on a signed column), scaled columns (TSCAL), 64-bit variable-length
    descriptors ('Q'), logical and string columns, and a header with CONTINUE / HISTORY cards."""
    import struct

    from ogip_cases import _header, _BLOCK

    n = 3
    chan = np.array([0, 40000, 65535])                     # stored as int16 with TZERO
    stored = (chan - 32768).astype('>i2')
    scaled = np.array([10, 20, 30], dtype='>i2')            # value = 0.5 * stored + 1
    flags = np.array([b'T', b'F', b'T'])
    names = [b'ab  ', b'c   ', b'    ']
    var = [np.array([1.5, 2.5], dtype='>f8'), np.zeros(0, dtype='>f8'), np.array([7.0], dtype='>f8')]
    heap = b''.join(v.tobytes() for v in var)
    rows, off = [], 0
    for i in range(n):
        desc = struct.pack('>qq', len(var[i]), off)
        off += var[i].nbytes
        rows.append(stored[i:i + 1].tobytes() + scaled[i:i + 1].tobytes() + flags[i] + names[i] + desc)
    row_bytes = len(rows[0])
    cards = [('XTENSION', 'BINTABLE'), ('BITPIX', 8), ('NAXIS', 2), ('NAXIS1', row_bytes), ('NAXIS2', n),
             ('PCOUNT', len(heap)), ('GCOUNT', 1), ('TFIELDS', 5),
             ('TTYPE1', 'CHANNEL'), ('TFORM1', '1I'), ('TZERO1', 32768), ('TSCAL1', 1),
             ('TTYPE2', 'SCALED'), ('TFORM2', 'I'), ('TSCAL2', 0.5), ('TZERO2', 1.0),
             ('TTYPE3', 'FLAG'), ('TFORM3', '1L'), ('TTYPE4', 'NAME'), ('TFORM4', '4A'),
             ('TTYPE5', 'VAR'), ('TFORM5', '1QD(2)'), ('EXTNAME', 'THINGS'), ('EXTVER', 2),
             ('LONGSTR', "it's"), ('EXPOSURE', 1.5e3)]
    hdr = _header(cards)
    # splice a HISTORY and a CONTINUE card in front of END (they carry no '= ' value indicator)
    body = hdr[:80 * len(cards)]
    body += b'HISTORY made by a test'.ljust(80) + b"CONTINUE  'ignored'".ljust(80) + b'END'.ljust(80)
    hdr = body + b' ' * (-len(body) % _BLOCK)
    data = b''.join(rows) + heap
    primary = _header([('SIMPLE', True), ('BITPIX', 16), ('NAXIS', 2), ('NAXIS1', 3), ('NAXIS2', 2), ('EXTEND', True)])
    image = np.arange(6, dtype='>i2').tobytes()              # a primary image to be skipped over
    p = tmp_path / 't.fits'
    p.write_bytes(primary + image + bytes(-len(image) % _BLOCK) + hdr + data + bytes(-len(data) % _BLOCK))
    hdus = read_fits(str(p))
    assert [h.name for h in hdus] == ['PRIMARY', 'THINGS'] and hdus[1].ver == 2
    h = hdus[1]
    assert h.header['LONGSTR'] == "it's" and h.header['EXPOSURE'] == 1500.0 and 'HISTORY' not in h.header
    assert np.array_equal(h['CHANNEL'], chan) and h['CHANNEL'].dtype.kind == 'i'
    assert np.allclose(h['SCALED'], [6.0, 11.0, 16.0])
    assert h['FLAG'].tolist() == [True, False, True] and h['NAME'].tolist() == ['ab', 'c', '']
    assert [v.tolist() for v in h['VAR']] == [[1.5, 2.5], [], [7.0]]
