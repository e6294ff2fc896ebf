#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE for the hot path.

The reference (wcxve/elisa, /root/reference) is pure Python on JAX/numpyro, neither of
which is installable here.  Its arithmetic for this path, however, only touches
`jax.numpy` element-wise functions, `jax.scipy.special.xlogy`, `jax.scipy.stats.norm`
and a few numpyro primitives that merely record values.  This script therefore
  * installs thin stand-in modules (`jax.numpy` -> NumPy float64, `numpyro` -> a recorder),
  * imports `elisa/infer/likelihood.py` UNMODIFIED from /root/reference through them, and
  * pulls the bodies of the component `continuum` / `integral` static methods
    (models/add.py, models/mul.py), `_powerlaw_integral`, `PLFluxNorm.eval`'s inner
    integral and `NumericalIntegral._make_integral`'s trapz / Simpson closures
    (models/model.py:1731-1761) out of the reference files with `ast` and executes
    them as they are written,
so every number in the fixtures is produced by reference code, not by a restatement.
NumPy and XLA:CPU both evaluate these float64 primitives to within a few ulp, which is
far inside the 1e-10 parity tolerance.  Gradients are central differences (Richardson)
of those same reference functions.

Run in the BUILD container only (needs /root/reference); the fixtures are committed and
travel to the GPU box, this script's input does not.

    python tools/make_golden.py [--photonabs | --conv | --ogip]
"""

from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.special
import scipy.stats

REF = '/root/reference/src/elisa'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


# --------------------------------------------------------------------------- jax.numpy stand-in
class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Ref:
            def set(self, v):
                out = arr.copy()
                out[idx] = v
                return out

            def add(self, v):
                out = arr.copy()
                np.add.at(out, idx, v)
                return out

        return _Ref()


class A(np.ndarray):
    """ndarray with the `.at[...]` accessor of jax arrays (functional updates)."""

    @property
    def at(self):
        return _At(self)


def _a(x):
    return np.asarray(x, dtype=np.float64).view(A) if not isinstance(x, (bool, np.bool_)) else x


def _wrapped(f):
    def g(*a, **k):
        r = f(*a, **k)
        return r.view(A) if isinstance(r, np.ndarray) and r.dtype == np.float64 else r

    return g


def make_jnp():
    m = types.ModuleType('jax.numpy')
    for name in ('where', 'power', 'exp', 'log', 'expm1', 'arctan', 'not_equal', 'less', 'less_equal',
                 'greater_equal', 'greater', 'equal', 'bitwise_or', 'abs', 'sqrt', 'hstack', 'shape', 'size',
                 'broadcast_to', 'add', 'multiply', 'interp', 'sum', 'result_type'):
        setattr(m, name, _wrapped(getattr(np, name)))
    m.pi = np.pi
    m.ndarray = np.ndarray
    m.number = np.number
    m.array = lambda x, dtype=None: _a(x)
    m.asarray = lambda x, dtype=None: _a(x)
    m.full = lambda n, v: np.full(n, v)
    m.clip = lambda x, min=None, max=None: np.clip(x, min, max)

    def flatnonzero(mask, size=None, fill_value=0):
        idx = np.flatnonzero(mask)
        if size is not None:
            out = np.full(size, fill_value, dtype=idx.dtype)
            out[: min(size, idx.size)] = idx[:size]
            return out
        return idx

    m.flatnonzero = flatnonzero
    return m


def install_stand_ins():
    jnp = make_jnp()
    jax = types.ModuleType('jax')
    jax.numpy = jnp
    jax.jit = lambda f, **k: f
    jax.Array = np.ndarray
    jax.core = types.SimpleNamespace(Tracer=type('Tracer', (), {}))
    jax.device_get = lambda x: x
    lax = types.ModuleType('jax.lax')
    lax.broadcast_shapes = np.broadcast_shapes
    jax.lax = lax
    jscipy = types.ModuleType('jax.scipy')
    special = types.ModuleType('jax.scipy.special')
    special.xlogy = scipy.special.xlogy
    stats = types.ModuleType('jax.scipy.stats')
    stats.norm = types.SimpleNamespace(pdf=lambda x, loc=0.0, scale=1.0: scipy.stats.norm.pdf(x, loc=loc, scale=scale))
    jscipy.special, jscipy.stats = special, stats
    exp = types.ModuleType('jax.experimental')
    sparse = types.ModuleType('jax.experimental.sparse')
    sparse.BCSR = type('BCSR', (), {})
    exp.sparse = sparse

    # numpyro: record deterministic sites; distributions only need loc/scale/rate
    record = {}
    numpyro = types.ModuleType('numpyro')

    def deterministic(name, value):
        record[name] = np.array(value, dtype=np.float64)
        return value

    class plate:
        def __init__(self, name, size):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    numpyro.deterministic = deterministic
    numpyro.plate = plate
    numpyro.sample = lambda name, fn, obs=None: obs
    numpyro.primitives = types.SimpleNamespace(mutable=lambda name, value: value)
    dist = types.ModuleType('numpyro.distributions')

    class Normal:
        def __init__(self, loc, scale):
            self.loc, self.scale = loc, scale

    class Poisson:
        is_sparse = False
        _validate_args = False

        def __init__(self, rate):
            self.rate = rate

    dist.Normal, dist.Poisson = Normal, Poisson
    util = types.ModuleType('numpyro.distributions.util')
    util.validate_sample = lambda f: f
    dist.util = util
    numpyro.distributions = dist

    sys.modules.update({
        'jax': jax, 'jax.numpy': jnp, 'jax.lax': lax, 'jax.scipy': jscipy, 'jax.scipy.special': special,
        'jax.scipy.stats': stats, 'jax.experimental': exp, 'jax.experimental.sparse': sparse,
        'numpyro': numpyro, 'numpyro.distributions': dist, 'numpyro.distributions.util': util,
    })
    return jnp, stats, record


# --------------------------------------------------------------------------- source extraction
def _src(path):
    with open(os.path.join(REF, path)) as f:
        return f.read()


def _exec_function(node: ast.FunctionDef, glb: dict):
    node = ast.FunctionDef(name=node.name, args=node.args, body=node.body, decorator_list=[], returns=None,
                           type_comment=None, type_params=[])
    # annotations are strings under `from __future__ import annotations`; drop them anyway
    for a in node.args.args + node.args.kwonlyargs:
        a.annotation = None
    mod = ast.Module(body=[node], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, '<reference>', 'exec'), glb)
    return glb[node.name]


def load_components(jnp, stats):
    """{class name: ('continuum' | 'integral', function)} executed from the reference files."""
    out = {}
    glb = {'jnp': jnp, 'stats': stats, 'jax': sys.modules['jax']}
    for path in ('models/add.py', 'models/mul.py'):
        tree = ast.parse(_src(path))
        for node in tree.body:
            if isinstance(node, ast.FunctionDef) and node.name == '_powerlaw_integral':
                _exec_function(node, glb)
        for node in tree.body:
            if not isinstance(node, ast.ClassDef):
                continue
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in ('continuum', 'integral'):
                    fn = _exec_function(item, dict(glb))
                    out[node.name] = (item.name, fn)
            if node.name == 'PLFluxNorm':  # inner `integral` of the eval property (add.py:682-708)
                for item in node.body:
                    if isinstance(item, ast.FunctionDef) and item.name == 'eval':
                        inner = [n for n in ast.walk(item) if isinstance(n, ast.FunctionDef) and n.name == 'integral'][0]
                        out['PLFluxNorm.eval.integral'] = inner
    out['_glb'] = glb
    return out


def load_integrators(jnp):
    """The trapz / simpson closures of NumericalIntegral._make_integral (model.py:1731-1761)."""
    tree = ast.parse(_src('models/model.py'))
    fns = []
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == '_make_integral':
            fns = [n for n in ast.walk(node) if isinstance(n, ast.FunctionDef) and n.name == 'fn']
    assert len(fns) == 2, 'expected the trapz and simpson closures'

    def make(method, mtype, continuum):
        glb = {'jnp': jnp, 'mtype': mtype, 'continuum': continuum}
        return _exec_function(fns[0 if method == 'trapz' else 1], glb)

    return make


def load_likelihood():
    spec = importlib.util.spec_from_file_location('ref_likelihood', os.path.join(REF, 'infer/likelihood.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# --------------------------------------------------------------------------- reference evaluation
# (class, parameter names in _config order, defaults) -- read from the reference's _config tuples
def read_configs():
    cfgs = {}
    for path in ('models/add.py', 'models/mul.py'):
        tree = ast.parse(_src(path))
        for node in tree.body:
            if not isinstance(node, ast.ClassDef):
                continue
            for item in node.body:
                if isinstance(item, ast.Assign) and item.targets[0].id == '_config':
                    names, defaults = [], []
                    for call in item.value.elts:
                        args = [ast.literal_eval(a) for a in call.args]
                        names.append(args[0])
                        defaults.append(float(args[3]))
                    cfgs[node.name] = (names, defaults)
    return cfgs


class RefModel:
    """Evaluate reference components on a grid: bins = integral(egrid, params) or the
    reference's own trapz / simpson closure around continuum."""

    def __init__(self):
        self.jnp, self.stats, self.record = install_stand_ins()
        self.comps = load_components(self.jnp, self.stats)
        self.make_integral = load_integrators(self.jnp)
        self.cfgs = read_configs()
        self.add_names = ast.literal_eval(_all_list('models/add.py'))
        self.mul_names = ast.literal_eval(_all_list('models/mul.py'))

    def eval(self, name, egrid, values, method='trapz', emin=None, emax=None):
        egrid = _a(egrid)
        names = self.cfgs[name][0]
        params = {k: np.float64(v) for k, v in zip(names, values)}
        if name in ('PLPhFlux', 'PLEnFlux'):
            fn = self.comps['PLFluxNorm'][1]
            glb = dict(self.comps['_glb'], fn=fn, emin=emin, emax=emax, energy=(name == 'PLEnFlux'))
            integral = _exec_function(self.comps['PLFluxNorm.eval.integral'], glb)
            return np.asarray(integral(egrid, params))
        kind, fn = self.comps[name]
        if kind == 'integral':
            return np.asarray(fn(egrid, params)) + np.zeros(egrid.size - 1)
        mtype = 'add' if name in self.add_names else 'mul'
        return np.asarray(self.make_integral(method, mtype, fn)(egrid, params))


def _all_list(path):
    tree = ast.parse(_src(path))
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], 'id', None) == '__all__':
            return ast.unparse(node.value)
    raise RuntimeError('__all__ not found')


# --------------------------------------------------------------------------- fixtures
def component_fixture(ref: RefModel):
    """Bin values of every built-in add / mul component at its defaults and at two
    perturbed parameter vectors, trapz and simpson, on the reference test-suite grid
    geomspace(1e-4, 1e9, 1301) (tests/test_model.py:178-195) and on a keV-band grid."""
    rng = np.random.default_rng(2024)
    grids = {'wide': np.geomspace(1e-4, 1e9, 1301), 'kev': np.geomspace(0.3, 80.0, 257)}
    out = {f'grid_{k}': v for k, v in grids.items()}
    names = [n for n in ref.add_names + ref.mul_names if n in ref.cfgs and n not in ('PhAbs', 'TBAbs', 'WAbs')]
    out['names'] = np.array(names)
    for name in names:
        pnames, defaults = ref.cfgs[name]
        d = np.array(defaults)
        thetas = np.stack([d, d * (1.0 + 0.2 * rng.random(d.size)), d * (1.0 - 0.2 * rng.random(d.size))])
        if name in ('PowerLaw', 'PLPhFlux', 'PLEnFlux', 'BrokenPL'):
            thetas[1, 0] = 1.0  # the alpha == 1 branch of _powerlaw_integral
        out[f'{name}/theta'] = thetas
        numerical = ref.comps.get(name, ('integral',))[0] == 'continuum'
        for gname, grid in grids.items():
            for method in (('trapz', 'simpson') if numerical else ('trapz',)):
                kw = dict(emin=1.0, emax=10.0) if name in ('PLPhFlux', 'PLEnFlux') else {}
                with np.errstate(all='ignore'):
                    vals = np.stack([ref.eval(name, grid, t, method, **kw) for t in thetas])
                out[f'{name}/{gname}/{method}'] = vals
    return out


def likelihood_fixture(ref: RefModel):
    """The five statistic closures of infer/likelihood.py run unmodified (through the
    recorder stand-in of numpyro) on a small dense-response dataset with a
    CutoffPL + Blackbody model whose bins come from the reference component code."""
    lk = load_likelihood()
    rng = np.random.default_rng(7)
    E, C, N = 96, 40, 6
    egrid = np.geomspace(0.3, 60.0, E + 1)
    R = np.abs(rng.standard_normal((E, C))) * 20.0 + 1.0
    counts = rng.poisson(40.0, C).astype(float)
    back = rng.poisson(15.0, C).astype(float)
    counts[::6] = 0.0
    back[::5] = 0.0
    counts[10] = back[10] = 0.0
    ratio = np.full(C, 0.4)
    data = types.SimpleNamespace(
        name='g', spec_counts=counts, net_counts=counts - ratio * back, net_errors=np.sqrt(counts + 1.0),
        back_counts=back, back_errors=np.sqrt(back + 1.0), back_ratio=ratio, has_back=True, photon_egrid=egrid,
        channel_width=np.linspace(0.5, 2.0, C), response_matrix=R, response_sparse=False,
        area_scale=0.5 + rng.random(C), spec_exposure=3.0)
    truth = np.array([1.2, 20.0, 5.0, 1.5, 2.0])
    theta = truth * (1.0 + 0.1 * rng.standard_normal((N, 5)))
    theta[0, 2] = theta[0, 4] = 1e-12  # drives source counts towards the lower clip

    def model(photon_egrid, params):
        return _a(ref.eval('CutoffPL', photon_egrid, params[:3]) + ref.eval('Blackbody', photon_egrid, params[3:]))

    out = dict(photon_egrid=egrid, response_matrix=R, spec_counts=counts, net_counts=data.net_counts,
               net_errors=data.net_errors, back_counts=back, back_errors=data.back_errors, back_ratio=ratio,
               channel_width=data.channel_width, area_scale=data.area_scale, spec_exposure=np.float64(3.0), theta=theta)

    def total(stat, th):
        ref.record.clear()
        getattr(lk, stat)(data, model)(th, False)
        return float(np.sum(ref.record['g_loglike']))

    for stat in ('chi2', 'cstat', 'pstat', 'pgstat', 'wstat'):
        closure = getattr(lk, stat)(data, model)
        sites = {}
        for n in range(N):
            ref.record.clear()
            closure(theta[n], False)
            for k, v in ref.record.items():
                sites.setdefault(k, []).append(v.copy())
        for k, v in sites.items():
            out[f'{stat}/{k}'] = np.stack(v)
        # d(total loglike)/d theta by Richardson-extrapolated central differences
        g = np.zeros((N, 5))
        for n in range(1, N):
            for j in range(5):
                h = 1e-4 * abs(theta[n, j])
                d = []
                for hh in (h, h / 2):
                    tp, tm = theta[n].copy(), theta[n].copy()
                    tp[j] += hh
                    tm[j] -= hh
                    d.append((total(stat, tp) - total(stat, tm)) / (2 * hh))
                g[n, j] = (4 * d[1] - d[0]) / 3
        out[f'{stat}/fd_grad'] = g
    # the profile backgrounds on their own, over a grid that hits every branch
    s = np.array([1e-30, 0.5, 3.0, 8.0, 40.0, 1e3])[:, None]
    n_on = np.array([0.0, 0.0, 5.0, 12.0, 30.0])[None, :]
    n_off = np.array([0.0, 7.0, 0.0, 9.0, 25.0])[None, :]
    out['bg/s'], out['bg/n_on'], out['bg/n_off'] = s, n_on, n_off
    out['bg/wstat'] = np.asarray(lk.wstat_background(_a(s), _a(n_on), _a(n_off), 0.4))
    b_est = np.array([-2.0, 0.5, 3.0, 9.0, 25.0])[None, :]
    out['bg/b_est'] = b_est
    out['bg/pgstat'] = np.asarray(lk.pgstat_background(_a(s), _a(n_on), _a(b_est), 2.0, 0.4))
    return out


def synthetic_xsect_table():
    """A photoelectric-like cross-section table (decreasing power law with K edges), in the
    units of the reference's tables (1e-22 cm^2): plausible input, not the reference's data."""
    energy = np.geomspace(0.03, 40.0, 300)
    xsect = 2.4e-2 * energy**-2.6 * (1.0 + 1.5 * (energy > 0.532) + 0.8 * (energy > 1.84) + 2.5 * (energy > 7.11))
    return energy, xsect


def photonabs_fixture(ref: RefModel):
    """PhotonAbsorption (PhAbs / TBAbs / WAbs): the reference's `_make_interp`
    (models/mul.py:274-283) and `PhotonAbsorption.continuum` (:360-388) executed as written,
    through its own trapz / Simpson closures.  `jnp.interp(..., left='extrapolate',
    right='extrapolate')` is third-party (jax): the stand-in below follows jax's
    documented implementation (jax/_src/numpy/lax_numpy.py `_interp`)."""
    jnp = ref.jnp

    def interp(x, xp, fp, left=None, right=None, period=None):
        assert left == 'extrapolate' and right == 'extrapolate' and period is None
        x, xp, fp = (np.asarray(v, dtype=np.float64) for v in (x, xp, fp))
        i = np.clip(np.searchsorted(xp, x, side='right'), 1, len(xp) - 1)
        df = fp[i] - fp[i - 1]
        dx = xp[i] - xp[i - 1]
        delta = x - xp[i - 1]
        dx0 = np.abs(dx) <= np.spacing(np.finfo(np.float64).eps)
        return np.where(dx0, fp[i - 1], fp[i - 1] + (delta / np.where(dx0, 1, dx)) * df)

    jnp.interp = interp
    tree = ast.parse(_src('models/mul.py'))
    glb = {'jnp': jnp, 'np': np, 'jax': sys.modules['jax']}
    make_interp = None
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == '_make_interp':
            make_interp = _exec_function(node, glb)
    assert make_interp is not None
    energy, xsect = synthetic_xsect_table()
    kind, cont = ref.comps['PhotonAbsorption']
    assert kind == 'continuum'
    cont.__globals__['_XSECT_INTERP'] = {'phabs': {'vern': {'angr': make_interp(energy, xsect)}}}
    continuum = lambda egrid, params: cont(egrid, params, 'phabs', 'angr', 'vern')  # noqa: E731
    out = {'table_energy': energy, 'table_xsect': xsect}
    grids = {'inside': np.geomspace(0.1, 12.0, 151), 'beyond': np.geomspace(0.01, 300.0, 97)}
    nH = np.array([0.05, 1.0, 7.5])
    out['nH'] = nH
    for gname, g in grids.items():
        out[f'grid_{gname}'] = g
        for method in ('trapz', 'simpson'):
            fn = ref.make_integral(method, 'mul', continuum)
            out[f'{gname}_{method}'] = np.stack([np.asarray(fn(_a(g), {'nH': np.float64(v)})) for v in nH])
    # as the reference's own table file: FLOAT32 datasets (np.log of them is a float32 log,
    # mul.py:275-276) with repeated energies (xsect.hdf5 has 87), here also at the upper end
    idx = np.sort(np.concatenate([np.arange(energy.size), [17, 18, 140, energy.size - 1]]))
    e32, x32 = energy.astype(np.float32)[idx], xsect.astype(np.float32)[idx]
    assert e32.dtype == np.float32 and np.log(e32).dtype == np.float32 and np.sum(np.diff(e32) == 0) == 4
    cont.__globals__['_XSECT_INTERP'] = {'phabs': {'vern': {'angr': make_interp(e32, x32)}}}
    out['table32_energy'], out['table32_xsect'] = e32, x32
    for gname, g in grids.items():
        for method in ('trapz', 'simpson'):
            fn = ref.make_integral(method, 'mul', continuum)
            out[f'f32_{gname}_{method}'] = np.stack([np.asarray(fn(_a(g), {'nH': np.float64(v)})) for v in nH])
    return out


def load_convolutions(jnp):
    """{class name: convolve} executed from models/conv.py as written."""
    out = {}
    glb = {'jnp': jnp, 'jax': sys.modules['jax']}
    tree = ast.parse(_src('models/conv.py'))
    for node in tree.body:
        if not isinstance(node, ast.ClassDef):
            continue
        for item in node.body:
            if isinstance(item, ast.FunctionDef) and item.name == 'convolve' and node.name != 'NormConvolution':
                out[node.name] = _exec_function(item, dict(glb))
    return out


def conv_fixture(ref: RefModel):
    """tests/conv_cases.py evaluated by the reference: component bins from its continuum /
    integral code (RefModel.eval), composition on bin values (models/model.py:1285-1295),
    and every convolution through the `convolve` static methods of models/conv.py executed
    as written (ConvolvedModel.eval, models/model.py:1945-1955); the flux grid as
    NormConvolution.eval builds it (conv.py:81-84)."""
    sys.path.insert(0, os.path.join(os.path.dirname(OUT), ''))
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    import conv_cases

    jnp = ref.jnp
    jnp.geomspace = _wrapped(np.geomspace)
    jnp.linspace = _wrapped(np.linspace)
    conv = load_convolutions(jnp)
    assert set(conv) == {'PhFlux', 'EnFlux', 'ZAShift', 'ZMShift', 'VAShift', 'VMShift'}, sorted(conv)

    def val(ref_, row):
        return np.float64(row[ref_[1]]) if ref_[0] == 'theta' else np.float64(ref_[1])

    def ev(spec, egrid, row):
        kind = spec[0]
        if kind == 'comp':
            _, name, refs, extra = spec
            names = ref.cfgs[name][0]
            kw = {k: extra[k] for k in ('emin', 'emax') if k in extra}
            return ref.eval(name, egrid, [val(refs[n], row) for n in names], extra.get('method', 'trapz'), **kw)
        if kind == 'conv':
            _, name, refs, extra, inner = spec
            params = {k: val(v, row) for k, v in refs.items()}
            model_fn = lambda e: _a(ev(inner, e, row))  # noqa: E731
            if name in ('PhFlux', 'EnFlux'):
                make = jnp.geomspace if extra['elog'] else jnp.linspace
                return np.asarray(conv[name](_a(egrid), params, model_fn, make(extra['emin'], extra['emax'], extra['ngrid'])))
            return np.asarray(conv[name](_a(egrid), params, model_fn))
        lhs, rhs = ev(spec[1], egrid, row), ev(spec[2], egrid, row)
        return lhs + rhs if kind == '+' else lhs * rhs

    g = conv_cases.grid()
    out = {'grid': g}
    for name, (model, theta) in conv_cases.cases().items():
        spec = model.compile().to_spec()
        out[f'{name}/theta'] = theta
        out[f'{name}/bins'] = np.stack([ev(spec, g, row) for row in theta])
    return out


def ogip_fixture():
    """tests/ogip_cases.py read by the reference's own `elisa.data.ogip.Data` (data/ogip.py,
    data/base.py, data/grouping.py imported UNMODIFIED from /root/reference) and frozen with
    `get_fixed_data()`.  astropy is not installable here: `astropy.io.fits` is replaced by a
    facade over elisa_b200.fitsio (header dict + column access -- the decode of the columns into
    a matrix, the grouping, masks, scales and ratios are all the reference's code), and
    `astropy.units.Unit(u).to('keV')` by a table of the energy units."""
    import tempfile

    sys.path.insert(0, os.path.dirname(OUT))
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    import ogip_cases
    from elisa_b200.fitsio import read_fits

    class Rec:  # the slice of the FITS_rec interface data/ogip.py uses
        def __init__(self, hdu, rows=None):
            self.hdu, self.rows = hdu, rows
            self.names = list(hdu.names)

        def __len__(self):
            return len(self.hdu) if self.rows is None else len(range(len(self.hdu))[self.rows])

        def __getitem__(self, key):
            if isinstance(key, slice):
                return Rec(self.hdu, key)
            col = self.hdu[key]
            if isinstance(col, list):
                arr = np.empty(len(col), dtype=object)
                arr[:] = col
                col = arr
            return col if self.rows is None else col[self.rows]

    class FakeHDU:
        def __init__(self, h):
            self.name, self.ver, self.header = h.name, h.ver, h.header
            self.data = Rec(h) if h.names else None

    class HDUList(list):
        def __contains__(self, name):
            return any(h.name == name for h in self)

        def __getitem__(self, key):
            if isinstance(key, tuple):
                return next(h for h in self if h.name == key[0] and h.ver == key[1])
            if isinstance(key, str):
                return next(h for h in self if h.name == key)
            return list.__getitem__(self, key)

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    fits = types.ModuleType('astropy.io.fits')
    fits.open = lambda f: HDUList(FakeHDU(h) for h in read_fits(f))
    fits.Header = dict
    units = types.ModuleType('astropy.units')
    to_kev = {'ev': 1e-3, 'kev': 1.0, 'mev': 1e3}
    units.Unit = lambda u: types.SimpleNamespace(to=lambda target: to_kev[u.lower()] / to_kev[target.lower()])
    astropy, aio = types.ModuleType('astropy'), types.ModuleType('astropy.io')
    astropy.io, astropy.units, aio.fits = aio, units, fits
    mpl, plt = types.ModuleType('matplotlib'), types.ModuleType('matplotlib.pyplot')
    mpl.pyplot = plt
    plt.Figure = object
    stubs = {'astropy': astropy, 'astropy.io': aio, 'astropy.io.fits': fits, 'astropy.units': units,
             'matplotlib': mpl, 'matplotlib.pyplot': plt}
    # the package skeleton: only data/{grouping,base,ogip}.py run, from the reference tree
    for name in ('elisa', 'elisa.data', 'elisa.plot', 'elisa.plot.misc', 'elisa.util', 'elisa.util.misc'):
        stubs[name] = types.ModuleType(name)
    stubs['elisa.plot.misc'].get_colors = lambda *a, **k: []
    tree = ast.parse(_src('util/misc.py'))
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'to_native_byteorder'][0]
    stubs['elisa.util.misc'].to_native_byteorder = _exec_function(fn, {'np': np})
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        mods = {}
        for name in ('grouping', 'base', 'ogip'):
            spec = importlib.util.spec_from_file_location(f'elisa.data.{name}', os.path.join(REF, 'data', f'{name}.py'))
            mods[name] = importlib.util.module_from_spec(spec)
            sys.modules[f'elisa.data.{name}'] = mods[name]
            spec.loader.exec_module(mods[name])
        Data = mods['ogip'].Data
        out = {}
        import warnings

        with tempfile.TemporaryDirectory() as tmp:
            cwd = os.getcwd()
            os.chdir(tmp)  # BACKFILE / RESPFILE keywords are resolved relative to the working directory
            try:
                for case, kw in ogip_cases.build(tmp).items():
                    kw = {k: v for k, v in kw.items() if not k.startswith('_')}
                    with warnings.catch_warnings():
                        warnings.simplefilter('ignore')
                        fd = Data(**kw).get_fixed_data()
                    for f in ogip_cases.FIELDS:
                        v = getattr(fd, f)
                        if v is not None:
                            out[f'{case}/{f}'] = np.asarray(v, dtype=np.float64)
                    coo = fd.sparse_matrix
                    out[f'{case}/sparse_dense'] = np.asarray(coo.todense(), dtype=np.float64)
                    out[f'{case}/channel'] = np.asarray(fd.channel, dtype=str)
            finally:
                os.chdir(cwd)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for name in ('grouping', 'base', 'ogip'):
            sys.modules.pop(f'elisa.data.{name}', None)
    return out


def main():
    if '--ogip' in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        og = ogip_fixture()
        np.savez_compressed(os.path.join(OUT, 'ogip.npz'), **og)
        print('ogip.npz:', len(og), 'arrays,', os.path.getsize(os.path.join(OUT, 'ogip.npz')), 'bytes')
        return
    ref = RefModel()
    os.makedirs(OUT, exist_ok=True)
    if '--conv' in sys.argv or '--all' in sys.argv or len(sys.argv) == 1:
        cv = conv_fixture(ref)
        np.savez_compressed(os.path.join(OUT, 'conv.npz'), **cv)
        print('conv.npz:', len(cv), 'arrays,', os.path.getsize(os.path.join(OUT, 'conv.npz')), 'bytes')
        if '--conv' in sys.argv:
            return
        ref = RefModel()
    if '--photonabs' in sys.argv or '--all' in sys.argv or len(sys.argv) == 1:
        pa = photonabs_fixture(ref)
        np.savez_compressed(os.path.join(OUT, 'photonabs.npz'), **pa)
        print('photonabs.npz:', len(pa), 'arrays,', os.path.getsize(os.path.join(OUT, 'photonabs.npz')), 'bytes')
        if '--photonabs' in sys.argv:
            return
        ref = RefModel()  # fresh stand-ins (photonabs_fixture replaces jnp.interp)
    comp = component_fixture(ref)
    np.savez_compressed(os.path.join(OUT, 'components.npz'), **comp)
    like = likelihood_fixture(ref)
    np.savez_compressed(os.path.join(OUT, 'likelihood.npz'), **like)
    print('components.npz:', len(comp), 'arrays;', 'likelihood.npz:', len(like), 'arrays')
    for f in ('components.npz', 'likelihood.npz'):
        print(f, os.path.getsize(os.path.join(OUT, f)), 'bytes')


if __name__ == '__main__':
    main()
