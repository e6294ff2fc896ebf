"""Host-side mirror of ``elisa.models`` for the hot path: built-in components, ``+`` / ``*``
composition and ``Model.compile()`` -> ``CompiledModel`` (reference: models/model.py:91,
293-303, 415; models/add.py; models/mul.py).

Only what the forward-folding likelihood path needs is mirrored: component names,
parameter names/order/defaults (the reference's ``_config`` tuples), fixed / free /
linked parameters, the integration ``method`` of numerical-integral components and the
composition type rules (models/model.py:1228-1258).  A compiled model is a small postfix
program handed to the CUDA library; ``CompiledModel.eval`` runs it on the GPU.
"""

from __future__ import annotations

from typing import NamedTuple, Sequence

import numpy as np

from . import _lib

__all__ = [
    'Band', 'BandEp', 'BentPL', 'Blackbody', 'BlackbodyRad', 'BrokenPL', 'Compt', 'CutoffPL', 'Gauss',
    'Lorentz', 'OTTB', 'OTTS', 'PLEnFlux', 'PLPhFlux', 'PowerLaw', 'SmoothlyBrokenPL',
    'Constant', 'Edge', 'ExpAbs', 'ExpFac', 'GAbs', 'HighECut', 'PLAbs',
    'PhAbs', 'TBAbs', 'WAbs', 'PhotonAbsorption',
    'EnFlux', 'PhFlux', 'ZAShift', 'ZMShift', 'VAShift', 'VMShift',
    'UniformParameter', 'Model', 'CompiledModel',
]


class ParamConfig(NamedTuple):
    """models/model.py:1542-1567 (latex / unit dropped: not used on this path)."""

    name: str
    default: float
    min: float
    max: float
    log: bool = False
    fixed: bool = False


class UniformParameter:
    """Minimal stand-in for elisa.models.parameter.UniformParameter (parameter.py:453):
    a named scalar with default / bounds, free or fixed.  Using ONE instance in two
    components links them (one column of theta)."""

    def __init__(self, name, default, min, max, log=False, fixed=False):
        self.name = str(name)
        self.default = float(default)
        self.min = float(min)
        self.max = float(max)
        self.log = bool(log)
        self.fixed = bool(fixed)
        if not (self.min <= self.default <= self.max):
            raise ValueError(f'default {default} of {name} outside [{min}, {max}]')

    def __repr__(self):
        return f'UniformParameter({self.name!r}, {self.default}, fixed={self.fixed})'


class Model:
    """Base of the model algebra (models/model.py:1204-1295)."""

    type: str  # 'add' | 'mul'

    def __add__(self, other):
        return CompositeModel(self, other, '+')

    def __mul__(self, other):
        return CompositeModel(self, other, '*')

    def _leaves(self):
        raise NotImplementedError

    def _postfix(self, index):
        raise NotImplementedError

    def compile(self) -> 'CompiledModel':
        """Model.compile, models/model.py:91."""
        return CompiledModel(self)


class Component(Model):
    """One built-in component (models/model.py:1570-1874 + the metaclass ctor :1334-1416):
    each parameter argument may be None (free, reference defaults), a number (fixed to
    it) or a UniformParameter instance."""

    _kind: int
    _config: tuple[ParamConfig, ...]
    _numerical: bool = False
    type = 'add'

    def __init__(self, *args, method: str | None = None, **params):
        names = [c.name for c in self._config]
        if len(args) > len(names):
            raise TypeError(f'{type(self).__name__} takes at most {len(names)} positional parameters')
        for n, a in zip(names, args):
            if n in params:
                raise TypeError(f'{n} given twice')
            params[n] = a
        unknown = set(params) - set(names)
        if unknown:
            raise TypeError(f'{type(self).__name__}: unknown parameter(s) {sorted(unknown)}')
        if method is None:
            method = 'trapz'
        if method not in _lib.METHOD:
            raise NotImplementedError(f"integration method '{method}'")  # models/model.py:1759
        self.method = method
        self.name = type(self).__name__
        self.params: dict[str, UniformParameter] = {}
        for cfg in self._config:
            p = params.get(cfg.name)
            if p is None:
                p = UniformParameter(cfg.name, cfg.default, cfg.min, cfg.max, cfg.log, cfg.fixed)
            elif isinstance(p, (int, float, np.number)) or (isinstance(p, np.ndarray) and p.shape == ()):
                p = UniformParameter(cfg.name, float(p), min(cfg.min, float(p)), max(cfg.max, float(p)), cfg.log, True)
            elif isinstance(p, np.ndarray):
                raise ValueError('scalar value is expected')
            elif not isinstance(p, UniformParameter):
                raise TypeError(f'{cfg.name}: expected None, a number or a UniformParameter')
            self.params[cfg.name] = p
        self.aux = (0.0, 0.0)

    @property
    def param_names(self):
        return tuple(c.name for c in self._config)

    def __getitem__(self, key):
        return self.params[key]

    def _leaves(self):
        return [self]

    def _postfix(self, index):
        return [index[id(self)]]

    def __repr__(self):
        return self.name


class CompositeModel(Model):
    """models/model.py:1204-1295: '+' needs two additive operands; '*' forbids
    additive * additive; the result of '*' is additive if either operand is."""

    def __init__(self, lhs, rhs, op):
        if not (isinstance(lhs, Model) and isinstance(rhs, Model)):
            raise TypeError(
                f"unsupported operand types for {op}: '{type(lhs).__name__}' and '{type(rhs).__name__}'"
            )
        if 'conv' in (lhs.type, rhs.type):
            raise TypeError(f'unsupported model type for {op}: a convolution model must be applied to a model first')
        if op == '+':
            if lhs.type != 'add':
                raise TypeError(f'{lhs} is not an additive model')
            if rhs.type != 'add':
                raise TypeError(f'{rhs} is not an additive model')
            self.type = 'add'
        elif op == '*':
            if lhs.type == 'add' and rhs.type == 'add':
                raise TypeError(f'unsupported model type for *: {lhs} (additive) and {rhs} (additive)')
            self.type = 'add' if 'add' in (lhs.type, rhs.type) else 'mul'
        else:
            raise NotImplementedError(f'op {op}')
        self.lhs, self.rhs, self.op = lhs, rhs, op

    def _leaves(self):
        out = list(self.lhs._leaves())
        for c in self.rhs._leaves():
            if all(c is not o for o in out):
                out.append(c)
        return out

    def _postfix(self, index):
        return self.lhs._postfix(index) + self.rhs._postfix(index) + [_lib.OP_ADD if self.op == '+' else _lib.OP_MUL]

    def __repr__(self):
        def wrap(m):
            s = repr(m)
            return f'({s})' if self.op == '*' and isinstance(m, CompositeModel) and m.op == '+' else s

        return f'{wrap(self.lhs)} {self.op} {wrap(self.rhs)}'


# --------------------------------------------------------------------------- built-ins
def _component(name, kind, mtype, numerical, config, extra_args=()):
    ns = {'_kind': kind, '_config': tuple(ParamConfig(*c) for c in config), '_numerical': numerical, 'type': mtype}
    if extra_args:  # PLPhFlux / PLEnFlux: (emin, emax) come first (models/add.py:660-680)

        def __init__(self, emin, emax, *args, **kw):
            emin, emax = float(emin), float(emax)
            if emin >= emax:
                raise ValueError('emin must be less than emax')
            Component.__init__(self, *args, **kw)
            self.aux = (emin, emax)

        ns['__init__'] = __init__
    return type(name, (Component,), ns)


# (name, default, min, max[, log[, fixed]]) in the reference's _config order
Band = _component('Band', 0, 'add', True, [('alpha', -1.0, -10.0, 5.0), ('beta', -2.0, -10.0, 10.0), ('Ec', 300.0, 10.0, 1e4), ('K', 1.0, 1e-10, 1e10)])
BandEp = _component('BandEp', 1, 'add', True, [('alpha', -1.0, -1.99, 10.0), ('beta', -2.0, -10.0, 10.0), ('Ep', 300.0, 10.0, 1e4), ('K', 1.0, 1e-10, 1e10)])
BentPL = _component('BentPL', 2, 'add', True, [('alpha', 1.5, -3.0, 10.0), ('Eb', 5.0, 1e-2, 1e6), ('K', 1.0, 1e-10, 1e10)])
Blackbody = _component('Blackbody', 3, 'add', True, [('kT', 3.0, 1e-4, 200.0), ('K', 1.0, 1e-10, 1e10)])
BlackbodyRad = _component('BlackbodyRad', 4, 'add', True, [('kT', 3.0, 1e-4, 200.0), ('K', 1.0, 1e-10, 1e10)])
BrokenPL = _component('BrokenPL', 5, 'add', False, [('alpha1', 1.0, -3.0, 10.0), ('Eb', 5.0, 1e-2, 1e6), ('alpha2', 2.0, -3.0, 10.0), ('K', 1.0, 1e-10, 1e10)])
Compt = _component('Compt', 6, 'add', True, [('alpha', 1.0, -3.0, 10.0), ('Ep', 15.0, 0.01, 1e4), ('K', 1.0, 1e-10, 1e10)])
CutoffPL = _component('CutoffPL', 7, 'add', True, [('alpha', 1.0, -3.0, 10.0), ('Ec', 15.0, 0.01, 1e4), ('K', 1.0, 1e-10, 1e10)])
Gauss = _component('Gauss', 8, 'add', True, [('El', 6.5, 0.0, 1e6), ('sigma', 0.1, 0.0, 20.0), ('K', 1.0, 1e-10, 1e10)])
Lorentz = _component('Lorentz', 9, 'add', False, [('El', 6.5, 1e-6, 1e6), ('sigma', 0.1, 1e-3, 20.0), ('K', 1.0, 1e-10, 1e10)])
OTTB = _component('OTTB', 10, 'add', True, [('kT', 30.0, 0.1, 1e3), ('K', 1.0, 1e-10, 1e10)])
OTTS = _component('OTTS', 11, 'add', True, [('Ec', 100.0, 1e-3, 1e3), ('K', 1.0, 1e-10, 1e10)])
PLEnFlux = _component('PLEnFlux', 12, 'add', False, [('alpha', 1.01, -3.0, 10.0), ('F', 1e-12, 1e-30, 1e30, True)], ('emin', 'emax'))
PLPhFlux = _component('PLPhFlux', 13, 'add', False, [('alpha', 1.01, -3.0, 10.0), ('F', 1.0, 0.01, 1e10)], ('emin', 'emax'))
PowerLaw = _component('PowerLaw', 14, 'add', False, [('alpha', 1.01, -3.0, 10.0), ('K', 1.0, 1e-10, 1e10)])
SmoothlyBrokenPL = _component('SmoothlyBrokenPL', 15, 'add', True, [('alpha1', 1.0, -3.0, 10.0), ('Eb', 5.0, 1e-2, 1e6), ('alpha2', 2.0, -3.0, 10.0), ('rho', 1.0, 1e-2, 1e2, False, True), ('K', 1.0, 1e-10, 1e10)])
Constant = _component('Constant', 16, 'mul', False, [('f', 1.0, 1e-5, 1e5)])
Edge = _component('Edge', 17, 'mul', True, [('Ec', 7.0, 0.0, 100.0), ('D', 1.0, 0.0, 10.0)])
ExpAbs = _component('ExpAbs', 18, 'mul', True, [('Ec', 2.0, 0.0, 200.0)])
ExpFac = _component('ExpFac', 19, 'mul', True, [('A', 1.0, 0.0, 1e6), ('f', 1.0, 0.0, 1e6), ('Ec', 0.5, 0.0, 1e6)])
GAbs = _component('GAbs', 20, 'mul', True, [('El', 1.0, 0.0, 1e6), ('sigma', 0.01, 0.0, 20.0), ('tau', 1.0, 0.0, 1e6)])
HighECut = _component('HighECut', 21, 'mul', True, [('Ec', 10.0, 1e-4, 1e6), ('Ef', 15.0, 1e-4, 1e6)])
PLAbs = _component('PLAbs', 22, 'mul', False, [('alpha', 2.0, 0.0, 5.0), ('K', 1.0, 0.0, 100.0)])



class PhotonAbsorption(Component):
    """PhotonAbsorption (models/mul.py:302-388): ``exp(-nH * sigma(E))`` with sigma(E)
    log-log interpolated (and extrapolated) from a cross-section table.  The reference reads
    its tables from ``models/tables/xsect.hdf5`` (dataset ``energy`` and
    ``{phabs|tbabs|wabs}/{xsect}/{abund}``) with h5py; here the caller hands the two arrays
    in -- ``PhAbs(f['energy'][:], f['phabs/vern/angr'][:])`` on the reference side."""

    _kind = 23
    _config = (ParamConfig('nH', 1.0, 0.0, 1e6),)
    _numerical = True
    type = 'mul'

    def __init__(self, table_energy, table_xsect, nH=None, method: str | None = None):
        super().__init__(nH=nH, method=method)
        e, x = np.asarray(table_energy), np.asarray(table_xsect)
        if e.dtype.kind != 'f':
            e = e.astype(np.float64)
        if x.dtype.kind != 'f':
            x = x.astype(np.float64)
        if e.ndim != 1 or e.shape != x.shape or e.size < 2:
            raise ValueError('cross-section table: two 1-d arrays of equal length >= 2')
        if not (np.all(e > 0) and np.all(np.diff(e) >= 0) and e[-1] > e[0] and np.all(x > 0)):
            raise ValueError('cross-section table: positive values, non-decreasing energies')
        self.table_energy, self.table_xsect = e, x
        # `_make_interp` (mul.py:275-276) takes np.log of the arrays AS STORED: the float32
        # datasets of xsect.hdf5 give float32 logarithms, promoted to float64 by jnp.interp
        self.table_log_energy = np.ascontiguousarray(np.log(e), dtype=np.float64)
        self.table_log_xsect = np.ascontiguousarray(np.log(x), dtype=np.float64)

    @classmethod
    def from_tables(cls, path, abund=None, xsect=None, nH=None, method: str | None = None):
        """The component as the reference builds it, ``PhAbs(nH, abund=..., xsect=...)``
        (models/mul.py:302-339, 429-640), from the reference's ``models/tables/xsect.hdf5``
        read by ``elisa_b200.tables`` (no h5py)."""
        from .tables import load_xsect

        which = cls.__name__.lower()
        if which not in ('phabs', 'tbabs', 'wabs'):
            raise TypeError('from_tables: use PhAbs, TBAbs or WAbs')
        t = load_xsect(path, dtype=None)
        abund = cls._default_abund if abund is None else abund
        xsect = cls._default_xsect if xsect is None else xsect
        try:
            sigma = t[which][xsect][abund]
        except KeyError:
            have = {xs: sorted(t[which][xs]) for xs in t.get(which, {})}
            raise ValueError(f'{which}: no table for xsect={xsect!r}, abund={abund!r}; available: {have}') from None
        return cls(t['energy'], sigma, nH=nH, method=method)


class PhAbs(PhotonAbsorption):  # models/mul.py:429-511
    _default_abund, _default_xsect = 'angr', 'vern'


class TBAbs(PhotonAbsorption):  # models/mul.py:514-585
    _default_abund, _default_xsect = 'wilm', 'vern'


class WAbs(PhotonAbsorption):  # models/mul.py:588-640
    _default_abund, _default_xsect = 'aneb', 'wabs'


# --------------------------------------------------------------------------- convolution components
class ConvolutionComponent(Component):
    """models/model.py:1877-2000: a component that wraps a model, ``op(model)``.  As in the
    reference (ConvolvedModel.__init__, :1917-1919) its parameters come AFTER those of the
    model it wraps."""

    type = 'conv'
    _kind = -1
    _supported = frozenset({'add', 'mul'})

    def __call__(self, model: Model) -> 'ConvolvedModel':
        if not isinstance(model, Model) or model.type not in self._supported:
            accepted = ', '.join(f"'{i}'" for i in sorted(self._supported))
            raise TypeError(
                f'{self.name} convolution model supports models with type: {accepted}; got '
                f"'{getattr(model, 'type', type(model).__name__)}' type model {model}"
            )
        return ConvolvedModel(self, model)

    def __add__(self, other):
        raise TypeError(f'{self.name} is a convolution model: apply it to a model first')

    __mul__ = __add__


class ConvolvedModel(Model):
    """models/model.py:1905-1959."""

    def __init__(self, op: ConvolutionComponent, model: Model):
        self.op, self.inner = op, model
        self.type = model.type

    def _leaves(self):
        out = list(self.inner._leaves())
        if all(self.op is not o for o in out):
            out.append(self.op)
        return out

    def __repr__(self):
        return f'{self.op.name}({self.inner!r})'


class _Shift(ConvolutionComponent):
    """Energy shifts (models/conv.py:259-432).  The shift parameter is fixed by default in the
    reference, and then the shift is a compile-time scale of the leaves' grids.  A FREE z / v (pass a
    ``UniformParameter``) is an ordinary column of theta: the kernels evaluate the wrapped components on
    dual-number energies, so d/dz comes with the other derivatives (``ElisaComponentDesc.shift_*``)."""

    _divide = False  # ZAShift divides the shifted model by the factor (conv.py:299-300)
    _velocity = False  # V*Shift: factor 1 - v / c

    @property
    def shift_param(self) -> UniformParameter:
        return self.params[self._config[0].name]

    def factor(self) -> float | None:
        """The grid factor of a FIXED shift; None when the shift parameter is free."""
        p = self.shift_param
        return None if not p.fixed else self._factor(p.default)


def _shift(name, supported, pcfg, factor, divide=False, velocity=False):
    return type(name, (_Shift,), {'_config': (ParamConfig(*pcfg),), '_supported': frozenset({supported}),
                                  '_factor': staticmethod(factor), '_divide': divide, '_velocity': velocity})


_C_KMS = 299792.458  # conv.py:383
ZAShift = _shift('ZAShift', 'add', ('z', 0.0, -0.999, 15.0, False, True), lambda z: 1.0 + z, divide=True)
ZMShift = _shift('ZMShift', 'mul', ('z', 0.0, -0.999, 15.0, False, True), lambda z: 1.0 + z)
VAShift = _shift('VAShift', 'add', ('v', 0.0, -1e4, 1e4, False, True), lambda v: 1.0 - v / _C_KMS, velocity=True)
VMShift = _shift('VMShift', 'mul', ('v', 0.0, -1e4, 1e4, False, True), lambda v: 1.0 - v / _C_KMS, velocity=True)


class _FluxNorm(ConvolutionComponent):
    """NormConvolution (models/conv.py:21-143): ``PhFlux(emin, emax, F=None, ngrid=1000, elog=True)``."""

    _supported = frozenset({'add'})

    def __init__(self, emin, emax, *args, ngrid=None, elog=None, **params):
        emin, emax = float(emin), float(emax)
        if emin >= emax:
            raise ValueError('emin must be less than emax')
        super().__init__(*args, **params)
        self.emin, self.emax = emin, emax
        self.ngrid = 1000 if ngrid is None else int(ngrid)
        self.elog = True if elog is None else bool(elog)
        if self.ngrid < 2:
            raise ValueError('ngrid must be at least 2')

    def flux_egrid(self) -> np.ndarray:
        """conv.py:81-84."""
        fn = np.geomspace if self.elog else np.linspace
        return np.ascontiguousarray(fn(self.emin, self.emax, self.ngrid), dtype=np.float64)


class PhFlux(_FluxNorm):  # conv.py:145-195
    _config = (ParamConfig('F', 1.0, 0.01, 1e10),)


class EnFlux(_FluxNorm):  # conv.py:198-256
    _config = (ParamConfig('F', 1e-12, 1e-30, 1e30, True),)


CONVOLUTION = (EnFlux, PhFlux, ZAShift, ZMShift, VAShift, VMShift)

ADDITIVE = (Band, BandEp, BentPL, Blackbody, BlackbodyRad, BrokenPL, Compt, CutoffPL, Gauss, Lorentz, OTTB, OTTS,
            PLEnFlux, PLPhFlux, PowerLaw, SmoothlyBrokenPL)
MULTIPLICATIVE = (Constant, Edge, ExpAbs, ExpFac, GAbs, HighECut, PLAbs)
TABLE_MULTIPLICATIVE = (PhAbs, TBAbs, WAbs)  # need a cross-section table argument


# --------------------------------------------------------------------------- compiled model
class CompiledModel:
    """Model.compile() result (models/model.py:293-303): the free parameters in order of
    first appearance (component order, then _config order) and the postfix program.

    ``eval(egrid, params)`` mirrors CompiledModel.eval (models/model.py:415-436): ``params``
    is an array (..., nparam), a mapping name -> array, or None (defaults); the result has
    shape (..., E) and is computed on the GPU by ``elisa_b200_unfold_dev``."""

    def __init__(self, model: Model):
        if not isinstance(model, Model):
            raise TypeError('model must be a Model')
        self.model = model
        self.type = model.type
        if model.type == 'conv':
            raise TypeError(f'{model} is a convolution model: apply it to a model first')
        # distinct components in the reference's order (naming, free parameters)
        self.leaves: list[Component] = model._leaves()
        # the program: one entry per (component, energy-shift context), flux normalisations, postfix ops
        # (component, grid_scale, out_scale, norm_grid_scale, norm_out_scale, in_norm, free shifts around it)
        self.entries: list[tuple] = []
        self.norms: list[_FluxNorm] = []
        self.ops: list[int] = self._emit(model, (1.0, 1.0, 1.0, 1.0, False, ()))
        if self.norms and any(e[6] for e in self.entries):
            raise NotImplementedError('a free energy shift together with a flux normalisation (PhFlux / EnFlux)')
        if len(self.entries) > _lib.MAX_COMPONENTS:
            raise NotImplementedError(f'more than {_lib.MAX_COMPONENTS} components in one model')
        # component display names: PowerLaw, PowerLaw_2, ... (models/model.py naming)
        seen: dict[str, int] = {}
        self.comp_names = []
        for c in self.leaves:
            seen[c.name] = seen.get(c.name, 0) + 1
            self.comp_names.append(c.name if seen[c.name] == 1 else f'{c.name}_{seen[c.name]}')
        self.free: list[UniformParameter] = []
        self.params_name: list[str] = []
        for cname, c in zip(self.comp_names, self.leaves):
            for pname in c.param_names:
                p = c.params[pname]
                if not p.fixed and all(p is not q for q in self.free):
                    self.free.append(p)
                    self.params_name.append(f'{cname}.{pname}')
        self.nparam = len(self.free)
        self.params_default = np.array([p.default for p in self.free], dtype=np.float64)
        self._plan = None
        self._plan_key = None

    def _emit(self, m: Model, ctx) -> list[int]:
        gs, os_, ngs, nos, in_norm, shifts = ctx
        if isinstance(m, ConvolvedModel):
            op = m.op
            if isinstance(op, _Shift):
                f = op.factor()
                if f is None:  # free z / v: the leaves below carry it as a dual-number factor
                    if len(shifts) >= _lib.MAX_SHIFTS:
                        raise NotImplementedError(f'more than {_lib.MAX_SHIFTS} nested free energy shifts')
                    return self._emit(m.inner, (gs, os_, ngs, nos, in_norm, shifts + (op,)))
                g = 1.0 / f if op._divide else 1.0
                return self._emit(m.inner, (gs * f, os_ * g, ngs * f if in_norm else ngs,
                                            nos * g if in_norm else nos, in_norm, shifts))
            if isinstance(op, _FluxNorm):
                if in_norm:
                    raise NotImplementedError('nested flux normalisations (PhFlux / EnFlux) are not supported')
                if len(self.norms) >= _lib.MAX_NORMS:
                    raise NotImplementedError(f'more than {_lib.MAX_NORMS} flux normalisations in one model')
                k = len(self.norms)
                self.norms.append(op)
                return self._emit(m.inner, (gs, os_, 1.0, 1.0, True, shifts)) + [_lib.OP_NORM0 - k]
            raise NotImplementedError(f'convolution component {op.name}')
        if isinstance(m, CompositeModel):
            return self._emit(m.lhs, ctx) + self._emit(m.rhs, ctx) + [_lib.OP_ADD if m.op == '+' else _lib.OP_MUL]
        if not isinstance(m, Component) or m.type == 'conv':
            raise TypeError(f'cannot compile {m!r}')
        # ZAShift's 1 / (1 + z) acts on the additive factor of each term of the shifted sum
        if shifts and isinstance(m, PhotonAbsorption):
            raise NotImplementedError('a free energy shift around a table component (PhAbs / TBAbs / WAbs)')
        # (shift kind, parameter): ZAShift's division acts on additive leaves only, like its fixed form
        free_shifts = tuple((_lib.SHIFT_V if op._velocity else (_lib.SHIFT_ZA if op._divide and m.type == 'add' else _lib.SHIFT_Z),
                             op.shift_param) for op in shifts)
        key = (m, gs, os_ if m.type == 'add' else 1.0, ngs, (nos if m.type == 'add' else 1.0), in_norm, free_shifts)
        for i, e in enumerate(self.entries):
            if e[0] is m and e[1:6] == key[1:6] and len(e[6]) == len(free_shifts) and all(
                    a[0] == b[0] and a[1] is b[1] for a, b in zip(e[6], free_shifts)):
                return [i]
        self.entries.append(key)
        return [len(self.entries) - 1]

    # -- descriptors ---------------------------------------------------------------
    def column_of(self, param: UniformParameter, free: Sequence[UniformParameter] | None = None) -> int:
        free = self.free if free is None else free
        for j, q in enumerate(free):
            if q is param:
                return j
        return -1

    def to_desc(self, free: Sequence[UniformParameter] | None = None):
        """Build the C descriptor; ``free`` is the global free-parameter list of a joint
        fit (columns of theta), defaulting to this model's own.  Returns (ModelDesc, keepalive)."""
        comps = (_lib.ComponentDesc * len(self.entries))()
        for i, (c, gs, os_, ngs, nos, _, free_shifts) in enumerate(self.entries):
            d = comps[i]
            d.grid_scale, d.out_scale, d.norm_grid_scale, d.norm_out_scale = gs, os_, ngs, nos
            for k, (kind, sp) in enumerate(free_shifts):
                j = self.column_of(sp, free)
                if j < 0:
                    raise ValueError(f'the free shift parameter {sp.name} is not among the fit parameters')
                d.shift_kind[k], d.shift_src[k] = kind, j
            d.kind = c._kind
            d.method = _lib.METHOD[c.method]
            for k in range(_lib.MAX_COMP_PARAMS):
                d.param_src[k] = -1
                d.param_const[k] = 0.0
            for k, pname in enumerate(c.param_names):
                p = c.params[pname]
                if p.fixed:
                    d.param_const[k] = p.default
                else:
                    j = self.column_of(p, free)
                    if j < 0:
                        raise ValueError(f'{c.name}.{pname} is free but not among the fit parameters')
                    d.param_src[k] = j
            d.aux[0], d.aux[1] = c.aux
            if isinstance(c, PhotonAbsorption):
                d.table_energy = c.table_log_energy.ctypes.data_as(_lib.c_double_p)
                d.table_xsect = c.table_log_xsect.ctypes.data_as(_lib.c_double_p)
                d.table_len = c.table_log_energy.size
                d.table_log = 1
        ops = (_lib.C.c_int32 * len(self.ops))(*self.ops)
        norms = (_lib.NormDesc * max(len(self.norms), 1))()
        grids = []
        for k, nm in enumerate(self.norms):
            p = nm.params['F']
            norms[k].kind = _lib.NORM[type(nm).__name__]
            norms[k].f_src = -1 if p.fixed else self.column_of(p, free)
            if not p.fixed and norms[k].f_src < 0:
                raise ValueError(f'{nm.name}.F is free but not among the fit parameters')
            norms[k].f_const = p.default
            g = nm.flux_egrid()
            grids.append(g)
            norms[k].grid = g.ctypes.data_as(_lib.c_double_p)
            norms[k].n_grid = g.size
        md = _lib.ModelDesc()
        md.n_components = len(self.entries)
        md.components = comps
        md.n_ops = len(self.ops)
        md.ops = ops
        md.n_norms = len(self.norms)
        md.norms = norms
        return md, (comps, ops, norms, grids)

    def to_spec(self, free: Sequence[UniformParameter] | None = None):
        """Neutral description (nested tuples) of the same program, the format the test
        oracle consumes: ('comp', name, {pname: ('theta', j) | ('const', v)}, extras),
        ('+', lhs, rhs), ('*', lhs, rhs)."""

        def prefs(m):
            return {pn: (('const', m.params[pn].default) if m.params[pn].fixed
                         else ('theta', self.column_of(m.params[pn], free))) for pn in m.param_names}

        def rec(m):
            if isinstance(m, ConvolvedModel):
                op = m.op
                extra = {}
                if isinstance(op, _FluxNorm):
                    extra = {'emin': op.emin, 'emax': op.emax, 'ngrid': op.ngrid, 'elog': op.elog}
                return ('conv', type(op).__name__, prefs(op), extra, rec(m.inner))
            if isinstance(m, Component):
                refs = {}
                for pname in m.param_names:
                    p = m.params[pname]
                    refs[pname] = ('const', p.default) if p.fixed else ('theta', self.column_of(p, free))
                extra = {'method': m.method}
                if m._kind in (PLEnFlux._kind, PLPhFlux._kind):
                    extra.update(emin=m.aux[0], emax=m.aux[1])
                if isinstance(m, PhotonAbsorption):
                    extra.update(table_energy=m.table_energy, table_xsect=m.table_xsect)
                    return ('comp', 'PhotonAbsorption', refs, extra)
                return ('comp', m.name, refs, extra)
            return (m.op, rec(m.lhs), rec(m.rhs))

        return rec(self.model)

    # -- evaluation ----------------------------------------------------------------
    def _theta_from(self, params):
        if params is None:
            return self.params_default.copy()
        if isinstance(params, dict):
            missing = [n for n in self.params_name if n not in params]
            if missing:
                raise ValueError(f'missing parameters: {", ".join(missing)}')
            arrs = [np.asarray(params[n], dtype=np.float64) for n in self.params_name]
            if not arrs:
                raise ValueError('params are empty')
            if any(a.shape != arrs[0].shape for a in arrs[1:]):
                raise ValueError('all params must have the same shape')
            return np.stack(arrs, axis=-1)
        theta = np.atleast_1d(np.asarray(params, dtype=np.float64))
        if theta.shape[-1] != self.nparam:
            raise ValueError(f'expected params shape (..., {self.nparam}), got {theta.shape}')
        return theta

    def eval(self, egrid, params=None, device=None):
        """CompiledModel.eval (models/model.py:415-436) on the GPU; returns a numpy array
        of shape (..., E) (no clip: the clip belongs to the likelihood, likelihood.py:204)."""
        from .likelihood import unfold_only

        theta = self._theta_from(params)
        batch = theta.shape[:-1]
        out = unfold_only(self, np.asarray(egrid, dtype=np.float64), theta.reshape(-1, max(self.nparam, 1)), device)
        return out.reshape(*batch, out.shape[-1])
