"""A deliberately small unit system with the slice of the ``astropy.units`` surface
that The Joker's hot path touches (``5 * u.day``, ``q.to_value(u.km / u.s)``,
``q.unit``, ``q.value``, ``unit.is_equivalent``).

astropy is not available in the target image; if it is importable, astropy
``Quantity`` objects are accepted anywhere a quantity is expected (converted through
their ``to_value`` with a unit string).  The reference uses astropy throughout
(thejoker/data.py, prior.py, samples.py); only unit *conversion to the internal
units* ``[day, rad, rv-unit]`` matters for the hot path
(thejoker/src/fast_likelihood.pyx:41-45, 134-148).
"""
from __future__ import annotations

import numpy as np

# dimensions: (time, length, angle)
_BASE = {
    "day": ((1, 0, 0), 1.0),
    "d": ((1, 0, 0), 1.0),
    "s": ((1, 0, 0), 1.0 / 86400.0),
    "hour": ((1, 0, 0), 1.0 / 24.0),
    "year": ((1, 0, 0), 365.25),
    "yr": ((1, 0, 0), 365.25),
    "km": ((0, 1, 0), 1.0),
    "m": ((0, 1, 0), 1.0e-3),
    "rad": ((0, 0, 1), 1.0),
    "radian": ((0, 0, 1), 1.0),
    "deg": ((0, 0, 1), np.pi / 180.0),
    "": ((0, 0, 0), 1.0),
    "one": ((0, 0, 0), 1.0),
}


class UnitsError(ValueError):
    pass


class Unit:
    """scale * day^a km^b rad^c"""

    __array_priority__ = 1000

    def __init__(self, dims=(0, 0, 0), scale=1.0, name=None):
        self.dims = tuple(dims)
        self.scale = float(scale)
        self._name = name

    # -- algebra ----------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(tuple(a + b for a, b in zip(self.dims, other.dims)), self.scale * other.scale,
                        _join(self._name, other._name, " "))
        return Quantity(other, self)

    def __rmul__(self, other):
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(tuple(a - b for a, b in zip(self.dims, other.dims)), self.scale / other.scale,
                        _join(self._name, other._name, " / "))
        return Quantity(1.0 / np.asarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        return Quantity(other, Unit(tuple(-a for a in self.dims), 1.0 / self.scale,
                                    _join("1", self._name, " / ")))

    def __pow__(self, p):
        return Unit(tuple(a * p for a in self.dims), self.scale**p,
                    None if self._name is None else f"({self._name})^{p}")

    def __eq__(self, other):
        other = as_unit(other)
        return self.dims == other.dims and np.isclose(self.scale, other.scale, rtol=1e-14, atol=0)

    def __hash__(self):
        return hash((self.dims, round(self.scale, 12)))

    def is_equivalent(self, other):
        return self.dims == as_unit(other).dims

    def to(self, other, value=1.0):
        other = as_unit(other)
        if self.dims != other.dims:
            raise UnitsError(f"'{self}' and '{other}' are not convertible")
        factor = self.scale / other.scale
        value = np.asarray(value, dtype=float)
        # same unit: hand back the array itself (as astropy does) -- a 2^28-row prior
        # column must not be copied just to be multiplied by one
        return value if factor == 1.0 else value * factor

    def __repr__(self):
        return f"Unit('{self}')"

    def __str__(self):
        if self._name is not None:
            return self._name
        names = ("d", "km", "rad")
        parts = [f"{n}^{p}" if p != 1 else n for n, p in zip(names, self.dims) if p != 0]
        body = " ".join(parts)
        return (f"{self.scale:g} " if self.scale != 1.0 else "") + body


def _join(a, b, op):
    if a is None or b is None:
        return None
    if a == "":
        return b if op == " " else f"1{op}{b}"
    if b == "":
        return a
    return f"{a}{op}{b}"


class Quantity:
    __array_priority__ = 1001

    def __init__(self, value, unit=None):
        if isinstance(value, Quantity):
            unit = value.unit if unit is None else as_unit(unit)
            value = value.to_value(unit)
        elif _is_astropy_quantity(value):
            unit = as_unit(str(value.unit)) if unit is None else as_unit(unit)
            value = np.asarray(value.to_value(value.unit)) * as_unit(str(value.unit)).to(unit)
        self.value = np.asarray(value, dtype=float)
        self.unit = as_unit(unit)

    def to_value(self, unit=None):
        if unit is None:
            return self.value
        return self.unit.to(unit, self.value)

    def to(self, unit):
        return Quantity(self.to_value(unit), as_unit(unit))

    @property
    def shape(self):
        return self.value.shape

    @property
    def ndim(self):
        return self.value.ndim

    @property
    def size(self):
        return self.value.size

    @property
    def isscalar(self):
        return self.value.ndim == 0

    def __len__(self):
        return len(self.value)

    def __getitem__(self, idx):
        return Quantity(self.value[idx], self.unit)

    def __setitem__(self, idx, val):
        self.value[idx] = Quantity(val, self.unit).value if isinstance(val, Quantity) else val

    def __iter__(self):
        for v in self.value:
            yield Quantity(v, self.unit)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit * other)
        if isinstance(other, Quantity):
            return Quantity(self.value * other.value, self.unit * other.unit)
        return Quantity(self.value * np.asarray(other), self.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit / other)
        if isinstance(other, Quantity):
            return Quantity(self.value / other.value, self.unit / other.unit)
        return Quantity(self.value / np.asarray(other), self.unit)

    def __rtruediv__(self, other):
        return Quantity(np.asarray(other) / self.value, Unit() / self.unit)

    def __pow__(self, p):
        return Quantity(self.value**p, self.unit**p)

    def __add__(self, other):
        o = other.to_value(self.unit) if isinstance(other, Quantity) else other
        return Quantity(self.value + o, self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        o = other.to_value(self.unit) if isinstance(other, Quantity) else other
        return Quantity(self.value - o, self.unit)

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __lt__(self, other):
        o = other.to_value(self.unit) if isinstance(other, Quantity) else other
        return self.value < o

    def __gt__(self, other):
        o = other.to_value(self.unit) if isinstance(other, Quantity) else other
        return self.value > o

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value, dtype=dtype)

    def copy(self):
        return Quantity(self.value.copy(), self.unit)

    def __repr__(self):
        return f"<Quantity {self.value!r} {self.unit}>"


def _is_astropy_quantity(x):
    return type(x).__module__.startswith("astropy.") and hasattr(x, "to_value") and hasattr(x, "unit")


def as_unit(u):
    """Unit from a Unit, None (dimensionless), a string like 'km / s', or an astropy unit."""
    if isinstance(u, Unit):
        return u
    if u is None:
        return one
    s = str(u).strip()
    return _parse(s)


def _parse(s):
    # tiny parser: products / quotients / integer powers of base names
    s = s.replace("**", "^")
    num, _, den = s.partition("/")

    def prod(txt, sign):
        out = Unit()
        for tok in txt.replace("(", " ").replace(")", " ").replace("*", " ").split():
            if tok == "1":
                continue
            name, _, p = tok.partition("^")
            p = int(p) if p else 1
            for suffix in ("2", "3"):  # 's2' style
                if name not in _BASE and name.endswith(suffix) and name[:-1] in _BASE:
                    name, p = name[:-1], int(suffix)
            if name not in _BASE:
                raise UnitsError(f"unknown unit '{tok}' in '{s}'")
            dims, sc = _BASE[name]
            out = out * (Unit(dims, sc) ** (sign * p))
        return out

    u_ = prod(num, 1)
    if den:
        for part in den.split("/"):
            u_ = u_ * prod(part, -1)
    u_._name = s
    return u_


def to_value(x, unit, default_unit=None):
    """Numeric value of x in `unit`.  Bare numbers are taken to be in `default_unit`
    (or already in `unit`)."""
    unit = as_unit(unit)
    if isinstance(x, Quantity):
        return x.to_value(unit)
    if _is_astropy_quantity(x):
        return Quantity(x).to_value(unit)
    x = np.asarray(x, dtype=float)
    if default_unit is None:
        return x
    return as_unit(default_unit).to(unit, x)


# the names the reference imports from astropy.units
one = Unit((0, 0, 0), 1.0, "")
dimensionless_unscaled = one
day = Unit((1, 0, 0), 1.0, "d")
d = day
s = Unit((1, 0, 0), 1.0 / 86400.0, "s")
hour = Unit((1, 0, 0), 1.0 / 24.0, "h")
year = Unit((1, 0, 0), 365.25, "yr")
yr = year
km = Unit((0, 1, 0), 1.0, "km")
m = Unit((0, 1, 0), 1.0e-3, "m")
rad = Unit((0, 0, 1), 1.0, "rad")
radian = rad
deg = Unit((0, 0, 1), np.pi / 180.0, "deg")
_BASE["h"] = _BASE["hour"]
