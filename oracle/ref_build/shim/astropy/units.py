"""Minimal unit algebra: products of named bases with integer/fractional powers.
Conversions between different units are refused -- the harness hands the reference
helper values that are already in its internal units (day, rad, one rv unit)."""


class Unit:
    def __init__(self, bases=None):
        self.bases = {k: v for k, v in (bases or {}).items() if v != 0}

    def _combine(self, other, sign):
        if not isinstance(other, Unit):
            if other == 1:
                other = Unit()
            else:
                return NotImplemented
        out = dict(self.bases)
        for k, v in other.bases.items():
            out[k] = out.get(k, 0) + sign * v
        return Unit(out)

    def __mul__(self, other):
        return self._combine(other, +1)

    def __truediv__(self, other):
        return self._combine(other, -1)

    def __rtruediv__(self, other):
        if other != 1:
            return NotImplemented
        return Unit({k: -v for k, v in self.bases.items()})

    def __pow__(self, p):
        return Unit({k: v * p for k, v in self.bases.items()})

    def __eq__(self, other):
        return isinstance(other, Unit) and self.bases == other.bases

    def __hash__(self):
        return hash(tuple(sorted(self.bases.items())))

    def __repr__(self):
        return "Unit(%r)" % (self.bases,)


class Quantity:
    def __init__(self, value, unit):
        self.value, self.unit = value, unit

    def to_value(self, unit):
        if unit != self.unit:
            raise ValueError("shim Quantity cannot convert %r -> %r" % (self.unit, unit))
        return self.value


day = Unit({"d": 1})
radian = Unit({"rad": 1})
one = Unit()
