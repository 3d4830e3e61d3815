"""JokerPrior: the prior specification the hot path consumes, without pymc.

The reference describes priors as pymc random variables with units attached
(thejoker/prior.py, distributions.py, units.py) and extracts plain numbers from them
in CJokerHelper.__init__ (thejoker/src/fast_likelihood.pyx:204-252).  pymc / pytensor
are not part of the target image, so the distributions here are small numpy classes
that carry exactly those numbers and can draw samples.  Kept: ``JokerPrior(pars,
poly_trend, v0_offsets)``, ``JokerPrior.default(...)``, ``.sample(size,
generate_linear, return_logprobs, rng)``, ``.par_names``, ``.par_units``,
``.poly_trend``, ``.n_offsets``, ``.pars``, ``.v0_offsets``.  Arbitrary pymc models
are out of scope.
"""
from __future__ import annotations

import numpy as np

from . import units as u

__all__ = ["JokerPrior", "Normal", "FixedCompanionMass", "UniformLog", "Uniform", "Angle", "Beta",
           "Kipping13Global", "Kipping13Long", "Kipping13Short", "Constant", "LogNormal"]


# distribution kinds of the library's prior sampler (include/thejoker_b200.h TJB_PRIOR_*)
PRIOR_CONSTANT, PRIOR_UNIFORM, PRIOR_UNIFORMLOG, PRIOR_BETA, PRIOR_LOGNORMAL, PRIOR_NORMAL = range(6)

# ----------------------------------------------------------------------------
# distributions


class Distribution:
    is_normal = False

    def __init__(self, name=None, unit=None):
        self.name = name
        self.unit = u.as_unit(unit)

    def draw(self, rng, size, **ctx):
        raise NotImplementedError

    def logp(self, value, **ctx):
        raise NotImplementedError

    def device_spec(self):
        """(kind, p0, p1) for the library's counter-based sampler (include/thejoker_b200.h
        TJB_PRIOR_*, csrc/prior_gen.cuh), or None: no device sampler for this family."""
        return None


class Normal(Distribution):
    """Independent Normal prior (the only kind allowed on linear parameters,
    prior.py:158-177)."""
    is_normal = True

    def __init__(self, name=None, mu=0.0, sigma=1.0, unit=None):
        super().__init__(name, unit)
        self.mu, self.sigma = float(mu), float(sigma)

    def mean_std(self, to_unit):
        """(mean, std) in to_unit -- what utils._pytensor_get_mean_std returns."""
        f = self.unit.to(to_unit)
        return self.mu * f, self.sigma * f

    def draw(self, rng, size, **ctx):
        return rng.normal(self.mu, self.sigma, size=size)

    def device_spec(self):
        return (PRIOR_NORMAL, self.mu, self.sigma) if type(self) is Normal else None

    def logp(self, value, **ctx):
        return -0.5 * (np.log(2 * np.pi * self.sigma**2) + ((value - self.mu) / self.sigma) ** 2)


class FixedCompanionMass(Normal):
    """K ~ N(mu, sigma_K), sigma_K = min(max_K, sigma_K0 (P/P0)^(-1/3) (1-e^2)^(-1/2))
    (distributions.py:103-152)."""

    def __init__(self, name="K", sigma_K0=None, P0=None, mu=0.0, max_K=None, unit=None):
        sigma_K0 = u.Quantity(sigma_K0) if not isinstance(sigma_K0, u.Quantity) else sigma_K0
        unit = sigma_K0.unit if unit is None else u.as_unit(unit)
        Distribution.__init__(self, name, unit)
        self.mu = float(mu)
        self.sigma = np.nan
        self._sigma_K0 = sigma_K0
        self._P0 = P0 if isinstance(P0, u.Quantity) else u.Quantity(P0, u.day)
        if max_K is None:
            max_K = 500.0 * u.km / u.s
        self._max_K = max_K if isinstance(max_K, u.Quantity) else u.Quantity(max_K, unit)

    def sigma_of(self, P_day, e):
        sK0 = self._sigma_K0.to_value(self.unit)
        P0 = self._P0.to_value(u.day)
        sig = sK0 * (P_day / P0) ** (-1 / 3) / np.sqrt(1 - e**2)
        return np.clip(sig, 0.0, self._max_K.to_value(self.unit))

    def mean_std(self, to_unit):
        return self.mu * self.unit.to(to_unit), np.nan

    def draw(self, rng, size, P_day=None, e=None, **ctx):
        return rng.normal(self.mu, self.sigma_of(P_day, e), size=size)

    def logp(self, value, P_day=None, e=None, **ctx):
        sig = self.sigma_of(P_day, e)
        return -0.5 * (np.log(2 * np.pi * sig**2) + ((value - self.mu) / sig) ** 2)


class UniformLog(Distribution):
    """p(x) ~ 1/x on (a, b) (distributions.py:17-51)."""

    def __init__(self, name=None, a=1.0, b=2.0, unit=None):
        super().__init__(name, unit)
        self.a, self.b = float(a), float(b)
        if not (0 < self.a < self.b):
            raise ValueError("a > 0 and a < b")

    def draw(self, rng, size, **ctx):
        fac = np.log(self.b) - np.log(self.a)
        return np.exp(rng.uniform(size=size) * fac + np.log(self.a))  # distributions.py:25-28

    def device_spec(self):
        return (PRIOR_UNIFORMLOG, self.a, self.b)

    def logp(self, value, **ctx):
        # normalised density of 1/x on (a,b).  (distributions.py:44-46 writes
        # ``-value - log(fac)``, missing the log; the density form is used here.)
        return -np.log(value) - np.log(np.log(self.b) - np.log(self.a))


class Uniform(Distribution):
    def __init__(self, name=None, lower=0.0, upper=1.0, unit=None):
        super().__init__(name, unit)
        self.lower, self.upper = float(lower), float(upper)

    def draw(self, rng, size, **ctx):
        return rng.uniform(self.lower, self.upper, size=size)

    def device_spec(self):
        return (PRIOR_UNIFORM, self.lower, self.upper)

    def logp(self, value, **ctx):
        return np.full(np.shape(value), -np.log(self.upper - self.lower))


class Angle(Uniform):
    """Uniform angle.  The reference uses pymc_ext.distributions.angle
    (prior.py:437, 469-472), which returns values in (-pi, pi]."""

    def __init__(self, name=None, unit=None):
        super().__init__(name, -np.pi, np.pi, u.rad if unit is None else unit)


class Beta(Distribution):
    def __init__(self, name=None, alpha=1.0, beta=1.0, unit=None):
        super().__init__(name, unit)
        self.alpha, self.beta = float(alpha), float(beta)

    def draw(self, rng, size, **ctx):
        return rng.beta(self.alpha, self.beta, size=size)

    def device_spec(self):
        return (PRIOR_BETA, self.alpha, self.beta)

    def logp(self, value, **ctx):
        from math import lgamma
        lB = lgamma(self.alpha) + lgamma(self.beta) - lgamma(self.alpha + self.beta)
        return (self.alpha - 1) * np.log(value) + (self.beta - 1) * np.log1p(-value) - lB


class Kipping13Long(Beta):  # distributions.py:155-160
    def __init__(self, name=None):
        super().__init__(name, 1.12, 3.09)


class Kipping13Short(Beta):  # distributions.py:163-168
    def __init__(self, name=None):
        super().__init__(name, 0.697, 3.27)


class Kipping13Global(Beta):  # distributions.py:171-176
    def __init__(self, name=None):
        super().__init__(name, 0.867, 3.03)


class Constant(Distribution):
    """Deterministic constant (the default jitter prior, prior.py:476-479)."""

    def __init__(self, name=None, value=0.0, unit=None):
        super().__init__(name, unit)
        self.value = float(value)

    def draw(self, rng, size, **ctx):
        return np.full(size, self.value)

    def device_spec(self):
        return (PRIOR_CONSTANT, self.value, 0.0)

    def logp(self, value, **ctx):
        return np.zeros(np.shape(value))


class LogNormal(Distribution):
    """exp(N(mu, sigma)); the usual non-trivial jitter prior in the reference docs."""

    def __init__(self, name=None, mu=0.0, sigma=1.0, unit=None):
        super().__init__(name, unit)
        self.mu, self.sigma = float(mu), float(sigma)

    def draw(self, rng, size, **ctx):
        return np.exp(rng.normal(self.mu, self.sigma, size=size))

    def device_spec(self):
        return (PRIOR_LOGNORMAL, self.mu, self.sigma)

    def logp(self, value, **ctx):
        lv = np.log(value)
        return -lv - 0.5 * (np.log(2 * np.pi * self.sigma**2) + ((lv - self.mu) / self.sigma) ** 2)


# ----------------------------------------------------------------------------
# helpers (prior_helpers.py)


def validate_poly_trend(poly_trend):
    try:
        poly_trend = int(poly_trend)
    except Exception:
        raise ValueError("poly_trend must be an integer that specifies the number of polynomial "
                         "(in time) trend terms to include in The Joker.")
    return poly_trend, [f"v{i}" for i in range(poly_trend)]


def validate_n_offsets(n_offsets):
    try:
        n_offsets = int(n_offsets)
    except Exception:
        raise ValueError("n_offsets must be an integer that specifies the number of v0 offset "
                         "parameters to include in The Joker.")
    return n_offsets, [f"dv0_{i}" for i in range(1, n_offsets + 1)]


def get_nonlinear_equiv_units():
    return {"P": u.day, "e": u.one, "omega": u.radian, "M0": u.radian, "s": u.m / u.s}


def get_linear_equiv_units(poly_trend):
    _, v_names = validate_poly_trend(poly_trend)
    return {"K": u.m / u.s, **{name: u.m / u.s / u.day**i for i, name in enumerate(v_names)}}


def get_v0_offsets_equiv_units(n_offsets):
    _, names = validate_n_offsets(n_offsets)
    return {name: u.m / u.s for name in names}


def validate_sigma_v(sigma_v, poly_trend, v_names):
    """prior_helpers.py:41-75."""
    if isinstance(sigma_v, u.Quantity) and sigma_v.isscalar:
        sigma_v = {"v0": sigma_v}
    elif isinstance(sigma_v, u.Quantity):
        raise ValueError("You must pass in a scalar value for sigma_v if passing in a single "
                         "quantity.")
    if hasattr(sigma_v, "keys"):
        for name in v_names:
            if name not in sigma_v.keys():
                raise ValueError("If specifying the standard-deviations of the polynomial trend "
                                 "parameter prior, you must pass in values for all parameter "
                                 f"names. Expected keys: {v_names}, received: {sigma_v.keys()}")
        return sigma_v
    try:
        if len(sigma_v) != poly_trend:
            raise ValueError("You must pass in a single sigma value for each velocity trend "
                             f"parameter: You passed in {len(sigma_v)} values, but "
                             f"poly_trend={poly_trend}")
        return {name: val for name, val in zip(v_names, sigma_v)}
    except TypeError:
        raise TypeError("Invalid input for velocity trend prior sigma values. This must either "
                        "be a scalar Quantity (if poly_trend=1) or an iterable of Quantity "
                        "objects (if poly_trend>1)")


# ----------------------------------------------------------------------------


class JokerPrior:
    """Prior over [P, e, omega, M0, s] (nonlinear) and [K, v0, v1.., dv0_1..] (linear,
    independent Normals).  See the module docstring for what is kept from
    thejoker/prior.py:40-180."""

    def __init__(self, pars=None, poly_trend=1, v0_offsets=None, model=None):
        if pars is None:
            raise ValueError("pars must be given (there is no pymc model context here)")
        if isinstance(pars, Distribution):
            pars = {pars.name: pars}
        else:
            try:
                pars = dict(pars)
            except Exception:
                try:
                    pars = {p.name: p for p in pars}
                except Exception as e:
                    raise ValueError("Invalid input parameters: The input `pars` must either be "
                                     "a dictionary, list, or a single distribution, not a "
                                     f"'{type(pars)}'.") from e

        self.poly_trend, self._v_trend_names = validate_poly_trend(poly_trend)
        if v0_offsets is None:
            v0_offsets = []
        try:
            v0_offsets = list(v0_offsets)
        except Exception as e:
            raise TypeError("Constant velocity offsets must be an iterable of Normal "
                            "distributions that define the priors on each offset term.") from e
        self.v0_offsets = v0_offsets
        pars.update({p.name: p for p in self.v0_offsets})

        self._nonlinear_equiv_units = get_nonlinear_equiv_units()
        self._linear_equiv_units = get_linear_equiv_units(self.poly_trend)
        self._v0_offsets_equiv_units = get_v0_offsets_equiv_units(self.n_offsets)
        self._all_par_unit_equiv = {**self._nonlinear_equiv_units, **self._linear_equiv_units,
                                    **self._v0_offsets_equiv_units}

        for name in self.par_names:
            if name not in pars:
                raise ValueError(f"Missing prior for parameter '{name}': you must specify a prior "
                                 "distribution for all parameters.")
            if not isinstance(pars[name], Distribution):
                raise TypeError(f"Invalid type for prior on parameter {name}: {type(pars[name])}")
            equiv = self._all_par_unit_equiv[name]
            if not pars[name].unit.is_equivalent(equiv):
                raise ValueError(f"Parameter '{name}' has an invalid unit: The units for this "
                                 f"parameter must be transformable to '{equiv}'")
        for name in list(self._linear_equiv_units) + list(self._v0_offsets_equiv_units):
            if not pars[name].is_normal:
                raise ValueError("Priors on the linear parameters (K, v0, etc.) must be "
                                 f"independent Normal distributions, not "
                                 f"'{type(pars[name]).__name__}' (for {name})")
        for name, p in pars.items():
            if p.name is None:
                p.name = name
        self.pars = pars
        self.model = model

    @classmethod
    def default(cls, P_min=None, P_max=None, sigma_K0=None, P0=1 * u.year, sigma_v=None, s=None,
                poly_trend=1, v0_offsets=None, model=None, pars=None):
        r"""The default prior (prior.py:182-284):
        p(P) ~ 1/P on (P_min, P_max); e ~ Beta(0.867, 3.03); omega, M0 uniform angles;
        s constant; K ~ FixedCompanionMass(sigma_K0, P0); v_i ~ Normal(0, sigma_v_i)."""
        pars = {} if pars is None else (dict(pars) if hasattr(pars, "keys")
                                        else {p.name: p for p in pars})
        out = {}
        # nonlinear (prior.py:410-493)
        if "e" not in pars:
            out["e"] = Kipping13Global("e")
        if "omega" not in pars:
            out["omega"] = Angle("omega")
        if "M0" not in pars:
            out["M0"] = Angle("M0")
        if "s" not in pars:
            if s is None:
                s = 0 * u.m / u.s
            if isinstance(s, Distribution):
                s.name = s.name or "s"
                out["s"] = s
            else:
                s = s if isinstance(s, u.Quantity) else u.Quantity(s, u.km / u.s)
                if not s.unit.is_equivalent(u.km / u.s):
                    raise u.UnitsError("Invalid unit for s: must be equivalent to km/s")
                out["s"] = Constant("s", float(s.value), s.unit)
        if "P" not in pars:
            if P_min is None or P_max is None:
                raise ValueError("If you are using the default period prior, you must pass in "
                                 "both P_min and P_max to set the period prior domain.")
            P_min = P_min if isinstance(P_min, u.Quantity) else u.Quantity(P_min, u.day)
            P_max = P_max if isinstance(P_max, u.Quantity) else u.Quantity(P_max, u.day)
            out["P"] = UniformLog("P", float(P_min.value), float(P_max.to_value(P_min.unit)),
                                  P_min.unit)
        # linear (prior.py:496-575)
        poly_trend, v_names = validate_poly_trend(poly_trend)
        if v_names and "v0" not in pars:
            sigma_v = validate_sigma_v(sigma_v, poly_trend, v_names)
        if "K" not in pars:
            if sigma_K0 is None or P0 is None:
                raise ValueError("If using the default prior form on K, you must pass in a "
                                 "variance scale (sigma_K0) and a reference period (P0)")
            out["K"] = FixedCompanionMass("K", sigma_K0=sigma_K0, P0=P0)
        for name in v_names:
            if name not in pars:
                sv = sigma_v[name]
                sv = sv if isinstance(sv, u.Quantity) else u.Quantity(sv, u.km / u.s)
                out[name] = Normal(name, 0.0, float(sv.value), sv.unit)
        out.update(pars)
        return cls(pars=out, poly_trend=poly_trend, v0_offsets=v0_offsets, model=model)

    @property
    def par_names(self):
        return (list(self._nonlinear_equiv_units) + list(self._linear_equiv_units)
                + list(self._v0_offsets_equiv_units))

    @property
    def par_units(self):
        return {name: p.unit for name, p in self.pars.items()}

    @property
    def n_offsets(self):
        return len(self.v0_offsets)

    def __repr__(self):
        return f'<JokerPrior [{", ".join(self.par_names)}]>'

    def __str__(self):
        return ", ".join(self.par_names)

    _NONLINEAR_INTERNAL = ("P", "e", "omega", "M0", "s")

    def device_generator(self, seed, rv_unit):
        """The nonlinear prior as the library's counter-based sampler sees it
        (SURVEY.md section 8 f2): a ``_lib.TjbPriorGen`` -- per parameter the distribution
        kind, its two numbers and the factor into the helper's internal units
        [day, -, rad, rad, rv_unit] -- or None if a parameter's family has no device
        sampler (the caller then samples on the host).  Sample i of the prior is a pure
        function of (seed, i), see csrc/prior_gen.cuh; like the reference's pm.draw path
        the stream is not numpy's, the distributions are the same."""
        from . import _lib

        gen = _lib.TjbPriorGen()
        to = {"P": u.day, "e": u.one, "omega": u.rad, "M0": u.rad, "s": rv_unit}
        for k, name in enumerate(self._NONLINEAR_INTERNAL):
            spec = self.pars[name].device_spec()
            if spec is None:
                return None
            gen.par[k].kind, gen.par[k].p0, gen.par[k].p1 = int(spec[0]), float(spec[1]), float(spec[2])
            gen.par[k].scale = float(self.pars[name].unit.to(to[name]))
        gen.seed = int(seed) & (2**64 - 1)
        return gen

    def ln_prior_rows(self, rows, rv_unit):
        """ln prior density of packed rows [P, e, omega, M0, s] given in internal units
        (what ``sample(..., return_logprobs=True)`` stores per sample, prior.py:388-400)."""
        rows = np.asarray(rows, dtype=np.float64).reshape(-1, 5)
        to = {"P": u.day, "e": u.one, "omega": u.rad, "M0": u.rad, "s": rv_unit}
        logp = np.zeros(len(rows))
        for k, name in enumerate(self._NONLINEAR_INTERNAL):
            p = self.pars[name]
            logp = logp + p.logp(rows[:, k] / float(p.unit.to(to[name])))
        return logp

    def sample_device(self, size, device, seed, rv_unit, return_logprobs=False, index0=0):
        """Materialise ``size`` prior samples (global indices index0 ...) on a GPU:
        ``([P, e, omega, M0] float64 CUDA tensors in [day, -, rad, rad], s, ln_prior)``
        with ``s`` a tensor in ``rv_unit`` or a python float for a constant jitter and
        ``ln_prior`` a host array or None.  None if a distribution has no device sampler.
        One hand-written kernel (``tjb_prior_sample``); the rejection sampler itself does
        not call this -- its likelihood kernel generates the same samples in registers."""
        import torch

        from .helper import prior_sample_device

        gen = self.device_generator(seed, rv_unit)
        if gen is None:
            return None
        dev = torch.device(device)
        const_s = self.pars["s"].device_spec()[0] == PRIOR_CONSTANT
        cols = prior_sample_device(gen, int(index0), int(size), dev.index or 0, with_s=not const_s)
        s = float(gen.par[4].p0 * gen.par[4].scale) if const_s else cols[4]
        logp = None
        if return_logprobs:
            rows = np.stack([c.cpu().numpy() for c in cols[:4]]
                            + [np.full(int(size), s) if const_s else cols[4].cpu().numpy()], axis=1)
            logp = self.ln_prior_rows(rows, rv_unit)
        return cols[:4], s, logp

    def sample(self, size=1, generate_linear=False, return_logprobs=False, rng=None, dtype=None,
               **kwargs):
        """Draw prior samples (prior.py:297-407) with a numpy Generator.

        Not stream-compatible with the reference (which draws through pymc.draw);
        the distributions are the same.  Draw order: P, e, omega, M0, s, then linear.
        """
        from .samples import JokerSamples

        if rng is None:
            rng = np.random.default_rng()
        elif not isinstance(rng, np.random.Generator):
            rng = np.random.default_rng(rng)
        dtype = np.float64 if dtype is None else dtype
        size = int(size)

        names = list(self._nonlinear_equiv_units)
        if generate_linear:
            names = self.par_names
        raw, ctx = {}, {}
        for name in names:
            p = self.pars[name]
            raw[name] = np.asarray(p.draw(rng, size, **ctx), dtype=dtype)
            if name == "P":
                ctx["P_day"] = raw["P"] * p.unit.to(u.day)
            if name == "e":
                ctx["e"] = raw["e"]

        samples = JokerSamples(poly_trend=self.poly_trend, n_offsets=self.n_offsets, **kwargs)
        for name in names:
            samples[name] = u.Quantity(np.atleast_1d(raw[name]), self.pars[name].unit)
        # constant jitter is the default; remember it so the drivers can skip a scan
        if isinstance(self.pars["s"], Constant):
            samples._uniform_s = True
        if return_logprobs:
            logp = np.zeros(size)
            for name in names:
                logp = logp + self.pars[name].logp(raw[name], **ctx)
            samples["ln_prior"] = logp
        return samples
