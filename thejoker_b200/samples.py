"""JokerSamples: prior / posterior samples as named columns with units.

Keeps the slice of thejoker/samples.py the hot path uses: column access, ``pack`` /
``unpack`` (samples.py:404-478), ``wrap_K`` (:392-401), ``median_period`` (:378-384),
``mean/std``, ``t_ref / poly_trend / n_offsets`` metadata, ``par_names``, ``copy``,
indexing, ``ln_unmarginalized_likelihood`` (:611-632) and a simple ``write/read``.
The reference wraps an astropy QTable and writes HDF5 / FITS; astropy and h5py are
not in the target image, so storage here is a numpy ``.npz`` (one array per column =
the SoA layout the GPU wants).  Orbit objects (twobody) are out of scope.
"""
from __future__ import annotations

import copy as _copy
from collections import OrderedDict

import numpy as np

from . import units as u
from .prior import (get_linear_equiv_units, get_nonlinear_equiv_units,
                    get_v0_offsets_equiv_units, validate_n_offsets, validate_poly_trend)

__all__ = ["JokerSamples"]

# pyx:41-45
_nonlinear_packed_order = ["P", "e", "omega", "M0", "s"]
_nonlinear_internal_units = {"P": u.day, "e": u.one, "omega": u.radian, "M0": u.radian}


def file_format(filename):
    """'npz' | 'hdf5' | None from the first bytes of a file, whatever its name: this
    package's ``JokerSamples.write`` produces a zip (.npz) container, the reference's an
    HDF5 file (signature at offset 0 or, after a user block, at 512, 1024, ...)."""
    with open(filename, "rb") as f:
        head = f.read(8)
        if head[:2] == b"PK":
            return "npz"
        off = 0
        while len(head) == 8:
            if head == b"\x89HDF\r\n\x1a\n":
                return "hdf5"
            off = 512 if off == 0 else off * 2
            if off > (1 << 26):
                break
            f.seek(off)
            head = f.read(8)
    return None


class JokerSamples:
    _hdf5_path = "samples"

    def __init__(self, samples=None, t_ref=None, n_offsets=None, poly_trend=None, **kwargs):
        poly_trend = 1 if poly_trend is None else poly_trend
        n_offsets = 0 if n_offsets is None else n_offsets
        if isinstance(samples, JokerSamples):
            t_ref = samples.t_ref if t_ref is None else t_ref
            poly_trend, n_offsets = samples.poly_trend, samples.n_offsets
            samples = samples.tbl
        poly_trend, _ = validate_poly_trend(poly_trend)
        n_offsets, _ = validate_n_offsets(n_offsets)
        valid_units = {**get_nonlinear_equiv_units(), **get_linear_equiv_units(poly_trend),
                       **get_v0_offsets_equiv_units(n_offsets)}
        valid_units["ln_prior"] = u.one
        valid_units["ln_likelihood"] = u.one
        valid_units["ln_posterior"] = u.one
        self._valid_units = valid_units
        self.tbl = OrderedDict()
        self.meta = {"t_ref": t_ref, "poly_trend": poly_trend, "n_offsets": n_offsets}
        self.meta.update(kwargs)
        self._uniform_s = False
        if samples is not None:
            for k in samples.keys():
                self[k] = samples[k]

    # -- dict-like ---------------------------------------------------------
    def __getitem__(self, key):
        if isinstance(key, str):
            return self.tbl[key]
        new = self.__class__(**self.meta)
        for k, v in self.tbl.items():
            new.tbl[k] = u.Quantity(np.atleast_1d(v.value[key]), v.unit)
        new._uniform_s = self._uniform_s
        return new

    def __setitem__(self, key, val):
        if key not in self._valid_units:
            raise ValueError(f"Invalid parameter name '{key}'. Must be one of: "
                             f"{list(self._valid_units)}")
        if not isinstance(val, u.Quantity):
            val = u.Quantity(val, u.one) if self._valid_units[key] == u.one else u.Quantity(val)
        expected = self._valid_units[key]
        if not val.unit.is_equivalent(expected):
            raise u.UnitsError(f"Units of '{key}' samples ({val.unit}) are not compatible with "
                               f"expected units ({expected})")
        val = u.Quantity(np.atleast_1d(val.value), val.unit)
        if self.tbl and len(val) != len(self):
            raise ValueError(f"Column '{key}' has length {len(val)}, expected {len(self)}")
        if key == "s":
            self._uniform_s = False
        self.tbl[key] = val

    def __len__(self):
        for v in self.tbl.values():
            return len(v)
        return 0

    def keys(self):
        return self.tbl.keys()

    def __contains__(self, key):
        return key in self.tbl

    def __repr__(self):
        return f'<JokerSamples [{", ".join(self.tbl.keys())}] ({len(self)} samples)>'

    @property
    def t_ref(self):
        return self.meta["t_ref"]

    @property
    def poly_trend(self):
        return self.meta["poly_trend"]

    @property
    def n_offsets(self):
        return self.meta["n_offsets"]

    @property
    def isscalar(self):
        """samples.py:263-268: True when the object holds a single, un-indexed sample (the
        reference backs that case with an astropy ``Row``)."""
        return bool(self.tbl) and all(np.ndim(v.value) == 0 for v in self.tbl.values())

    @property
    def par_names(self):
        return [k for k in self.tbl.keys() if k not in ("ln_prior", "ln_likelihood", "ln_posterior")]

    def copy(self):
        new = self.__class__(**_copy.deepcopy(self.meta))
        for k, v in self.tbl.items():
            new.tbl[k] = v.copy()
        new._uniform_s = self._uniform_s
        return new

    # -- statistics --------------------------------------------------------
    def _apply(self, func):
        new = self.__class__(**self.meta)
        for k, v in self.tbl.items():
            new.tbl[k] = u.Quantity(np.atleast_1d(func(v.value)), v.unit)
        return new

    def mean(self):
        return self._apply(np.mean)

    def std(self):
        return self._apply(np.std)

    def median_period(self):
        """The sample at the median period (samples.py:378-384)."""
        P = self["P"].value
        idx = np.argpartition(P, len(P) // 2)[len(P) // 2]
        return self[idx]

    def wrap_K(self):
        """Flip negative K to positive and rotate omega by pi (samples.py:392-401)."""
        K = self.tbl["K"].value
        mask = K < 0
        if np.any(mask):
            K[mask] = np.abs(K[mask])
            om = self.tbl["omega"]
            f = u.rad.to(om.unit)
            om.value[mask] = (om.value[mask] + np.pi * f) % (2 * np.pi * f)
        return self

    # -- packing -----------------------------------------------------------
    def pack(self, units=None, names=None, nonlinear_only=True):
        """samples.py:404-445: one float64 (n, len(names)) array, units stripped."""
        units = {} if units is None else dict(units)
        out_units = OrderedDict()
        for k, v in _nonlinear_internal_units.items():
            units.setdefault(k, v)
        if names is None:
            names = _nonlinear_packed_order if nonlinear_only else self.par_names
        arrs = []
        for name in names:
            unit = units.get(name, self.tbl[name].unit)
            arrs.append(self.tbl[name].to_value(unit))
            out_units[name] = u.as_unit(unit)
        return np.stack(arrs, axis=1), out_units

    def columns(self, units=None, names=None):
        """SoA view used by the device path: list of contiguous float64 arrays in
        the requested units (no (n,5) packing copy)."""
        units = {} if units is None else dict(units)
        for k, v in _nonlinear_internal_units.items():
            units.setdefault(k, v)
        names = _nonlinear_packed_order if names is None else names
        return [np.ascontiguousarray(self.tbl[n].to_value(units.get(n, self.tbl[n].unit)),
                                     dtype=np.float64) for n in names]

    @classmethod
    def unpack(cls, packed_samples, units, **kwargs):
        """samples.py:447-478."""
        packed_samples = np.array(packed_samples)
        nsamples, npars = packed_samples.shape
        samples = cls(**kwargs)
        for i, k in enumerate(list(units.keys())[:npars]):
            samples[k] = u.Quantity(packed_samples[:, i], units[k])
        return samples

    # -- storage -----------------------------------------------------------
    def write(self, output, overwrite=False, append=False):
        import os

        if append and os.path.exists(output):
            old = self.read(output)
            merged = self.__class__(**old.meta)
            for k in old.tbl:
                merged.tbl[k] = u.Quantity(
                    np.concatenate([old.tbl[k].value, self.tbl[k].to_value(old.tbl[k].unit)]),
                    old.tbl[k].unit)
            return merged.write(output, overwrite=True)
        if os.path.exists(output) and not overwrite:
            raise OSError(f"File {output} exists: use overwrite=True")
        if str(output).lower().endswith((".hdf5", ".h5", ".fits")):
            import warnings

            # the reference writes HDF5 / FITS under these names (samples.py:480-545); no
            # HDF5 / FITS writer exists here (h5py / astropy are not dependencies)
            warnings.warn(f"{output}: written in thejoker_b200's own .npz container, not "
                          "HDF5 / FITS -- thejoker_b200 reads it back under any name (the format "
                          "is detected from the file's first bytes); the reference cannot",
                          UserWarning, stacklevel=2)
        payload = {f"col:{k}": v.value for k, v in self.tbl.items()}
        payload["__units__"] = np.array([f"{k}={v.unit.scale!r}|{v.unit.dims!r}"
                                         for k, v in self.tbl.items()])
        payload["__meta__"] = np.array([repr(self.meta)])
        with open(output, "wb") as f:
            np.savez(f, **payload)

    @classmethod
    def read(cls, filename, path=None):
        """samples.py:565-609.  ``path`` names the HDF5 group in the reference's files; the
        .npz container written by ``write`` has a single table, so it is ignored."""
        import ast

        kind = file_format(filename)
        if kind == "hdf5":  # a file written by the reference's JokerSamples.write
            from .cache import read_reference_hdf5

            return read_reference_hdf5(filename)
        if kind != "npz":
            raise OSError(f"{filename}: neither an .npz container written by JokerSamples.write "
                          "nor an HDF5 file written by the reference")
        with np.load(filename, allow_pickle=False) as z:
            meta = ast.literal_eval(str(z["__meta__"][0]))
            new = cls(**meta)
            for entry in z["__units__"]:
                k, spec = str(entry).split("=", 1)
                scale, dims = spec.split("|")
                unit = u.Unit(ast.literal_eval(dims), float(scale))
                new.tbl[k] = u.Quantity(z[f"col:{k}"], unit)
        return new

    # -- (unmarginalised) likelihood of posterior samples ----------------------
    def ln_unmarginalized_likelihood(self, data, helper=None, device=None):
        """samples.py:611-632, evaluated on the GPU (``tjb_unmarginalized_ll``; there is
        no host implementation).

        Without ``helper`` the model is the reference's: K z(t) plus the polynomial trend
        ``v0 + v1 dt + ...`` about ``t_ref`` -- survey offsets are not part of it, as in
        ``get_orbit`` (samples.py:296-343).  With a ``CJokerHelper`` built for the same
        data the helper's full design matrix (offsets included) is used."""
        if helper is not None:
            names = list(helper.internal_units.keys())[: 5 + helper.n_linear]
            rows, _ = self.pack(units=helper.internal_units, names=names, nonlinear_only=False)
            return helper.ln_unmarginalized_likelihood(rows)
        from .helper import CJokerHelper

        unit = data.rv.unit
        t_ref = self.t_ref if self.t_ref is not None else data._t_ref_bmjd
        dt = np.asarray(data._t_bmjd, dtype=float) - t_ref
        cols = [np.ones_like(dt)]
        for _ in range(1, self.poly_trend):
            cols.append(cols[-1] * dt)
        L = 1 + self.poly_trend
        spec = dict(t=data._t_bmjd, rv=data.rv.value, ivar=data.ivar.to_value(u.one / unit**2),
                    t0=t_ref, trend_M=np.stack(cols, axis=1), mu=np.zeros(L), Lambda=np.ones(L),
                    K_prior_kind=1, sigma_K0=1.0, P0=1.0, max_K=1.0, jitter_mode=1)
        n = len(self)
        rows = np.zeros((n, 5 + L))
        rows[:, 0] = self["P"].to_value(u.day)
        rows[:, 1] = self["e"].value
        rows[:, 2] = self["omega"].to_value(u.rad)
        rows[:, 3] = self["M0"].to_value(u.rad)
        rows[:, 4] = self["s"].to_value(unit) if "s" in self.tbl else 0.0
        rows[:, 5] = self["K"].to_value(unit)
        for j in range(self.poly_trend):
            rows[:, 6 + j] = self[f"v{j}"].to_value(unit / u.day**j)
        return CJokerHelper.from_spec(spec, device=device).ln_unmarginalized_likelihood(rows)
