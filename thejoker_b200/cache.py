"""Prior-cache storage.

The reference keeps the prior cache in an HDF5 file written by ``JokerSamples.write``
(dataset ``samples`` of compound rows = AoS, units in a YAML header;
thejoker/samples.py:480-545, samples_helpers.py:35-272) and every pool worker re-reads
its slice with PyTables, converts units column by column and packs an (n, 5) array
(thejoker/utils.py:106-198) -- at 2^28 samples that path is I/O bound.

Native format here: a directory with one little-endian float64 ``.npy`` file per column
(SoA, exactly the layout the GPU kernel reads) already converted to the internal units
[day, -, rad, rad, rv-unit], plus ``meta.json``.  Columns are memory-mapped, so a shard
``[lo, hi)`` is read straight from the page cache / disk into the device upload without
touching the rest of the file, and a constant jitter column is stored as a scalar.

``read_reference_hdf5`` / ``convert_reference_hdf5`` read the reference's own HDF5 layout
through ``h5py`` when it is installed and through the package's dependency-free reader
``hdf5_min`` otherwise (h5py is not in the build image); the conversion is blockwise.
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import units as u
from .samples import JokerSamples

__all__ = ["write_prior_cache", "PriorCache", "read_reference_hdf5", "convert_reference_hdf5",
           "parse_table_column_meta", "read_batch",
           "read_batch_slice", "read_batch_idx", "read_random_batch"]

_COLS = ("P", "e", "omega", "M0", "s")
_INTERNAL = {"P": u.day, "e": u.one, "omega": u.rad, "M0": u.rad}


def write_prior_cache(samples: JokerSamples, path: str, rv_unit=None, overwrite=False):
    """Store prior samples as a native SoA cache directory."""
    rv_unit = u.as_unit(u.km / u.s if rv_unit is None else rv_unit)
    if os.path.exists(path):
        if not overwrite:
            raise OSError(f"{path} exists: use overwrite=True")
    os.makedirs(path, exist_ok=True)
    n = len(samples)
    meta = {"n": n, "rv_unit_scale": rv_unit.scale, "rv_unit_dims": list(rv_unit.dims),
            "poly_trend": samples.poly_trend, "n_offsets": samples.n_offsets, "columns": [],
            "s_const": None}
    for name in _COLS:
        unit = _INTERNAL.get(name, rv_unit)
        col = np.ascontiguousarray(samples[name].to_value(unit), dtype="<f8")
        if name == "s" and n and np.all(col == col[0]):
            meta["s_const"] = float(col[0])
            continue
        np.save(os.path.join(path, f"{name}.npy"), col)
        meta["columns"].append(name)
    if "ln_prior" in samples:
        np.save(os.path.join(path, "ln_prior.npy"),
                np.ascontiguousarray(samples["ln_prior"].value, dtype="<f8"))
        meta["columns"].append("ln_prior")
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump(meta, f)
    return path


class PriorCache:
    """Memory-mapped view of a native prior cache."""

    def __init__(self, path):
        self.path = path
        with open(os.path.join(path, "meta.json")) as f:
            self.meta = json.load(f)
        self.n = int(self.meta["n"])
        self.rv_unit = u.Unit(tuple(self.meta["rv_unit_dims"]), self.meta["rv_unit_scale"])
        self._mm = {c: np.load(os.path.join(path, f"{c}.npy"), mmap_mode="r")
                    for c in self.meta["columns"]}
        self.s_const = self.meta.get("s_const")

    def __len__(self):
        return self.n

    @property
    def has_ln_prior(self):
        return "ln_prior" in self._mm

    def columns(self, rv_unit=None, lo=0, hi=None):
        """[P, e, omega, M0, s] for rows [lo, hi) in internal units; ``s`` is a python
        float when the cache holds a constant jitter."""
        hi = self.n if hi is None else hi
        f = 1.0 if rv_unit is None else float(self.rv_unit.to(u.as_unit(rv_unit)))
        cols = [self._mm[c][lo:hi] for c in _COLS[:4]]
        if self.s_const is not None:
            s = self.s_const * f
        else:
            s = self._mm["s"][lo:hi]
            s = s * f if f != 1.0 else s
        return cols + [s]

    def ln_prior(self, idx=None):
        lp = self._mm["ln_prior"]
        return np.asarray(lp if idx is None else lp[np.asarray(idx)])

    def to_samples(self, lo=0, hi=None):
        hi = self.n if hi is None else hi
        out = JokerSamples(poly_trend=self.meta["poly_trend"], n_offsets=self.meta["n_offsets"])
        cols = self.columns(lo=lo, hi=hi)
        for name, c in zip(_COLS, cols):
            unit = _INTERNAL.get(name, self.rv_unit)
            out[name] = u.Quantity(np.full(hi - lo, c) if np.ndim(c) == 0 else np.array(c), unit)
        if self.s_const is not None:
            out._uniform_s = True
        if self.has_ln_prior:
            out["ln_prior"] = np.array(self._mm["ln_prior"][lo:hi])
        return out


# -- row-block readers with the reference's names and signatures (utils.py:106-245) ------
def _open_columns(prior_samples_file):
    """(getter(name) -> (array-like column, unit), n_rows) for a native cache directory or
    a file written by JokerSamples.write."""
    if os.path.isdir(prior_samples_file):
        cache = PriorCache(prior_samples_file)

        def get(name):
            unit = _INTERNAL.get(name, cache.rv_unit)
            if name == "ln_prior":
                return cache._mm[name], u.one
            if name == "s" and cache.s_const is not None:
                return np.broadcast_to(np.float64(cache.s_const), (cache.n,)), unit
            return cache._mm[name], unit

        return get, cache.n
    samples = JokerSamples.read(prior_samples_file)
    return (lambda name: (samples[name].value, samples[name].unit)), len(samples)


def _read_rows(prior_samples_file, columns, rows, units):
    get, _ = _open_columns(prior_samples_file)
    batch = None
    for i, name in enumerate(columns):
        col, unit = get(name)
        arr = np.asarray(col[rows], dtype=np.float64)
        if batch is None:
            batch = np.zeros((len(arr), len(columns)))
        batch[:, i] = arr
        if units is not None and name in units:
            batch[:, i] *= float(unit.to(u.as_unit(units[name])))
    return np.zeros((0, len(columns))) if batch is None else batch


def read_batch_slice(prior_samples_file, columns, slice, units=None):
    """Rows ``slice`` of the named columns as a plain (n, len(columns)) float64 array,
    converted to ``units`` where given (utils.py:168-198)."""
    return _read_rows(prior_samples_file, columns, slice, units)


def read_batch_idx(prior_samples_file, columns, idx, units=None):
    """The same for an index array (utils.py:201-228)."""
    return _read_rows(prior_samples_file, columns, np.asarray(idx), units)


def read_random_batch(prior_samples_file, columns, size, units=None, rng=None):
    """``size`` distinct random rows (utils.py:231-245)."""
    if rng is None:
        rng = np.random.default_rng()
    _, n = _open_columns(prior_samples_file)
    idx = rng.choice(n, size=size, replace=False)
    return read_batch_idx(prior_samples_file, columns, idx=idx, units=units)


def read_batch(prior_samples_file, columns, slice_or_idx, units=None, rng=None):
    """Single entry point (utils.py:106-165): a slice or (start, stop) tuple reads a
    contiguous block, an integer a random batch of that size, an array those rows."""
    if isinstance(slice_or_idx, tuple):
        slice_or_idx = slice(*slice_or_idx)
    if isinstance(slice_or_idx, slice):
        return read_batch_slice(prior_samples_file, columns, slice_or_idx, units=units)
    if isinstance(slice_or_idx, (int, np.integer)):
        return read_random_batch(prior_samples_file, columns, slice_or_idx, units=units, rng=rng)
    if isinstance(slice_or_idx, np.ndarray):
        return read_batch_idx(prior_samples_file, columns, slice_or_idx, units=units)
    raise ValueError("Invalid input for slice_or_idx: must be a slice, int, or numpy array.")


# -- the reference's HDF5 prior cache ------------------------------------------------------
def parse_table_column_meta(lines):
    """The YAML header astropy stores next to a serialised table -- dataset
    ``samples.__table_column_meta__`` of the reference's prior cache (written by
    ``write_table_hdf5(..., serialize_meta=True)``, thejoker/samples.py:535-545; read back
    by ``get_header_from_yaml`` in thejoker/utils.py:75-88 and samples.py:548-563), one
    line per array element -- as ``(units, columns, meta)``: column name -> unit string
    ('' = dimensionless), the column order, and the table's own meta (``poly_trend``,
    ``n_offsets``, ``t_ref`` ...).

    The header carries a column's unit twice when the table held Quantity columns: in its
    ``datatype`` entry and in ``meta.__serialized_columns__.<name>.unit`` (an
    ``!astropy.units.Unit`` node); the first wins, the second is the fallback.  astropy's
    custom tags are read as plain mappings, so astropy itself is not needed."""
    import yaml

    class _Loader(yaml.SafeLoader):
        pass

    def _plain(loader, suffix, node):
        if isinstance(node, yaml.MappingNode):
            return loader.construct_mapping(node, deep=True)
        if isinstance(node, yaml.SequenceNode):
            return loader.construct_sequence(node, deep=True)
        return loader.construct_scalar(node)

    _Loader.add_multi_constructor("!", _plain)
    text = "\n".join(ln.decode("utf-8") if isinstance(ln, bytes) else str(ln) for ln in lines)
    header = yaml.load(text, Loader=_Loader)
    if not isinstance(header, dict) or "datatype" not in header:
        raise ValueError("not an astropy table header: no 'datatype' list")
    meta = header.get("meta") or {}
    if isinstance(meta, list):  # !!omap -> list of (key, value) pairs
        meta = {k: v for k, v in meta}
    serialized = meta.pop("__serialized_columns__", None) or {}
    units, columns = {}, []
    for row in header["datatype"]:
        name = row["name"]
        columns.append(name)
        unit = row.get("unit")
        if unit is None:
            ser = serialized.get(name) or {}
            unit = ser.get("unit")
            if isinstance(unit, dict):
                unit = unit.get("unit")
        units[name] = "" if unit is None else str(unit)
    return units, columns, meta


def _meta_kwargs(meta):
    """poly_trend / n_offsets / t_ref of the stored table, as JokerSamples takes them
    (samples.py:68-73).  A serialised astropy Time comes back as its jd1 + jd2 pair."""
    kw = {}
    for k in ("poly_trend", "n_offsets"):
        if meta.get(k) is not None:
            kw[k] = int(meta[k])
    t_ref = meta.get("t_ref")
    if isinstance(t_ref, dict) and "jd1" in t_ref:
        t_ref = (float(t_ref["jd1"]) - 2400000.5) + float(t_ref.get("jd2", 0.0))  # -> MJD
    if isinstance(t_ref, (int, float)):
        kw["t_ref"] = float(t_ref)
    return kw


def _open_reference_hdf5(filename):
    """h5py when it is installed (every user of the reference has it), else the package's
    own minimal reader (hdf5_min: the structures the reference's writer produces -- version-0
    superblock, old-style groups, chunked compound dataset, fixed-string header)."""
    try:
        import h5py
    except ImportError:
        from . import hdf5_min

        return hdf5_min.File(filename, "r")
    return h5py.File(filename, "r")


def read_reference_hdf5(filename, lo=0, hi=None):
    """JokerSamples from rows [lo, hi) of a file written by the reference's
    ``JokerSamples.write`` (HDF5 dataset ``samples`` of compound rows; column units in the
    YAML header stored next to it, thejoker/samples.py:480-563, utils.py:75-103), through
    h5py or, without it, hdf5_min (the reference's workers read the same dataset with
    PyTables, utils.py:168-198)."""
    with _open_reference_hdf5(filename) as f:
        dset = f[JokerSamples._hdf5_path]
        units, columns, meta = parse_table_column_meta(
            f[JokerSamples._hdf5_path + ".__table_column_meta__"][()])
        hi = len(dset) if hi is None else hi
        data = dset[lo:hi]
    out = JokerSamples(**_meta_kwargs(meta))
    for col in data.dtype.names:
        if col in out._valid_units or col in ("ln_prior", "ln_likelihood"):
            out[col] = u.Quantity(np.ascontiguousarray(data[col], dtype=np.float64),
                                  u.as_unit(units.get(col, "")))
    return out


def convert_reference_hdf5(filename, path, rv_unit=None, overwrite=False, rows_per_block=1 << 22):
    """Convert the reference's HDF5 prior cache into a native SoA cache directory, block by
    block: the compound (AoS) rows are read ``rows_per_block`` at a time and scattered into
    memory-mapped per-column files in internal units, so a 2^28-row cache (10.7 GB of rows)
    needs ~170 MB of host memory at a time, not the whole table twice.  Returns ``path``."""
    from numpy.lib.format import open_memmap

    rv_unit = u.as_unit(u.km / u.s if rv_unit is None else rv_unit)
    if os.path.exists(path) and not overwrite:
        raise OSError(f"{path} exists: use overwrite=True")
    os.makedirs(path, exist_ok=True)
    with _open_reference_hdf5(filename) as f:
        dset = f[JokerSamples._hdf5_path]
        units, columns, tmeta = parse_table_column_meta(
            f[JokerSamples._hdf5_path + ".__table_column_meta__"][()])
        n = len(dset)
        missing = [c for c in _COLS[:4] if c not in columns]
        if missing:
            raise ValueError(f"{filename}: prior cache lacks the columns {missing}")
        names = [c for c in _COLS + ("ln_prior",) if c in columns]
        factor = {c: (1.0 if c == "ln_prior" else
                      float(u.as_unit(units.get(c, "")).to(_INTERNAL.get(c, rv_unit))))
                  for c in names}
        mm = {c: open_memmap(os.path.join(path, f"{c}.npy"), mode="w+", dtype="<f8", shape=(n,))
              for c in names}
        s_first, s_uniform = None, "s" in names
        for lo in range(0, n, rows_per_block):
            block = dset[lo:lo + rows_per_block]
            for c in names:
                col = np.asarray(block[c], dtype=np.float64)
                if factor[c] != 1.0:
                    col = col * factor[c]
                mm[c][lo:lo + len(col)] = col
                if c == "s" and s_uniform and len(col):
                    s_first = col[0] if s_first is None else s_first
                    s_uniform = bool(np.all(col == s_first))
        for m in mm.values():
            m.flush()
        del mm
    kw = _meta_kwargs(tmeta)
    meta = {"n": n, "rv_unit_scale": rv_unit.scale, "rv_unit_dims": list(rv_unit.dims),
            "poly_trend": kw.get("poly_trend", 1), "n_offsets": kw.get("n_offsets", 0),
            "columns": list(names), "s_const": None}
    if "s" not in names:
        meta["s_const"] = 0.0
    elif s_uniform and n:  # a constant jitter column is stored as a scalar, like write_prior_cache
        meta["s_const"] = float(s_first)
        meta["columns"].remove("s")
        os.remove(os.path.join(path, "s.npy"))
    with open(os.path.join(path, "meta.json"), "w") as f:
        json.dump(meta, f)
    return path
