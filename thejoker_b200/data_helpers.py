"""validate_prepare_data (thejoker/data_helpers.py:56-133): one RVData, or several
surveys concatenated with an id per point and the trend design matrix."""
from __future__ import annotations

import numpy as np

from . import units as u
from .data import RVData
from .likelihood_helpers import get_trend_design_matrix

__all__ = ["validate_prepare_data"]


def validate_prepare_data(data, poly_trend, n_offsets):
    if isinstance(data, RVData):
        if n_offsets != 0:
            raise ValueError("If sampling over velocity offsets between data sources, you must "
                             "pass in multiple data sources. To do this, pass in a list of RVData "
                             "instances or a dictionary with RVData instances as values.")
        trend_M = get_trend_design_matrix(data, None, poly_trend)
        return data, np.zeros(len(data), dtype=int), trend_M

    if not hasattr(data, "keys"):
        try:
            data = {i: d for i, d in enumerate(data)}
        except Exception:
            raise TypeError("Failed to parse input data: data must either be an RVData instance, "
                            "an iterable of RVData instances, or a dictionary with RVData "
                            f"instances as values. Received: {type(data)}")

    rv_unit = None
    t, rv, err, ids = [], [], [], []
    for k in data.keys():
        d = data[k]
        if not isinstance(d, RVData):
            raise TypeError(f"All data must be specified as RVData instances: Object at key "
                            f"'{k}' is a '{type(d)}' instead.")
        if d._has_cov:
            raise NotImplementedError("We currently don't support multi-survey data when a full "
                                      "covariance matrix is specified.")
        if rv_unit is None:
            rv_unit = d.rv.unit
        t.append(d._t_bmjd)
        rv.append(d.rv.to_value(rv_unit))
        err.append(d.rv_err.to_value(rv_unit))
        ids.append([k] * len(d))

    t = np.concatenate(t)
    rv = np.concatenate(rv)
    err = np.concatenate(err)
    ids = np.concatenate(ids)

    if (len(np.unique(ids)) - 1) != n_offsets:
        raise ValueError("Number of data IDs + 1 must equal the number of priors on constant "
                         "offsets specified (i.e. v0_offsets)")

    # The reference builds the combined RVData (which sorts by time) but leaves `ids`
    # in concatenation order (data_helpers.py:117-131), so interleaved surveys get the
    # wrong indicator columns.  Here ids follow the same time sort as the data.
    order = np.argsort(t, kind="stable")
    all_data = RVData(t=t[order], rv=u.Quantity(rv[order], rv_unit),
                      rv_err=u.Quantity(err[order], rv_unit))
    ids = ids[order]
    trend_M = get_trend_design_matrix(all_data, ids, poly_trend)
    return all_data, ids, trend_M
