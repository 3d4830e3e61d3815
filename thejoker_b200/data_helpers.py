"""validate_prepare_data (thejoker/data_helpers.py:56-133): one RVData, or several
surveys concatenated with an id per point and the trend design matrix."""
from __future__ import annotations

import numpy as np

from . import units as u
from .data import RVData
from .likelihood_helpers import get_trend_design_matrix

__all__ = ["validate_prepare_data"]


def validate_prepare_data(data, poly_trend, n_offsets):
    """Returns (RVData, survey id per point, trend_M).  A single RVData passes through;
    several (list or dict, one per survey) are merged into one time-sorted RVData with
    the rv unit of the first survey."""
    if isinstance(data, RVData):
        if n_offsets != 0:
            raise ValueError("If sampling over velocity offsets between data sources, you must "
                             "pass in multiple data sources. To do this, pass in a list of RVData "
                             "instances or a dictionary with RVData instances as values.")
        return data, np.zeros(len(data), dtype=int), get_trend_design_matrix(data, None, poly_trend)

    if hasattr(data, "keys"):
        surveys = [(k, data[k]) for k in data.keys()]
    else:
        try:
            surveys = list(enumerate(data))
        except TypeError:
            raise TypeError("Failed to parse input data: data must either be an RVData instance, "
                            "an iterable of RVData instances, or a dictionary with RVData "
                            f"instances as values. Received: {type(data)}")
    for key, d in surveys:
        if not isinstance(d, RVData):
            raise TypeError(f"All data must be specified as RVData instances: Object at key "
                            f"'{key}' is a '{type(d)}' instead.")
        if d._has_cov:
            raise NotImplementedError("We currently don't support multi-survey data when a full "
                                      "covariance matrix is specified.")
    if len(surveys) - 1 != n_offsets or len({k for k, _ in surveys}) != len(surveys):
        raise ValueError("Number of data IDs + 1 must equal the number of priors on constant "
                         "offsets specified (i.e. v0_offsets)")

    rv_unit = surveys[0][1].rv.unit
    t = np.concatenate([d._t_bmjd for _, d in surveys])
    rv = np.concatenate([d.rv.to_value(rv_unit) for _, d in surveys])
    err = np.concatenate([d.rv_err.to_value(rv_unit) for _, d in surveys])
    ids = np.concatenate([np.full(len(d), k) for k, d in surveys])

    # The reference builds the combined RVData (which sorts by time) but leaves `ids`
    # in concatenation order (data_helpers.py:117-131), so interleaved surveys get the
    # wrong indicator columns.  Here ids follow the same time sort as the data.
    order = np.argsort(t, kind="stable")
    all_data = RVData(t=t[order], rv=u.Quantity(rv[order], rv_unit),
                      rv_err=u.Quantity(err[order], rv_unit))
    ids = ids[order]
    return all_data, ids, get_trend_design_matrix(all_data, ids, poly_trend)
