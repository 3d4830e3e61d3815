"""Stand-in for thejoker/utils.py::_pytensor_get_mean_std (utils.py:317-334), which
evaluates a pytensor Normal's (mu, sigma) and converts units.  The harness's fake
distributions carry plain floats already in the target unit."""


def _pytensor_get_mean_std(dist, in_unit, out_unit):
    if in_unit != out_unit:
        raise ValueError("shim cannot convert %r -> %r" % (in_unit, out_unit))
    return float(dist.mu), float(dist.sigma)
