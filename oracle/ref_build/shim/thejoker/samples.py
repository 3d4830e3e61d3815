"""Stand-in for thejoker/samples.py: likelihood_helpers.py:69-88 only calls
JokerSamples.unpack(raw, units, t_ref=, poly_trend=, n_offsets=) and then item-assigns
"ln_prior" / "ln_likelihood".  The harness wants the raw packed array back."""


class JokerSamples(dict):
    @classmethod
    def unpack(cls, raw_samples, units, t_ref=None, poly_trend=1, n_offsets=0):
        out = cls()
        out["raw"] = raw_samples
        out["poly_trend"], out["n_offsets"] = poly_trend, n_offsets
        return out
