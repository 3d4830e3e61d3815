"""Empty parent package so that the compiled reference module can be imported under its
own qualified name, thejoker.src.fast_likelihood (its `from ..distributions import ...`
is relative).  Stand-in modules only; TEST INFRASTRUCTURE ONLY."""
