"""Stand-in for the absent `astropy` package: just enough of `astropy.units` for the
module-level code and CJokerHelper.__init__ of thejoker/src/fast_likelihood.pyx
(pyx:13, 41-45, 133-166, 236-242).  TEST INFRASTRUCTURE ONLY; installed in sys.modules
only while oracle/ref_cython.py loads / constructs the compiled reference helper."""
