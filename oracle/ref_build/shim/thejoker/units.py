"""Stand-in for thejoker/units.py: only the attribute name the pyx reads (pyx:20, 213, 232)."""
UNIT_ATTR_NAME = "__ref_shim_unit__"
