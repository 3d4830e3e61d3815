"""Stand-in for thejoker/logging.py (which subclasses astropy's logger): a plain logger."""
import logging

logger = logging.getLogger("thejoker_reference_shim")
