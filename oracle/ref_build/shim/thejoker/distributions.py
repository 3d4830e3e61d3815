"""Stand-in for thejoker/distributions.py: the pyx imports the name only (pyx:224)."""


class FixedCompanionMass:
    pass
