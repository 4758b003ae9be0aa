"""microclimc_b200 — B200-native batched implementation of microclimc's canopy energy-balance hot path."""
__version__ = "0.1.0"
