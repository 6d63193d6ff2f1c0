"""Random reference-named state dicts for the tests: re-exported from oracle/weights.py (bench.py's CPU legs use the
same generator without importing anything under tests/)."""
from oracle.weights import random_d_sd, random_g_sd  # noqa: F401
