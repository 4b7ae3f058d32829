"""sobol_seq stand-in: the reference calls i4_sobol_generate (GP_Safe.py:211) and never uses the result."""
from scipy.stats import qmc


def i4_sobol_generate(dim, n, skip=1):
    return qmc.Sobol(d=dim, scramble=False).random(n + skip)[skip:]
