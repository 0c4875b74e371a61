"""B200-native collapsed sparse-GP hot path (SGPR bound + gradient, SVGP ELBO, sparse predictive).

Drop-in for the arithmetic vr308/Generalised-Gaussian-Processes delegates to gpytorch / pymc3 at
models/sgpr.py:123-129, models/bayesian_sgpr_hmc.py:60-78, models/svgp.py:104-110.  The directory name has a hyphen;
import it through the `ggp_b200` shim at the repository root.
"""
from ._lib import build, load, LIB_PATH, GgpError  # noqa: F401
from .engine import Engine, NotPSDError, jitter_ladder  # noqa: F401
