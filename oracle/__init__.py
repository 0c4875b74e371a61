"""CPU float64 oracle for the collapsed sparse-GP hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain torch-CPU float64, what the reference
(vr308/Generalised-Gaussian-Processes) evaluates at its call sites into
gpytorch / pymc3 / gpflow (none of which is vendored under /root/reference, none
of which is installed in the build container, and none of which is version
pinned by the reference).

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this
path and its third-party evaluators cannot be imported here, so the oracle is
pinned only against (a) the dense textbook definitions (N x N multivariate
normal log-density, unwhitened SVGP ELBO, brute-force quadrature) and (b) the
real RNG objects that *are* available (numpy RandomState, torch DataLoader).
See DESIGN.md section "Oracle".

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product path
(generalised-gaussian-processes_b200/) never does.
"""
from . import kernels, linalg, sgpr, priors, svgp, sgpmc  # noqa: F401
