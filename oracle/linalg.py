"""Cholesky policies the reference's evaluators apply (oracle).

* gpytorch ``psd_safe_cholesky`` (every Cholesky under models/sgpr.py:123-125, models/svgp.py:104):
  try plain; on failure set the *total* diagonal jitter to j0*10^i, i=0..2 (j0 = 1e-8 in float64,
  1e-6 in float32) and retry; raise NotPSDError after three failures.
* pymc3 ``stabilize``: K + 1e-6*I unconditionally (models/bayesian_sgpr_hmc.py:66 -> MarginalSparse).
"""
import torch


class NotPSDError(RuntimeError):
    pass


def jitter_ladder(policy, dtype=torch.float64):
    """Return the list of total-jitter values to try, in order, for a policy.

    policy: "gpytorch" | "pymc3" | float (fixed, single attempt).
    """
    if policy == "gpytorch":
        j0 = 1e-8 if dtype == torch.float64 else 1e-6
        return [0.0, j0, j0 * 10.0, j0 * 100.0]
    if policy == "pymc3":
        return [1e-6]
    return [float(policy)]


def psd_safe_cholesky(A, policy="gpytorch"):
    """Returns (L, jitter_used).  A is a single [M,M] matrix."""
    eye = torch.eye(A.shape[-1], dtype=A.dtype)
    last = None
    for j in jitter_ladder(policy, A.dtype):
        Aj = A if j == 0.0 else A + j * eye
        L, info = torch.linalg.cholesky_ex(Aj)
        if int(info) == 0:
            return L, j
        if torch.isnan(A).any():
            raise NotPSDError("NaN in matrix passed to Cholesky")
        last = int(info)
    raise NotPSDError(f"matrix not PD after jitter ladder {jitter_ladder(policy, A.dtype)}; leading minor {last}")
