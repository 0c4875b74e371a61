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


PIVOT_RTOL = 1e-12


def psd_safe_cholesky(A, policy="gpytorch", pivot_rtol=PIVOT_RTOL):
    """Returns (L, jitter_used).  A is a single [M,M] matrix.

    pivot_rtol: a factorisation whose smallest pivot L_jj^2 is <= pivot_rtol * max diag(A + jI) also counts as a failure.  LAPACK
    (hence upstream) only fails for a pivot <= 0, but with exactly duplicated inducing rows (the with-replacement draw of
    experiments/regression.py:83) the true pivot is 0 and the computed one is rounding noise of either sign -- LAPACK's outcome is a
    coin flip that no other implementation can reproduce, and proceeding on a noise pivot means cond ~ 1e16.  The CUDA Cholesky
    (csrc/kernel_tiles.cuh GGP_PIVOT_RTOL) applies the same deterministic rule, so both settle on the same ladder level.
    pivot_rtol=0 gives plain LAPACK semantics."""
    eye = torch.eye(A.shape[-1], dtype=A.dtype)
    last = None
    for j in jitter_ladder(policy, A.dtype):
        Aj = A if j == 0.0 else A + j * eye
        L, info = torch.linalg.cholesky_ex(Aj)
        if int(info) == 0 and pivot_rtol > 0 and float((torch.diagonal(L) ** 2).min()) <= pivot_rtol * float(torch.diagonal(Aj).max()):
            info = torch.argmin(torch.diagonal(L)) + 1
        if int(info) == 0:
            return L, j
        if torch.isnan(A).any():
            raise NotPSDError("NaN in matrix passed to Cholesky")
        last = int(info)
    raise NotPSDError(f"matrix not PD after jitter ladder {jitter_ladder(policy, A.dtype)}; leading minor {last}")
