"""Titsias collapsed SGPR bound, its gradient and the sparse predictive (oracle; torch CPU float64).

What the reference evaluates (it only wires third-party code):
  * models/sgpr.py:114,123-129      -mll(output, y)   with ExactMarginalLogLikelihood over
                                    InducingPointKernel(ScaleKernel(RBFKernel(ard)))  -> F / N
  * models/bayesian_sgpr_hmc.py:60-71  pm.gp.MarginalSparse(approx="VFE").marginal_likelihood -> F (jitter 1e-6)
  * models/sgpr.py:150-160          likelihood(self(test_x)) in eval mode             -> predictive
Formulas: SURVEY Appendix A.4 (bound), section 8a row R5 (closed-form gradient), A.6 (predictive).
"""
import math
import torch

from .kernels import ard_kernel
from .linalg import psd_safe_cholesky

LOG2PI = math.log(2.0 * math.pi)


def _tri_solve(L, B, upper=False):
    return torch.linalg.solve_triangular(L, B, upper=upper)


def sgpr_bound(X, y, Z, ell, sf2, s2, jitter_policy="gpytorch", normalize="n", kind="rbf",
               return_state=False):
    """Collapsed bound in the A/B form (pymc3's formulation).

    normalize="n"   -> F / N  (gpytorch ExactMarginalLogLikelihood divides by num_data; models/sgpr.py:125)
    normalize="none"-> F      (pymc3 MarginalSparse VFE logp; models/bayesian_sgpr_hmc.py:71)
    The jitter found by the policy is a constant w.r.t. autograd (as upstream).
    """
    N = X.shape[0]
    M = Z.shape[0]
    Kzz = ard_kernel(Z, Z, ell, sf2, kind)
    with torch.no_grad():
        _, jit = psd_safe_cholesky(Kzz.detach(), jitter_policy)
    L = torch.linalg.cholesky(Kzz + jit * torch.eye(M, dtype=X.dtype))
    Kzx = ard_kernel(Z, X, ell, sf2, kind)
    A = _tri_solve(L, Kzx)
    S = A @ A.T
    B = torch.eye(M, dtype=X.dtype) + S / s2
    LB = torch.linalg.cholesky(B)
    b = A @ y
    c = _tri_solve(LB, b.unsqueeze(-1)).squeeze(-1) / s2
    yty = y @ y
    F = (-0.5 * N * LOG2PI - 0.5 * N * torch.log(s2) - torch.log(torch.diagonal(LB)).sum()
         - 0.5 * (yty / s2 - c @ c) - 0.5 * (N * sf2 - torch.trace(S)) / s2)
    out = F / N if normalize == "n" else F
    if return_state:
        return out, dict(L=L, LB=LB, A=A, S=S, b=b, c=c, jitter=jit, Kzx=Kzx, Kzz=Kzz)
    return out


def sgpr_bound_gpytorch_form(X, y, Z, ell, sf2, s2, jitter_policy="gpytorch", kind="rbf"):
    """Same number reached the way gpytorch does: R = Kxz U^{-1} with an *explicit*
    inverse of the upper factor, Woodbury on the capacitance I + R^T R / s, plus the added-loss
    trace term; divided by N."""
    N, M = X.shape[0], Z.shape[0]
    Kzz = ard_kernel(Z, Z, ell, sf2, kind, gpytorch_order=True)
    L, jit = psd_safe_cholesky(Kzz, jitter_policy)
    U = L.T
    Uinv = _tri_solve(U, torch.eye(M, dtype=X.dtype), upper=True)
    Kxz = ard_kernel(X, Z, ell, sf2, kind, gpytorch_order=True)
    R = Kxz @ Uinv
    cap = torch.eye(M, dtype=X.dtype) + R.T @ R / s2
    Lc = torch.linalg.cholesky(cap)
    logdet = 2.0 * torch.log(torch.diagonal(Lc)).sum() + N * torch.log(s2)
    Rty = R.T @ y
    v = _tri_solve(Lc, Rty.unsqueeze(-1)).squeeze(-1)
    quad = (y @ y) / s2 - (v @ v) / (s2 * s2)
    log_prob = -0.5 * (quad + logdet + N * LOG2PI)
    q_diag = (R * R).sum(-1)
    added = 0.5 * ((q_diag - sf2) / s2).sum()
    return (log_prob + added) / N


def sgpr_bound_dense(X, y, Z, ell, sf2, s2, jitter=0.0, kind="rbf"):
    """Textbook definition, O(N^3): log N(y; 0, Qnn + s I) - tr(Knn - Qnn)/(2s).  Known-answer anchor."""
    N, M = X.shape[0], Z.shape[0]
    Kzz = ard_kernel(Z, Z, ell, sf2, kind) + jitter * torch.eye(M, dtype=X.dtype)
    Kxz = ard_kernel(X, Z, ell, sf2, kind)
    Qnn = Kxz @ torch.linalg.solve(Kzz, Kxz.T)
    cov = Qnn + s2 * torch.eye(N, dtype=X.dtype)
    cov = 0.5 * (cov + cov.T)
    mvn = torch.distributions.MultivariateNormal(torch.zeros(N, dtype=X.dtype), covariance_matrix=cov)
    return mvn.log_prob(y) - 0.5 * (N * sf2 - torch.trace(Qnn)) / s2


def sgpr_bound_and_grads_autograd(X, y, Z, ell, sf2, s2, jitter_policy="gpytorch", normalize="n", kind="rbf"):
    """Bound and d/d(ell, sf2, s2, Z) by torch autograd through sgpr_bound (what loss.backward()
    does at models/sgpr.py:129, w.r.t. the constrained values)."""
    ell = ell.detach().clone().requires_grad_(True)
    sf2 = sf2.detach().clone().requires_grad_(True)
    s2 = s2.detach().clone().requires_grad_(True)
    Z = Z.detach().clone().requires_grad_(True)
    F = sgpr_bound(X, y, Z, ell, sf2, s2, jitter_policy, normalize, kind)
    g = torch.autograd.grad(F, [ell, sf2, s2, Z])
    return F.detach(), dict(ell=g[0], sf2=g[1], s2=g[2], Z=g[3])


def sgpr_grads_closed_form(X, y, Z, ell, sf2, s2, jitter_policy="gpytorch", normalize="n", form="QA"):
    """Closed-form gradient of the RBF bound; the algebra the CUDA backward implements.

    beta = B^{-1} b ; P_A = (I - B^{-1})/s - beta beta^T / s^3 ; u = L^{-T} beta / s^2
    dF/dKzx = Q A + u y^T with Q = L^{-T} P_A and A = L^{-1} Kzx  (form="QA", the default);
    dF/dKzz = -1/2 L^{-T} (B + B^{-1} - 2I + beta beta^T / s^2) L^{-1}

    form="P" is SURVEY R5 as written, dF/dKzx = P Kzx + u y^T with P = L^{-T} P_A L^{-1}: the same number in exact arithmetic, but
    P has entries of size 1/lambda_min(Kzz) and P Kzx cancels down by cond(Kzz): in float64 it loses cond(Kzz) * eps
    (1e-8 .. 2e-6 on the gradient at the headline Kzz, cond 7.6e7, measured against oracle/hp in tests/test_oracle_hp.py), whereas
    Q A -- which is what reverse-mode autograd through the triangular solve computes, i.e. what the reference's loss.backward() does
    -- loses only cond(L) * eps (1e-11 .. 1e-9).  form="P" is kept for that demonstration only.
    """
    N, M = X.shape[0], Z.shape[0]
    I = torch.eye(M, dtype=X.dtype)
    F, st = sgpr_bound(X, y, Z, ell, sf2, s2, jitter_policy, "none", "rbf", return_state=True)
    L, LB, S, b, Kzx, Kzz = st["L"], st["LB"], st["S"], st["b"], st["Kzx"], st["Kzz"]
    Linv = _tri_solve(L, I)
    LBinv = _tri_solve(LB, I)
    Binv = LBinv.T @ LBinv
    beta = Binv @ b
    PA = (I - Binv) / s2 - torch.outer(beta, beta) / s2 ** 3
    u = Linv.T @ beta / s2 ** 2
    if form == "P":
        P = Linv.T @ PA @ Linv
        G = P @ Kzx + torch.outer(u, y)                               # dF/dKzx   [M,N]
    else:
        G = (Linv.T @ PA) @ st["A"] + torch.outer(u, y)
    Bm = I + S / s2
    Gbar = Bm + Binv - 2.0 * I + torch.outer(beta, beta) / s2 ** 2
    Gzz = -0.5 * Linv.T @ Gbar @ Linv                                 # dF/dKzz   [M,M] (symmetric)
    W = G * Kzx
    V = Gzz * Kzz
    dzx = Z.unsqueeze(1) - X.unsqueeze(0)                             # [M,N,D]
    dzz = Z.unsqueeze(1) - Z.unsqueeze(0)                             # [M,M,D]
    d_ell = ((W.unsqueeze(-1) * dzx ** 2).sum((0, 1)) + (V.unsqueeze(-1) * dzz ** 2).sum((0, 1))) / ell ** 3
    d_sf2 = (W.sum() + V.sum()) / sf2 - N / (2.0 * s2)
    d_Z = (-(W.unsqueeze(-1) * dzx).sum(1) - 2.0 * (V.unsqueeze(-1) * dzz).sum(1)) / ell ** 2
    trS = torch.trace(S)
    d_s2 = (-N / (2 * s2) + 0.5 * torch.trace(Binv @ S) / s2 ** 2 + 0.5 * (y @ y) / s2 ** 2
            - (b @ beta) / s2 ** 3 + (beta @ S @ beta) / (2 * s2 ** 4) + (N * sf2 - trS) / (2 * s2 ** 2))
    scale = 1.0 / N if normalize == "n" else 1.0
    return F * scale, dict(ell=d_ell * scale, sf2=d_sf2 * scale, s2=d_s2 * scale, Z=d_Z * scale)


def sgpr_bound_and_grads_chunked(X, y, Z, ell, sf2, s2, jitter_policy="gpytorch", normalize="n",
                                 chunk=65536, form="QA"):
    """Two-pass, N-chunked evaluation of bound + closed-form gradients (RBF).  Never holds an N x M
    buffer larger than chunk x M, so BASELINE config 4 (N=1e6, M=1024) fits in host RAM.  Used as the
    CPU baseline in bench.py and cross-checked against the unchunked oracle."""
    N, M = X.shape[0], Z.shape[0]
    dt = X.dtype
    I = torch.eye(M, dtype=dt)
    Kzz = ard_kernel(Z, Z, ell, sf2)
    L, jit = psd_safe_cholesky(Kzz, jitter_policy)
    Linv = _tri_solve(L, I)
    S = torch.zeros(M, M, dtype=dt)
    b = torch.zeros(M, dtype=dt)
    yty = torch.zeros((), dtype=dt)
    keep_A = form != "P" and N * M * 8 <= 16 * 2 ** 30   # A = L^{-1} Kzx of pass 1 is reused by pass 2 when it fits in host memory
    A_chunks = []
    for i0 in range(0, N, chunk):
        Xc, yc = X[i0:i0 + chunk], y[i0:i0 + chunk]
        A = Linv @ ard_kernel(Z, Xc, ell, sf2)
        if keep_A:
            A_chunks.append(A)
        S += A @ A.T
        b += A @ yc
        yty += yc @ yc
    Bm = I + S / s2
    LB = torch.linalg.cholesky(Bm)
    c = _tri_solve(LB, b.unsqueeze(-1)).squeeze(-1) / s2
    F = (-0.5 * N * LOG2PI - 0.5 * N * torch.log(s2) - torch.log(torch.diagonal(LB)).sum()
         - 0.5 * (yty / s2 - c @ c) - 0.5 * (N * sf2 - torch.trace(S)) / s2)
    LBinv = _tri_solve(LB, I)
    Binv = LBinv.T @ LBinv
    beta = Binv @ b
    PA = (I - Binv) / s2 - torch.outer(beta, beta) / s2 ** 3
    QA = Linv.T @ PA
    QL = QA @ Linv if form == "P" else None
    u = Linv.T @ beta / s2 ** 2
    Gbar = Bm + Binv - 2.0 * I + torch.outer(beta, beta) / s2 ** 2
    Gzz = -0.5 * Linv.T @ Gbar @ Linv
    # moments of W = G o Kzx against [1, x, x^2]
    r = torch.zeros(M, dtype=dt)
    Q = torch.zeros(M, Z.shape[1], dtype=dt)
    T = torch.zeros(M, Z.shape[1], dtype=dt)
    for i0 in range(0, N, chunk):
        Xc, yc = X[i0:i0 + chunk], y[i0:i0 + chunk]
        Kc = ard_kernel(Z, Xc, ell, sf2)
        if form == "P":                                             # see sgpr_grads_closed_form on the two forms
            G = QL @ Kc
        else:
            G = QA @ (A_chunks[i0 // chunk] if keep_A else Linv @ Kc)
        W = (G + torch.outer(u, yc)) * Kc
        r += W.sum(1)
        Q += W @ Xc
        T += W @ (Xc * Xc)
    V = Gzz * Kzz
    rv = V.sum(1)
    Qv = V @ Z
    Tv = V @ (Z * Z)
    d_ell = ((Z * Z * r[:, None] - 2 * Z * Q + T).sum(0) + (Z * Z * rv[:, None] - 2 * Z * Qv + Tv).sum(0)) / ell ** 3
    d_sf2 = (r.sum() + rv.sum()) / sf2 - N / (2.0 * s2)
    d_Z = ((Q - Z * r[:, None]) + 2.0 * (Qv - Z * rv[:, None])) / ell ** 2
    trS = torch.trace(S)
    d_s2 = (-N / (2 * s2) + 0.5 * (Binv * S).sum() / s2 ** 2 + 0.5 * yty / s2 ** 2
            - (b @ beta) / s2 ** 3 + (beta @ S @ beta) / (2 * s2 ** 4) + (N * sf2 - trS) / (2 * s2 ** 2))
    scale = 1.0 / N if normalize == "n" else 1.0
    return F * scale, dict(ell=d_ell * scale, sf2=d_sf2 * scale, s2=d_s2 * scale, Z=d_Z * scale), jit


def sgpr_predict(Xs, X, y, Z, ell, sf2, s2, jitter_policy="gpytorch", full_cov=True, kind="rbf",
                 diag_correction=True, add_noise=True, train_diag_correction=True):
    """Eval-mode predictive of models/sgpr.py:150-160, `likelihood(self(test_x))` after `self.eval()`.

    UPSTREAM (gpytorch 1.3-1.8 ExactGP.__call__, posterior mode): the prediction strategy is built from
    `super().__call__(*train_inputs)` evaluated IN EVAL MODE, and the joint prior from the kernel on cat([train, test]); in both
    calls x1 == x2 and the module is not training, so InducingPointKernel._get_covariance adds the sgpr diagonal correction
    clamp(k_ii - q_ii, 0) to the whole diagonal -- to the TRAINING rows as well as the test rows.  The training covariance the
    strategy inverts is therefore Q_nn + diag(Lambda), Lambda_n = s + max(k_nn - q_nn, 0) (a FITC-like heteroscedastic noise), and
        mean* = Q*n (Q_nn + Lambda)^{-1} y ,  cov* = Q** + diag(corr*) - Q*n (Q_nn + Lambda)^{-1} Qn*  (+ s I by the likelihood).
    With A = L^{-1} Kzx, W = diag(1 / Lambda), B = I + A W A^T, t = L_B^{-1} a*:  mean* = t^T L_B^{-1} A W y, cov_f = t^T t + diag(corr*).
    train_diag_correction=False keeps plain s on the training rows (SURVEY A.6 as first written; B = I + A A^T / s)."""
    N, M = X.shape[0], Z.shape[0]
    Kzz = ard_kernel(Z, Z, ell, sf2, kind)
    L, _ = psd_safe_cholesky(Kzz, jitter_policy)
    A = _tri_solve(L, ard_kernel(Z, X, ell, sf2, kind))
    lam = s2 + ((sf2 - (A * A).sum(0)).clamp_min(0.0) if (train_diag_correction and diag_correction) else 0.0) * torch.ones(N, dtype=X.dtype)
    Aw = A / lam
    LB = torch.linalg.cholesky(torch.eye(M, dtype=X.dtype) + Aw @ A.T)
    c = _tri_solve(LB, (Aw @ y).unsqueeze(-1)).squeeze(-1)
    a = _tri_solve(L, ard_kernel(Z, Xs, ell, sf2, kind))
    t = _tri_solve(LB, a)
    mean = t.T @ c
    corr = (sf2 - (a * a).sum(0)).clamp_min(0.0) if diag_correction else torch.zeros(Xs.shape[0], dtype=X.dtype)
    noise = s2 if add_noise else 0.0
    if full_cov:
        cov = t.T @ t + torch.diag(corr + noise)
        return mean, cov
    var = (t * t).sum(0) + corr + noise
    return mean, var


def sgpr_predict_dense(Xs, X, y, Z, ell, sf2, s2, jitter=0.0, train_diag_correction=True):
    """Dense definition of the same predictive: GP with kernel Q(.,.) = K.z Kzz^{-1} Kz. + the eval-mode diagonal correction on every
    diagonal entry of the joint (train and test) covariance, Gaussian noise s."""
    M = Z.shape[0]
    Kzz = ard_kernel(Z, Z, ell, sf2) + jitter * torch.eye(M, dtype=X.dtype)
    Kxz = ard_kernel(X, Z, ell, sf2)
    Ksz = ard_kernel(Xs, Z, ell, sf2)
    Kinv = torch.linalg.inv(Kzz)
    Qnn = Kxz @ Kinv @ Kxz.T
    corr_n = (sf2 - torch.diagonal(Qnn)).clamp_min(0.0) if train_diag_correction else torch.zeros(X.shape[0], dtype=X.dtype)
    Qnn = Qnn + torch.diag(corr_n + s2)
    Qsn = Ksz @ Kinv @ Kxz.T
    Qss = Ksz @ Kinv @ Ksz.T
    sol = torch.linalg.solve(Qnn, Qsn.T)
    mean = sol.T @ y
    cov = Qss - Qsn @ sol
    corr = (sf2 - torch.diagonal(Qss)).clamp_min(0.0)
    cov = cov + torch.diag(corr + s2)
    return mean, cov
