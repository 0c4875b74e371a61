"""gpytorch-free mirrors of the reference's model classes, exposing the attribute surface its host loops touch
 with the arithmetic routed through the CUDA hot path.

  SparseGPR                         models/sgpr.py:22-160
  BayesianSparseGPR_HMC             models/bayesian_sgpr_hmc.py:26-231   (+ mixture_posterior_predictive)
  StochasticVariationalGP           models/svgp.py:24-141
  BayesianStochasticVariationalGP   models/bayesian_svgp.py:30-207
Parameterisation follows gpytorch: raw parameters, value = softplus(raw) (+1e-4 for the noise), all raws start at 0.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as Fnn

from . import functions as F
from .engine import Engine


def inv_softplus(v):
    v = torch.as_tensor(v, dtype=torch.float64)
    return v + torch.log(-torch.expm1(-v))


class RBFKernel(nn.Module):
    def __init__(self, ard_num_dims=1, dtype=torch.float64):
        super().__init__()
        self.ard_num_dims = ard_num_dims
        self.raw_lengthscale = nn.Parameter(torch.zeros(1, ard_num_dims, dtype=dtype))

    @property
    def lengthscale(self):
        return Fnn.softplus(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, v):
        v = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v, dtype=self.raw_lengthscale.dtype)
        self.raw_lengthscale.data.copy_(inv_softplus(v).reshape(1, -1).to(self.raw_lengthscale.device))


class ScaleKernel(nn.Module):
    def __init__(self, base_kernel, dtype=torch.float64):
        super().__init__()
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros((), dtype=dtype))

    @property
    def outputscale(self):
        return Fnn.softplus(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, v):
        self.raw_outputscale.data.copy_(inv_softplus(float(v)).to(self.raw_outputscale.device))


class _HomoskedasticNoise(nn.Module):
    def __init__(self, lower=1e-4, dtype=torch.float64):
        super().__init__()
        self.lower = lower
        self.raw_noise = nn.Parameter(torch.zeros(1, dtype=dtype))

    @property
    def noise(self):
        return Fnn.softplus(self.raw_noise) + self.lower

    @noise.setter
    def noise(self, v):
        self.raw_noise.data.copy_(inv_softplus(float(v) - self.lower).reshape(1).to(self.raw_noise.device))


class GaussianLikelihood(nn.Module):
    """gpytorch.likelihoods.GaussianLikelihood look-alike: noise = softplus(raw) + 1e-4 (GreaterThan(1e-4))."""

    def __init__(self, noise_lower_bound=1e-4):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise(noise_lower_bound)

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, v):
        self.noise_covar.noise = v


class BernoulliLikelihood(nn.Module):
    """Probit Bernoulli (no .noise attribute -> models/svgp.py:40-44 takes the multitask-strategy branch)."""


class InducingPointKernel(nn.Module):
    def __init__(self, base_kernel, inducing_points, likelihood=None):
        super().__init__()
        self.base_kernel = base_kernel
        Z = torch.as_tensor(inducing_points).detach().clone().to(torch.float64)
        if Z.dim() == 1:
            Z = Z.unsqueeze(-1)
        self.inducing_points = nn.Parameter(Z)


class PredictiveNormal:
    """What likelihood(model(test_x)) returns, as far as utils/metrics.py and utils/posterior_predictive.py use it."""

    def __init__(self, loc, covariance_matrix=None, variance=None):
        self.loc = self.mean = loc
        self._cov, self._var = covariance_matrix, variance

    @property
    def covariance_matrix(self):
        return self._cov if self._cov is not None else torch.diag(self._var)

    @property
    def variance(self):
        return self._var if self._var is not None else torch.diagonal(self._cov)

    @property
    def stddev(self):
        return self.variance.sqrt()

    def log_prob(self, y):
        cov = self.covariance_matrix
        n = cov.shape[-1]
        eye = torch.eye(n, dtype=cov.dtype, device=cov.device)
        for j in (0.0, 1e-8, 1e-7, 1e-6):
            L, info = torch.linalg.cholesky_ex(cov + j * eye)
            if int(info) == 0:
                break
        else:
            from .engine import NotPSDError
            raise NotPSDError("predictive covariance not positive definite after the jitter ladder (0, 1e-8, 1e-7, 1e-6)")
        r = torch.linalg.solve_triangular(L, (y.to(cov.dtype) - self.loc).unsqueeze(-1), upper=False).squeeze(-1)
        return -0.5 * (r @ r) - torch.log(torch.diagonal(L)).sum() - 0.5 * n * math.log(2.0 * math.pi)


class SparseGPR(nn.Module):
    """Titsias collapsed SGPR (models/sgpr.py:22).  Same constructor, train_model and posterior_predictive."""

    jitter_policy = "gpytorch"

    def __init__(self, train_x, train_y, likelihood, Z_init):
        super().__init__()
        self.train_x = train_x if train_x.dim() > 1 else train_x.unsqueeze(-1)
        self.train_y = train_y
        self.inducing_points = Z_init
        self.num_inducing = len(Z_init)
        self.likelihood = likelihood
        self.base_covar_module = ScaleKernel(RBFKernel(ard_num_dims=self.train_x.shape[-1]))
        self.covar_module = InducingPointKernel(self.base_covar_module, inducing_points=Z_init, likelihood=likelihood)

    def named_hyperparameters(self):
        return self.named_parameters()

    def _theta(self):
        return torch.cat([self.base_covar_module.base_kernel.lengthscale.reshape(-1), self.base_covar_module.outputscale.reshape(-1),
                          self.likelihood.noise.reshape(-1)])

    def bound(self):
        """mll(self.forward(train_x), train_y): the collapsed bound / N (models/sgpr.py:123-125)."""
        return F.sgpr_bound(self.train_x, self.train_y, self.covar_module.inducing_points,
                            self.base_covar_module.base_kernel.lengthscale, self.base_covar_module.outputscale,
                            self.likelihood.noise, dict(jitter_policy=self.jitter_policy, normalize="n"))

    def train_model(self, optimizer, combine_terms=True, n_restarts=10, max_steps=10000, verbose=False):
        self.train()
        losses = []
        for j in range(max_steps):
            optimizer.zero_grad()
            loss = -self.bound()
            losses.append(loss.item())
            loss.backward()
            if verbose and j % 1000 == 0:
                print('Iter %d/%d - Loss: %.3f   outputscale: %.3f  lengthscale: %s   noise: %.3f ' % (
                    j + 1, max_steps, loss.item(), self.base_covar_module.outputscale.item(),
                    self.base_covar_module.base_kernel.lengthscale, self.likelihood.noise.item()))
            optimizer.step()
        return losses

    def posterior_predictive(self, test_x, full_cov=True):
        """likelihood(self(test_x)) in eval mode (models/sgpr.py:150-160)."""
        self.eval()
        with torch.no_grad():
            test_x = test_x if test_x.dim() > 1 else test_x.unsqueeze(-1)
            eng = Engine.get(self.train_x.device)
            theta = self._theta()
            Z = self.covar_module.inducing_points
            eng.sgpr_predict_state(self.train_x, self.train_y, Z, theta, jitter_policy=self.jitter_policy)
            mean, var, cov = eng.sgpr_predict(test_x, Z, theta, full_cov=full_cov)
        return PredictiveNormal(mean[0], cov[0] if cov is not None else None, var[0])


class BayesianSparseGPR_HMC(SparseGPR):
    """Doubly collapsed SGPR: Adam on Z alternating with HMC over theta on the same bound (models/bayesian_sgpr_hmc.py:26)."""

    def __init__(self, train_x, train_y, likelihood, Z_init):
        super().__init__(train_x, train_y, likelihood, Z_init)
        self.data_dim = self.train_x.shape[1]

    def freeze_kernel_hyperparameters(self):
        for name, parameter in self.named_hyperparameters():
            if name != 'covar_module.inducing_points':
                parameter.requires_grad = False

    def sample_optimal_variational_hyper_dist(self, n_samples, input_dim, Z_opt, tune, sampler_params=None, chains=1):
        """pm.sample(n_samples, tune=tune, chains=1, step=pm.NUTS()) on the VFE model (models/bayesian_sgpr_hmc.py:58-80): NUTS with
        pymc3's defaults.  sampler_params: dict(sampler='nuts'|'hmc', max_treedepth, target_accept, step_size, n_leapfrog) -- the
        fixed-length sampler is opt-in."""
        from .hmc import sample_hyper
        Z = torch.as_tensor(Z_opt, dtype=torch.float64, device=self.train_x.device).reshape(-1, input_dim)
        sp = dict(sampler_params or {})
        traces, res = sample_hyper(self.train_x, self.train_y, Z, n_samples, tune, chains=chains, sampler=sp.get('sampler', 'nuts'),
                                   max_treedepth=sp.get('max_treedepth', 10), target_accept=sp.get('target_accept', 0.8),
                                   step_size=sp.get('step_size', sp.get('step_scale')), n_leapfrog=sp.get('n_leapfrog', 10))
        self.last_sampler_result = res
        return traces[0] if chains == 1 else traces

    def update_model_to_hyper(self, elbo, hyper_sample):
        self.likelihood.noise_covar.noise = hyper_sample['sig_n'] ** 2
        self.base_covar_module.outputscale = hyper_sample['sig_f'] ** 2
        self.base_covar_module.base_kernel.lengthscale = hyper_sample['ls']

    def stochastic_bound(self, trace_hyper):
        """(1/|trace|) sum_i bound(theta_i)/N with only Z differentiable (models/bayesian_sgpr_hmc.py:121-131): ONE batched
        evaluation over all draws instead of |trace| sequential ones."""
        thetas = trace_hyper.thetas().to(self.train_x.device)
        return _BatchedBoundMean.apply(self.train_x, self.train_y, self.covar_module.inducing_points, thetas, self.jitter_policy)

    def train_model(self, optimizer, max_steps=10000, hmc_scheduler=(200, 500, 1000, 1500), verbose=False):
        self.train()
        hmc_scheduler = list(hmc_scheduler)
        losses, trace_hyper, trace_step_size, trace_perf_time = [], None, [], []
        for n_iter in range(max_steps):
            optimizer.zero_grad()
            if n_iter < hmc_scheduler[0]:
                loss = -self.bound()
                losses.append(loss.item())
                loss.backward()
                optimizer.step()
            else:
                self.freeze_kernel_hyperparameters()
                if trace_hyper is not None:
                    loss = -self.stochastic_bound(trace_hyper)
                    losses.append(loss.item())
                    loss.backward()
                    optimizer.step()
                if n_iter in hmc_scheduler:
                    Z_opt = self.covar_module.inducing_points.detach().cpu().numpy()
                    if n_iter in (hmc_scheduler[0], hmc_scheduler[-1]):
                        num_tune, num_samples = 100, 20
                    else:
                        num_tune, num_samples = 25, 10
                    trace_hyper = self.sample_optimal_variational_hyper_dist(num_samples, self.data_dim, Z_opt, num_tune)
                    trace_step_size.append(trace_hyper.get_sampler_stats('step_size')[0])
                    trace_perf_time.append(trace_hyper.get_sampler_stats('perf_counter_diff').sum())
        return losses, trace_hyper, trace_step_size, trace_perf_time

    def train_fixed_model(self, num_tune=500, num_samples=500):
        self.train()
        Z_opt = self.covar_module.inducing_points.detach().cpu().numpy()
        trace_hyper = self.sample_optimal_variational_hyper_dist(num_samples, self.data_dim, Z_opt, num_tune)
        return (trace_hyper, [trace_hyper.get_sampler_stats('step_size')[0]],
                [trace_hyper.get_sampler_stats('perf_counter_diff').sum()])


class _BatchedBoundMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, y, Z, thetas, jitter_policy):
        eng = Engine.get(X.device, precision=F.pick_precision(X.shape[0], Z.shape[0], X.shape[1]))
        out = eng.sgpr_eval(X, y, Z, thetas, jitter_policy=jitter_policy, need_grad=True)
        n, B = out["n_total"][0], thetas.shape[0]
        D, M = X.shape[1], Z.shape[0]
        ctx.save_for_backward(out["grad"][:, D + 2:].mean(0).view(M, D) / n)
        return out["bound"].mean() / n

    @staticmethod
    def backward(ctx, gout):
        (gZ,) = ctx.saved_tensors
        return None, None, gout * gZ, None, None


def mixture_posterior_predictive(model, test_x, trace_hyper, full_cov=True):
    """List of predictive distributions, one per HMC draw (models/bayesian_sgpr_hmc.py:198-231) -- all draws are evaluated in one
    batched pass over the training rows instead of |trace| separate ones; draws whose predictive covariance is not PSD
    (+1e-4 I) are dropped like the reference does."""
    with torch.no_grad():
        test_x = test_x if test_x.dim() > 1 else test_x.unsqueeze(-1)
        eng = Engine.get(model.train_x.device)
        thetas = trace_hyper.thetas().to(model.train_x.device)
        Z = model.covar_module.inducing_points
        eng.sgpr_predict_state(model.train_x, model.train_y, Z, thetas, jitter_policy=model.jitter_policy)
        mean, var, cov = eng.sgpr_predict(test_x, Z, thetas, full_cov=full_cov)
    out = []
    eye = torch.eye(test_x.shape[0], dtype=torch.float64, device=mean.device) * 1e-4 if full_cov else None
    for i in range(thetas.shape[0]):
        if full_cov:
            _, info = torch.linalg.cholesky_ex(cov[i] + eye)
            if int(info) != 0:
                print('Not psd for sample ' + str(i))
                continue
        out.append(PredictiveNormal(mean[i], cov[i] if full_cov else None, var[i]))
    return out


class StochasticVariationalGP(nn.Module):
    """Hensman SVGP with a whitened Cholesky variational distribution (models/svgp.py:24)."""

    def __init__(self, train_x, train_y, likelihood, Z_init, num_tasks=None):
        super().__init__()
        self.train_x = train_x if train_x.dim() > 1 else train_x.unsqueeze(-1)
        self.train_y = train_y
        Z = torch.as_tensor(Z_init).detach().clone().to(torch.float64)
        self.inducing_inputs = nn.Parameter(Z if Z.dim() > 1 else Z.unsqueeze(-1))
        self.num_inducing = len(Z_init)
        self.likelihood = likelihood
        self.is_gaussian = hasattr(likelihood, "noise")
        M = self.num_inducing
        self.variational_mean = nn.Parameter(torch.zeros(M, dtype=torch.float64))
        self.chol_variational_covar = nn.Parameter(torch.eye(M, dtype=torch.float64))
        self.covar_module = ScaleKernel(RBFKernel(ard_num_dims=self.train_x.shape[-1]))
        self._variational_initialised = False

    def _maybe_init_variational(self):
        # gpytorch re-initialises q(u) on the FIRST call of the strategy: m <- 0 + 1e-3 * randn(M) on the global torch RNG
        if not self._variational_initialised:
            with torch.no_grad():
                self.variational_mean.copy_((1e-3 * torch.randn(self.num_inducing)).to(self.variational_mean))
                self.chol_variational_covar.copy_(torch.eye(self.num_inducing, dtype=torch.float64))
            self._variational_initialised = True

    def _noise(self):
        return self.likelihood.noise if self.is_gaussian else torch.ones(1, dtype=torch.float64, device=self.train_x.device)

    def elbo(self, x_batch, y_batch):
        self._maybe_init_variational()
        cfg = dict(jitter_policy="gpytorch", likelihood="gaussian" if self.is_gaussian else "bernoulli")
        return F.svgp_elbo(x_batch, y_batch, self.inducing_inputs, self.variational_mean, self.chol_variational_covar,
                           self.covar_module.base_kernel.lengthscale, self.covar_module.outputscale, self._noise(),
                           len(self.train_y), cfg)

    def train_model(self, optimizer, train_loader, minibatch_size=100, num_epochs=25, combine_terms=True):
        losses = []
        for _ in range(num_epochs):
            for x_batch, y_batch in train_loader:
                self.train()
                optimizer.zero_grad()
                loss = -self.elbo(x_batch.to(self.train_x.device), y_batch.to(self.train_x.device))
                losses.append(loss.item())
                loss.backward()
                optimizer.step()
        return losses

    def posterior_predictive(self, test_x):
        self.eval()
        self._maybe_init_variational()
        with torch.no_grad():
            test_x = test_x if test_x.dim() > 1 else test_x.unsqueeze(-1)
            eng = Engine.get(self.train_x.device)
            theta = torch.cat([self.covar_module.base_kernel.lengthscale.reshape(-1), self.covar_module.outputscale.reshape(-1),
                               self._noise().reshape(-1)])
            mean, var = eng.svgp_predict(test_x, self.inducing_inputs, self.variational_mean, self.chol_variational_covar, theta,
                                         add_noise=self.is_gaussian)
        return PredictiveNormal(mean[0], None, var[0])


class VariationalHyperDist(nn.Module):
    """q(log theta) = N(q_mu, L L^T + 1e-5 I) (models/bayesian_svgp.py:30-71); every lower-triangular entry of L (diagonal
    included) is overwritten from q_sigma_vec in tril_indices order."""

    def __init__(self, hyper_dim, n, prior_var=0.01):
        super().__init__()
        self.hyper_dim, self.n, self.prior_var = hyper_dim, n, prior_var
        self.q_mu = nn.Parameter(torch.randn(hyper_dim, dtype=torch.float64) * 1e-3)
        self.q_sigma_vec = nn.Parameter(torch.randn(hyper_dim * (hyper_dim + 1) // 2, dtype=torch.float64) * 1e-3)

    def construct_sigma(self):
        r, c = torch.tril_indices(self.hyper_dim, self.hyper_dim)
        L = torch.zeros(self.hyper_dim, self.hyper_dim, dtype=torch.float64, device=self.q_mu.device)
        L = L.index_put((r.to(L.device), c.to(L.device)), self.q_sigma_vec)
        return L @ L.T + 1e-5 * torch.eye(self.hyper_dim, dtype=torch.float64, device=L.device)

    def distribution(self):
        return torch.distributions.MultivariateNormal(self.q_mu, self.construct_sigma())

    def kl_per_point(self):
        p = torch.distributions.MultivariateNormal(torch.zeros_like(self.q_mu), self.prior_var * torch.eye(
            self.hyper_dim, dtype=torch.float64, device=self.q_mu.device))
        return torch.distributions.kl_divergence(self.distribution(), p) / self.n

    def forward(self, num_samples):
        return self.distribution().rsample(torch.Size([num_samples]))


class BayesianStochasticVariationalGP(StochasticVariationalGP):
    """SVGP + Gaussian q(log theta): 5 theta draws per minibatch, each an ELBO evaluation (models/bayesian_svgp.py:87-181).
    theta[0] -> outputscale, theta[1:-1] -> lengthscale, theta[-1]^2 -> noise (:129-133).  The draws are evaluated in one
    batched launch sequence.  As upstream (property setters copy into raw_*.data) the data term does not
    back-propagate into q(log theta): q is trained through its KL term only; pass reparam=True for the alternative."""

    num_hyper_draws = 5

    def __init__(self, train_x, train_y, likelihood, Z_init):
        super().__init__(train_x, train_y, likelihood, Z_init)
        self.n = len(train_y)
        self.input_dim = self.train_x.shape[1]
        self.log_theta = VariationalHyperDist(self.input_dim + 2, self.n)

    def _thetas(self, log_theta_samples):
        th = torch.exp(log_theta_samples)
        return torch.cat([th[:, 1:-1], th[:, :1], th[:, -1:] ** 2], dim=1)   # engine order: ell[D], sf2, s2

    def train_step_loss(self, x_batch, y_batch):
        self._maybe_init_variational()
        draws = self.log_theta(self.num_hyper_draws)
        thetas = self._thetas(draws.detach())
        val = _BatchedElboMean.apply(x_batch, y_batch, self.inducing_inputs, self.variational_mean, self.chol_variational_covar,
                                     thetas, float(self.n))
        return -(val - self.log_theta.kl_per_point())

    def train_model(self, optimizer, train_loader, minibatch_size=100, num_epochs=25, combine_terms=True):
        epoch_losses, batch_losses = [], []
        for _ in range(num_epochs):
            batch_losses = []
            for x_batch, y_batch in train_loader:
                optimizer.zero_grad()
                loss = self.train_step_loss(x_batch.to(self.train_x.device), y_batch.to(self.train_x.device))
                batch_losses.append(loss.item())
                loss.backward()
                optimizer.step()
            epoch_losses.append(float(np.sum(batch_losses)))
        return epoch_losses, batch_losses


    def sample_variational_log_hyper(self, num_samples):
        return self.log_theta(num_samples)

    def mixture_posterior_predictive(self, test_x, num_samples=100):
        """models/bayesian_svgp.py:183-207, followed literally: 100 draws l_i ~ q(log theta), theta_i = softplus(l_i) handed to
        forward() as `log_theta`, i.e. exponentiated again (:124) before update_covar_module_at_theta (:129-133).  All draws go
        through ONE batched predictive launch sequence; returns a list of predictive distributions (marginals), one per draw."""
        self.eval()
        self._maybe_init_variational()
        with torch.no_grad():
            test_x = test_x if test_x.dim() > 1 else test_x.unsqueeze(-1)
            draws = self.sample_variational_log_hyper(num_samples)
            thetas = self._thetas(Fnn.softplus(draws))
            eng = Engine.get(self.train_x.device)
            mean, var = eng.svgp_predict(test_x, self.inducing_inputs, self.variational_mean, self.chol_variational_covar, thetas,
                                         add_noise=self.is_gaussian)
        self.last_mixture_thetas = thetas
        return [PredictiveNormal(mean[i], None, var[i]) for i in range(thetas.shape[0])]


class _BatchedElboMean(torch.autograd.Function):
    """mean over theta draws of the SVGP ELBO; differentiable in Z, q_mean, q_chol."""

    @staticmethod
    def forward(ctx, xb, yb, Z, qm, qL, thetas, num_data):
        eng = Engine.get(xb.device)
        out = eng.svgp_eval(xb, yb, Z, qm, qL, thetas, num_data=num_data, need_grad=True)
        D, M = xb.shape[1], Z.shape[0]
        g = out["grad"].mean(0)
        o = D + 2
        ctx.save_for_backward(g[o:o + M * D].view(M, D), g[o + M * D:o + M * D + M], g[o + M * D + M:].view(M, M))
        return out["value"].mean()

    @staticmethod
    def backward(ctx, gout):
        gZ, gm, gL = ctx.saved_tensors
        return None, None, gout * gZ, gout * gm, gout * gL, None, None


MODEL_DICTIONARY = {"SGPR": SparseGPR, "Bayesian_SGPR_HMC": BayesianSparseGPR_HMC, "SVGP": StochasticVariationalGP,
                    "Bayesian_SVGP": BayesianStochasticVariationalGP}
