// Device-side tree bookkeeping of the lock-step multinomial NUTS sampler (host driver: hmc.py nuts_sample).
//
// The reference samples its hyperparameters with  pm.sample(n, tune, chains=1, step=pm.NUTS())
// (models/bayesian_sgpr_hmc.py:73-78, models/all_in_HMC.py:60).  pymc3 is not part of the reference tree; the algorithm restated here is
// its published one (multinomial NUTS: uniform progressive sampling inside a subtree, biased progressive sampling between the old tree
// and the new subtree, U-turn test p_sum . v_edge <= 0 at both edges of every balanced subtree, divergence at |dE| > 1000, acceptance
// statistic = mean over leaves of min(1, exp(-dE))).
//
// At the reference's own sizes (co2: N = 545, M = 100) one logp/dlogp evaluation is ~0.14 ms of graph-replayed GPU work, while the ~60
// elementwise torch calls of the per-leaf bookkeeping cost ~1.3 ms of host launches.  These kernels do that bookkeeping in ONE launch per
// leaf, captured in the same CUDA graph as the evaluation: a leapfrog is one graph replay and the host only reads one flag per doubling.
//
// One warp per chain (grid = C, block = 32): chains are independent, every reduction is a warp shuffle, no shared memory.  The leaf
// index is a per-chain device counter, so the same launch (same arguments) serves every leaf of a subtree and can be replayed.
// Elementwise updates use explicit __dmul_rn / __dadd_rn in the order of the host implementation (no FMA contraction): positions and
// momenta are bit-identical to it, energies agree to the rounding of the P-term sums.
#pragma once
#include <cstdint>
#include "../../include/ggp_b200.h"

namespace ggp {

__device__ __forceinline__ double nuts_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double nuts_neg_inf() { return __longlong_as_double(0xfff0000000000000ll); }
// log(exp(a) + exp(b)) with the -inf conventions of torch.logaddexp
__device__ __forceinline__ double nuts_logaddexp(double a, double b) {
  const double m = fmax(a, b);
  if (m == nuts_neg_inf()) return m;
  return m + log1p(exp(-fabs(a - b)));
}
// (sum p_sum p_left inv_mass <= 0) | (sum p_sum p_right inv_mass <= 0), all lanes return the same value
__device__ __forceinline__ bool nuts_turning(const double* pl, const double* pr, const double* ps, const double* im, int P, int lane) {
  double a = 0.0, b = 0.0;
  for (int i = lane; i < P; i += 32) {
    // host: (p_sum * p_edge * inv_mass).sum(1), products evaluated left to right
    a += __dmul_rn(__dmul_rn(ps[i], pl[i]), im[i]);
    b += __dmul_rn(__dmul_rn(ps[i], pr[i]), im[i]);
  }
  a = nuts_warp_sum(a);
  b = nuts_warp_sum(b);
  return a <= 0.0 || b <= 0.0;
}

// start of a transition: p0 = z / sqrt(inv_mass), e0 = -lp + 0.5 sum p0^2 inv_mass, tree = the single point (x, p0, g)
__global__ void __launch_bounds__(32) k_nuts_begin(ggp_nuts_state s, const double* __restrict__ z) {
  const int c = blockIdx.x, lane = threadIdx.x, P = s.P;
  const int64_t o = (int64_t)c * P;
  double ke = 0.0;
  for (int i = lane; i < P; i += 32) {
    const double im = s.inv_mass[o + i], p0 = z[o + i] / sqrt(im);
    const double xv = s.x[o + i], gv = s.g[o + i];
    s.xl[o + i] = xv; s.xr[o + i] = xv; s.x_prop[o + i] = xv;
    s.pl[o + i] = p0; s.pr[o + i] = p0; s.p_sum[o + i] = p0;
    s.gl[o + i] = gv; s.gr[o + i] = gv; s.g_prop[o + i] = gv;
    ke += __dmul_rn(__dmul_rn(p0, p0), im);
  }
  ke = nuts_warp_sum(ke);
  if (lane == 0) {
    s.e0[c] = -s.lp[c] + 0.5 * ke;
    s.lp_prop[c] = s.lp[c];
    s.log_w[c] = 0.0;
    s.sum_acc[c] = 0.0;
    s.n_leaf[c] = 0.0;
    s.depth[c] = 0;
    s.diverged[c] = 0;
    s.active[c] = 1;
  }
}

// first half of a leapfrog from the subtree's moving edge: p_half = pe + 0.5 e ge, x_eval = xe + e p_half inv_mass
__device__ __forceinline__ void nuts_half_step(const ggp_nuts_state& s, int c, int lane) {
  const int P = s.P;
  const int64_t o = (int64_t)c * P;
  const double e = s.e[c];
  for (int i = lane; i < P; i += 32) {
    const double ph = __dadd_rn(s.pe[o + i], __dmul_rn(__dmul_rn(0.5, e), s.ge[o + i]));
    s.p_half[o + i] = ph;
    s.x_eval[o + i] = __dadd_rn(s.xe[o + i], __dmul_rn(__dmul_rn(e, ph), s.inv_mass[o + i]));
  }
}

// start of a doubling: direction from u[0][c], moving edge = that end of the tree, empty subtree, first half step
__global__ void __launch_bounds__(32) k_nuts_subtree_begin(ggp_nuts_state s) {
  const int c = blockIdx.x, lane = threadIdx.x, P = s.P;
  const int64_t o = (int64_t)c * P;
  const bool right = s.u[c] < 0.5;
  for (int i = lane; i < P; i += 32) {
    const double xe = right ? s.xr[o + i] : s.xl[o + i], pe = right ? s.pr[o + i] : s.pl[o + i], ge = right ? s.gr[o + i] : s.gl[o + i];
    s.xe[o + i] = xe; s.pe[o + i] = pe; s.ge[o + i] = ge;
    s.s_x[o + i] = xe; s.s_g[o + i] = ge;
    s.s_p_sum[o + i] = 0.0;
  }
  if (lane == 0) {
    s.right[c] = right ? 1 : 0;
    s.e[c] = right ? s.eps[c] : -s.eps[c];
    s.s_log_w[c] = nuts_neg_inf();
    s.s_lp[c] = s.lp[c];
    s.s_turn[c] = 0;
    s.s_div[c] = 0;
    s.building[c] = s.active[c];
    s.leaf[c] = 0;
    if (c == 0) *s.any_active = 0;
  }
  __syncwarp();
  nuts_half_step(s, c, lane);
}

// one leaf: second half of the leapfrog with the fresh gradient, multinomial update of the subtree, U-turn checks against the
// checkpoints of every balanced sub-subtree ending here, then the first half of the next leapfrog
__global__ void __launch_bounds__(32) k_nuts_leaf(ggp_nuts_state s) {
  const int c = blockIdx.x, lane = threadIdx.x, P = s.P, C = s.C;
  const int64_t o = (int64_t)c * P;
  const int n = s.leaf[c];
  const double e = s.e[c];
  const double lpn = s.lp_eval[c];
  const double* im = s.inv_mass + o;
  bool building = s.building[c] != 0;
  // pn = p_half + 0.5 e gn (kept in p_half), kinetic energy
  double ke = 0.0;
  for (int i = lane; i < P; i += 32) {
    const double pn = __dadd_rn(s.p_half[o + i], __dmul_rn(__dmul_rn(0.5, e), s.g_eval[o + i]));
    s.p_half[o + i] = pn;
    ke += __dmul_rn(__dmul_rn(pn, pn), im[i]);
  }
  ke = nuts_warp_sum(ke);
  const double en = -lpn + 0.5 * ke;
  double dlt = s.e0[c] - en;
  const bool bad = !isfinite(dlt) || fabs(dlt) > s.max_energy_error;
  const bool div_now = building && bad;
  if (bad) dlt = nuts_neg_inf();
  const bool good = building && !bad;
  const double s_log_w = s.s_log_w[c];
  const double new_w = nuts_logaddexp(s_log_w, dlt);
  const double ul = s.u[(int64_t)(2 + n) * C + c];
  const bool take = good && (log(ul) < dlt - new_w);
  __syncwarp();
  for (int i = lane; i < P; i += 32) {
    const double xn = s.x_eval[o + i], pn = s.p_half[o + i], gn = s.g_eval[o + i];
    if (take) { s.s_x[o + i] = xn; s.s_g[o + i] = gn; }
    if (good) {
      s.s_p_sum[o + i] = __dadd_rn(s.s_p_sum[o + i], pn);
      s.xe[o + i] = xn; s.pe[o + i] = pn; s.ge[o + i] = gn;
    }
  }
  __syncwarp();
  // checkpoint slots (hmc.py _ckpt_range): save at even leaves, check at odd ones
  const int idx_max = __popc(n >> 1);
  int t = 0;
  for (int m = n; m & 1; m >>= 1) ++t;
  const int idx_min = idx_max - t + 1;
  bool turned = false;
  if ((n & 1) == 0) {
    double* pck = s.p_ck + ((int64_t)idx_max * C + c) * P;
    double* sck = s.ps_ck + ((int64_t)idx_max * C + c) * P;
    for (int i = lane; i < P; i += 32) { pck[i] = s.pe[o + i]; sck[i] = s.s_p_sum[o + i]; }
  } else {
    for (int k = idx_max; k >= idx_min; --k) {
      const double* pck = s.p_ck + ((int64_t)k * C + c) * P;
      const double* sck = s.ps_ck + ((int64_t)k * C + c) * P;
      double a = 0.0, b = 0.0;
      for (int i = lane; i < P; i += 32) {
        const double ps = __dadd_rn(__dadd_rn(s.s_p_sum[o + i], -sck[i]), pck[i]);   // momentum sum of the sub-subtree
        a += __dmul_rn(__dmul_rn(ps, pck[i]), im[i]);
        b += __dmul_rn(__dmul_rn(ps, s.pe[o + i]), im[i]);
      }
      a = nuts_warp_sum(a);
      b = nuts_warp_sum(b);
      turned = turned || a <= 0.0 || b <= 0.0;
    }
  }
  if (lane == 0) {
    if (div_now) s.s_div[c] = 1;
    if (building) {
      s.sum_acc[c] += exp(fmin(dlt, 0.0));
      s.n_leaf[c] += 1.0;
    }
    if (take) s.s_lp[c] = lpn;
    if (good) s.s_log_w[c] = new_w;
    bool b2 = good;
    if (b2 && turned) { s.s_turn[c] = 1; b2 = false; }
    s.building[c] = b2 ? 1 : 0;
    s.leaf[c] = n + 1;
  }
  __syncwarp();
  nuts_half_step(s, c, lane);
}

// end of a doubling: accept the subtree unless it turned or diverged, biased progressive sampling, extend the tree, global U-turn test
__global__ void __launch_bounds__(32) k_nuts_subtree_end(ggp_nuts_state s) {
  const int c = blockIdx.x, lane = threadIdx.x, P = s.P, C = s.C;
  const int64_t o = (int64_t)c * P;
  const bool active = s.active[c] != 0, s_turn = s.s_turn[c] != 0, s_div = s.s_div[c] != 0, right = s.right[c] != 0;
  const bool ok = active && !s_turn && !s_div;
  const double s_log_w = s.s_log_w[c], log_w = s.log_w[c];
  const bool take = ok && (log(s.u[C + c]) < s_log_w - log_w);
  for (int i = lane; i < P; i += 32) {
    if (take) { s.x_prop[o + i] = s.s_x[o + i]; s.g_prop[o + i] = s.s_g[o + i]; }
    if (ok) {
      s.p_sum[o + i] = __dadd_rn(s.p_sum[o + i], s.s_p_sum[o + i]);
      if (right) { s.xr[o + i] = s.xe[o + i]; s.pr[o + i] = s.pe[o + i]; s.gr[o + i] = s.ge[o + i]; }
      else       { s.xl[o + i] = s.xe[o + i]; s.pl[o + i] = s.pe[o + i]; s.gl[o + i] = s.ge[o + i]; }
    }
  }
  __syncwarp();
  const bool turn = nuts_turning(s.pl + o, s.pr + o, s.p_sum + o, s.inv_mass + o, P, lane);
  if (lane == 0) {
    if (active && s_div) s.diverged[c] = 1;
    if (take) s.lp_prop[c] = s.s_lp[c];
    if (ok) {
      s.log_w[c] = nuts_logaddexp(log_w, s_log_w);
      s.depth[c] += 1;
    }
    const int act = (ok && !turn) ? 1 : 0;
    s.active[c] = act;
    if (act) atomicOr(s.any_active, 1);
  }
}

// end of a transition: the proposal becomes the state; acceptance statistic; optional trace row `k`
__global__ void __launch_bounds__(32) k_nuts_end(ggp_nuts_state s, int k) {
  const int c = blockIdx.x, lane = threadIdx.x, P = s.P, C = s.C;
  const int64_t o = (int64_t)c * P;
  for (int i = lane; i < P; i += 32) {
    const double xv = s.x_prop[o + i];
    s.x[o + i] = xv;
    s.g[o + i] = s.g_prop[o + i];
    if (k >= 0) s.samples[((int64_t)k * C + c) * P + i] = xv;
  }
  if (lane == 0) {
    const double lp = s.lp_prop[c];
    s.lp[c] = lp;
    s.acc_prob[c] = s.sum_acc[c] / fmax(s.n_leaf[c], 1.0);
    if (k >= 0) {
      s.lps[(int64_t)k * C + c] = lp;
      s.depths[(int64_t)k * C + c] = s.depth[c];
      s.nleaps[(int64_t)k * C + c] = (int)s.n_leaf[c];
      s.divs[(int64_t)k * C + c] = s.diverged[c];
    }
  }
}

// ---- pymc3 target of models/bayesian_sgpr_hmc.py:60-71 around the collapsed bound: transforms, priors, Jacobians -------------------
// x = (ls_log__[D], sig_f_log__, sig_n_log__) per chain.  One thread per chain (C is a handful, D <= a few dozen): what was ~55
// elementwise torch launches per evaluation is two.
__global__ void k_vfe_theta(const double* __restrict__ x, int C, int D, double* __restrict__ theta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* xr = x + (int64_t)c * (D + 2);
  double* t = theta + (int64_t)c * (D + 2);
  for (int d = 0; d < D; ++d) t[d] = exp(xr[d]);            // ell = e^x
  const double sf = exp(xr[D]), sn = exp(xr[D + 1]);
  t[D] = sf * sf;                                            // outputscale = sig_f^2 (update_model_to_hyper, :82-86)
  t[D + 1] = sn * sn;                                        // noise variance = sig_n^2
}
// logp = F + [Gamma(2,1) on ell_d, HalfCauchy(1) on sig_f, sig_n, log-Jacobians of the log transforms]; dlogp by the
// chain rule from dF/d(ell, sf2, s2).  A failed factorisation (info / info_b) or a non-finite value gives logp = -inf, dlogp = 0.
__global__ void k_vfe_logp(const double* __restrict__ x, const double* __restrict__ bound, const double* __restrict__ grad, int64_t ldg,
                           const int* __restrict__ info, const int* __restrict__ info_b, int C, int D, int with_prior,
                           double* __restrict__ lp, double* __restrict__ dx) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double* xr = x + (int64_t)c * (D + 2);
  const double* g = grad + (int64_t)c * ldg;
  double* o = dx + (int64_t)c * (D + 2);
  const double LOG2 = 0.6931471805599453, LOGPI = 1.1447298858494002;
  double l = bound[c], sx = 0.0, sp = 0.0;
  for (int d = 0; d < D; ++d) {
    const double ell = exp(xr[d]);
    double v = g[d] * ell;
    if (with_prior) { sp += log(ell) - ell; sx += xr[d]; v += (1.0 - ell) + 1.0; }
    o[d] = v;
  }
  const double sf = exp(xr[D]), sn = exp(xr[D + 1]), sf2 = sf * sf, sn2 = sn * sn;
  double vf = g[D] * 2.0 * sf2, vn = g[D + 1] * 2.0 * sn2;
  if (with_prior) {
    sx += xr[D] + xr[D + 1];
    l += sp + (LOG2 - LOGPI - log1p(sf2)) + (LOG2 - LOGPI - log1p(sn2)) + sx;
    vf += -2.0 * sf2 / (1.0 + sf2) + 1.0;
    vn += -2.0 * sn2 / (1.0 + sn2) + 1.0;
  }
  o[D] = vf;
  o[D + 1] = vn;
  const bool bad = (info && info[c] != 0) || (info_b && info_b[c] != 0) || !isfinite(l);
  if (bad) {
    l = nuts_neg_inf();
    for (int d = 0; d < D + 2; ++d) o[d] = 0.0;
  }
  lp[c] = l;
}

}  // namespace ggp
