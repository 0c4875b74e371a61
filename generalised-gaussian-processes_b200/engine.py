"""Host driver of the CUDA hot path: owns a ggp handle per device and sequences
factor -> pass1 -> [all-reduce] -> finish -> pass2 -> [all-reduce]   (include/ggp_b200.h).

Replaces what gpytorch / pymc3 execute under the reference's calls
  models/sgpr.py:123-129 (forward, -mll, backward), models/bayesian_sgpr_hmc.py:66-78 (VFE logp/dlogp per leapfrog),
  models/sgpr.py:150-160 (eval-mode predictive).
PyTorch is used for device memory, streams and torch.distributed only.
"""
import ctypes

import torch

from . import _lib
from ._lib import GgpCfg, KERNELS, KERNEL_COMPOSITE, LIKELIHOODS, PRECISIONS, check


class NotPSDError(RuntimeError):
    """Kzz (or I + A A^T / s) not positive definite after the jitter ladder (gpytorch NotPSDError analogue)."""


def jitter_ladder(policy, dtype=torch.float64):
    """gpytorch psd_safe_cholesky ladder / pymc3 stabilize / fixed value."""
    if policy == "gpytorch":
        j0 = 1e-8 if dtype == torch.float64 else 1e-6
        return [0.0, j0, j0 * 10.0, j0 * 100.0]
    if policy == "pymc3":
        return [1e-6]
    return [float(policy)]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f64c(t, device):
    return t.detach().to(device=device, dtype=torch.float64).contiguous()


class Engine:
    """One per (device, kernel, precision).  Not thread-safe; one stream at a time."""

    _cache = {}

    @classmethod
    def get(cls, device=None, kernel="rbf", precision="fp64", chunk_rows=0, tile_cache_mib=None, kernel_param=0.0):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        device = torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if not isinstance(kernel, str):
            kernel = tuple(tuple(t) for t in kernel)     # composite program: hashable
        key = (device.index, kernel, precision, chunk_rows, tile_cache_mib, float(kernel_param))
        if key not in cls._cache:
            cls._cache[key] = cls(device, kernel, precision, chunk_rows, tile_cache_mib, kernel_param)
        return cls._cache[key]

    def __init__(self, device, kernel="rbf", precision="fp64", chunk_rows=0, tile_cache_mib=None, kernel_param=0.0):
        if not torch.cuda.is_available():
            raise RuntimeError("the sparse-GP hot path needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if tile_cache_mib is None:
            # keep the k(X,Z) tiles of pass 1 in HBM for pass 2 when they fit in a quarter of the free device memory (<= 32 GiB);
            # tile_cache_mib=0 restores the strictly streaming behaviour (never more than chunk_rows x m of k(X,Z) alive)
            free_b, _ = torch.cuda.mem_get_info(torch.device(device))
            tile_cache_mib = int(min(32 * 1024, free_b // (4 * 1024 * 1024)))
        if kernel == "rq" and not kernel_param > 0:
            raise ValueError("kernel='rq' needs kernel_param = alpha > 0")
        # composite kernel (a tuple of terms, each a tuple of factor names: sum of scaled products, the reference's CO2 model
        # experiments/co2_bayesian_sgpr_hmc.py:74-83): theta rows are [kernel parameters (P, _lib.make_kprog's order), s2] and the
        # gradient rows [d/d(kernel parameters), d_s2, d_Z]; SGPR entry points only, FP64 DMMA plan
        self.prog = None if isinstance(kernel, str) else tuple(tuple(t) for t in kernel)
        self._kprog = _lib.make_kprog(self.prog) if self.prog is not None else None
        self._kkeep = None
        kcode = KERNELS[kernel] if self.prog is None else KERNEL_COMPOSITE
        self.cfg = GgpCfg(kcode, PRECISIONS[precision], int(chunk_rows), int(tile_cache_mib), float(kernel_param))
        # the FP64 DMMA plan on the same handle: an fp64_i8 engine evaluates on it when the jitter ladder had to engage (see sgpr_eval)
        self.cfg_dmma = GgpCfg(kcode, PRECISIONS["fp64"], int(chunk_rows), int(tile_cache_mib), float(kernel_param))
        self.ladder_levels = []
        h = ctypes.c_void_p()
        check(self.lib.ggp_create(ctypes.byref(h), self.device.index), "ggp_create")
        self.h = h
        self._shape = None
        self.launch_count = 0
        self._side = None            # side stream: tile prefetch overlapped with the factorisation
        self._copy = None            # copy stream: host rows uploaded piece by piece ahead of their tile build
        self.prefetch_min_rows = 65536
        self._sgpr_state = None      # (Z ptr/shape, theta values) the handle's factor + B state belongs to, or None when clobbered
        self.i8_batch_min_elems = 1 << 26    # n_local * m above which a batch of theta rows runs draw by draw on the sliced-integer plans

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.ggp_destroy(self.h)
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _ident(t):
        return (t.data_ptr(), t._version, tuple(t.shape), str(t.device))

    def _mark_state(self, Z, theta, by_value=False):
        """Remember which (Z, theta) the handle's factor + B state belongs to.  The hot path (sgpr_eval) records the IDENTITY of the
        caller's tensors (pointer + version counter: no device read-back); sgpr_predict_state records the values."""
        if by_value:
            self._sgpr_state = ("values", Z.detach().cpu().clone(), theta.detach().cpu().reshape(-1).clone())
        else:
            self._sgpr_state = ("ident", self._ident(Z), self._ident(theta))

    def _check_state(self, Z, theta):
        st = self._sgpr_state
        ok = False
        if st is not None and st[0] == "ident":
            ok = st[1] == self._ident(Z) and st[2] == self._ident(theta)
        elif st is not None:
            zc, tc = Z.detach().cpu(), theta.detach().cpu().reshape(-1)
            ok = zc.shape == st[1].shape and tc.shape == st[2].shape and torch.equal(zc.to(st[1].dtype), st[1]) and torch.equal(tc.to(st[2].dtype), st[2])
        if not ok:
            raise RuntimeError("sgpr_predict: the handle does not hold the SGPR state of this (Z, theta) -- another evaluation "
                               "(SVGP / SGPMC, a different theta, chol, a re-reservation) has overwritten it; call "
                               "sgpr_predict_state(X, y, Z, theta) first")

    def reserve(self, n_local, m, d, batch):
        s = self._shape
        if s is not None and s[1] == m and s[2] == d and batch <= s[3] and n_local <= s[0]:
            return
        if s is not None and s[1] == m and s[2] == d:
            n_local, batch = max(n_local, s[0]), max(batch, s[3])
        torch.cuda.synchronize(self.device)
        self._sgpr_state = None
        with torch.cuda.device(self.device):
            check(self.lib.ggp_reserve(self.h, ctypes.byref(self.cfg), int(n_local), int(m), int(d), int(batch)), "ggp_reserve")
            if self._kprog is not None:
                check(self.lib.ggp_set_kernel_program(self.h, ctypes.byref(self._kprog), int(d)), "ggp_set_kernel_program")
        self._shape = (int(n_local), int(m), int(d), int(batch))

    def workspace_bytes(self, n_local, m, d, batch):
        out = ctypes.c_size_t()
        check(self.lib.ggp_workspace_bytes(ctypes.byref(self.cfg), int(n_local), int(m), int(d), int(batch), ctypes.byref(out)),
              "ggp_workspace_bytes")
        return out.value

    # ------------------------------------------------------------------------------------------------------
    def factor(self, Z, theta, jitter_policy="gpytorch", raise_on_fail=True, after_first_enqueue=None):
        """Kzz + jitter I = L L^T with the host-side jitter ladder.  Returns (jitter[batch] tensor, info[batch] cpu).
        after_first_enqueue: called once, right after the first attempt has been enqueued and before its status is read back
        (sgpr_eval enqueues the tile prefetch there, so that the factorisation is already running when the build arrives)."""
        m, d = Z.shape
        batch = theta.shape[0]
        ladder = jitter_ladder(jitter_policy)
        level = [0] * batch
        jit = torch.full((batch,), ladder[0], dtype=torch.float64, device=self.device)
        info = torch.zeros(batch, dtype=torch.int32, device=self.device)
        while True:
            if after_first_enqueue is not None:
                check(self.lib.ggp_sgpr_expect_prefetch(self.h, 1), "ggp_sgpr_expect_prefetch")
            check(self.lib.ggp_sgpr_factor(self.h, ctypes.byref(self.cfg), _stream(), _ptr(Z), _ptr(theta), _ptr(jit),
                                           m, d, batch, _ptr(info)), "ggp_sgpr_factor")
            if after_first_enqueue is not None:
                after_first_enqueue()
                after_first_enqueue = None
            self.ladder_levels = level
            if len(ladder) == 1 and not raise_on_fail:
                # fixed jitter (pymc3 stabilize / gpflow default) and the caller handles info[b] != 0 itself: nothing to retry, so
                # no host read-back -- the evaluation stays asynchronous and can be captured in a CUDA graph (hmc.GraphedTrajectory)
                return jit, info
            info_h = info.cpu()
            bad = [b for b in range(batch) if int(info_h[b]) != 0]
            if not bad:
                return jit, info_h
            if any(level[b] + 1 >= len(ladder) for b in bad):
                if raise_on_fail:
                    raise NotPSDError(f"Kzz not positive definite after jitter ladder {ladder}; potrf info={info_h.tolist()}")
                return jit, info_h
            for b in bad:
                level[b] += 1
            jit = torch.tensor([ladder[l] for l in level], dtype=torch.float64, device=self.device)

    @staticmethod
    def _resolve_group(group):
        """Row sharding is OPT-IN: group=False / None -> X, y are the whole data set on this rank, nothing is reduced (a model
        that passes its full train_x on every rank of a DDP / chain-sharded launch must not have its sums multiplied by the world
        size); group=True -> the default process group; a ProcessGroup -> that group.  Returns (distributed, pg)."""
        import torch.distributed as dist
        if group is None or group is False:
            return False, None
        if group is True:
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("group=True needs an initialised torch.distributed default process group")
            return dist.get_world_size() > 1, None
        return True, group

    # ---- composite kernels ----------------------------------------------------------------------------------
    def kernel_nparams(self, d):
        """Length of a theta row minus the noise entry: d + 1 (ell, sf2) for the single kernels, P of the program otherwise."""
        return d + 1 if self.prog is None else _lib.kprog_layout(self.prog, d)[0]

    def _composite_rows(self, theta, d, need_grad):
        """theta[batch, P + 1] (program parameters, s2) -> the [batch, d + 2] rows the C ABI takes (ell slots unused, sf2 slot =
        k(x, x) = sum of the amplitudes, s2) + the registered parameter / gradient buffers (kept alive on the engine)."""
        P, amp = _lib.kprog_layout(self.prog, d)
        if theta.shape[1] != P + 1:
            raise ValueError(f"composite kernel {self.prog}: theta rows have {P} kernel parameters + the noise variance = {P + 1} entries")
        batch = theta.shape[0]
        kth = theta[:, :P].contiguous()
        std = torch.ones(batch, d + 2, dtype=torch.float64, device=theta.device)
        std[:, d] = sum(kth[:, i] for i in amp)          # no index tensor: stays capturable in a CUDA graph
        std[:, d + 1] = theta[:, P]
        kg = torch.zeros(2, batch, P, dtype=torch.float64, device=theta.device) if need_grad else None
        self._kkeep = (kth, kg)
        return std, kth, kg

    def _set_kparams(self):
        kth, kg = self._kkeep
        check(self.lib.ggp_set_kernel_params(self.h, _ptr(kth), _ptr(kg[0]) if kg is not None else _ptr(None),
                                             _ptr(kg[1]) if kg is not None else _ptr(None)), "ggp_set_kernel_params")

    def sgpr_eval(self, X, y, Z, theta, jitter_policy="gpytorch", need_grad=True, group=False, raise_on_fail=True):
        """Collapsed bound F (not divided by N) and dF/d(ell, sf2, s2, Z) for each row of theta.

        X [n_local, d], y [n_local] are THIS rank's rows.  With `group` (True = default process group, or a ProcessGroup) they are a
        row shard and the partial sums are all-reduced; the default group=False never reduces.
        Returns dict(bound[batch], grad[batch, d+2+m*d] | None, jitter, info, info_b, n_total, path).
        """
        import torch.distributed as dist
        dev = self.device
        Z_in, theta_in = Z, theta
        Z, theta = _f64c(Z, dev), _f64c(theta, dev)
        if theta.dim() == 1:
            theta = theta.unsqueeze(0)
        n_local, d = X.shape
        m = Z.shape[0]
        batch = theta.shape[0]
        kg = None
        if self.prog is not None:
            theta, _, kg = self._composite_rows(theta, d, need_grad)
        assert theta.shape[1] == d + 2 and Z.shape[1] == d and y.shape[0] == n_local
        if (batch > 1 and self.prog is None and self.cfg.precision == PRECISIONS["fp64_i8"] and n_local * m >= self.i8_batch_min_elems
                and m >= 65 and d <= 16 and not torch.cuda.is_current_stream_capturing()):
            # several theta rows (hyper-parameter draws of the stochastic bound, models/bayesian_sgpr_hmc.py:121-134; HMC chains) on a
            # LARGE streamed problem: the sliced-integer plans hold one draw's tiles / digit planes at a time, so the draws are
            # evaluated one after another at the tcgen05 rate (2-3x the batched FP64 DMMA plan); small problems stay batched
            # (they are latency-bound: one launch sequence for all rows wins)
            if not X.is_cuda:
                X, y = _f64c(X, dev), _f64c(y, dev)
            outs = [self.sgpr_eval(X, y, Z, theta[b], jitter_policy, need_grad, group, raise_on_fail) for b in range(batch)]
            cat = lambda k: torch.cat([o[k] for o in outs]) if outs[0][k] is not None else None
            return dict(bound=cat("bound"), grad=cat("grad"), jitter=cat("jitter"), info=torch.cat([torch.as_tensor(o["info"]).reshape(-1).cpu() for o in outs]),
                        info_b=cat("info_b"), n_total=cat("n_total"), partial=cat("partial"),
                        path="fp64_i8" if all(o["path"] == "fp64_i8" for o in outs) else "mixed")
        self.reserve(n_local, m, d, batch)
        if self.prog is not None:
            self._set_kparams()
        with torch.cuda.device(dev):
            cfgp = ctypes.byref(self.cfg)
            use_side = (self.cfg.tile_cache_mib > 0 and n_local >= self.prefetch_min_rows
                        and not torch.cuda.is_current_stream_capturing())
            host_rows = use_side and not X.is_cuda and not y.is_cuda
            if use_side:
                if self._side is None:
                    self._side = torch.cuda.Stream(device=dev)
                main = torch.cuda.current_stream(dev)
                self._side.wait_stream(main)
            ev = None
            if host_rows:
                # HOST rows (pinned or not): upload them piece by piece on a copy stream, each piece followed by the tile build of
                # its rows on the side stream, so that the H2D copy of X (72 MB at the headline shape) overlaps both the tile build
                # of the previous piece and the factorisation (which only needs Z and theta)
                if self._copy is None:
                    self._copy = torch.cuda.Stream(device=dev)
                self._copy.wait_stream(main)
                Xh, yh = X.detach().to(dtype=torch.float64).contiguous(), y.detach().to(dtype=torch.float64).contiguous()
                X = torch.empty(Xh.shape, dtype=torch.float64, device=dev)
                y = torch.empty(yh.shape, dtype=torch.float64, device=dev)
                X.record_stream(self._copy); y.record_stream(self._copy)
                X.record_stream(self._side); y.record_stream(self._side)
                pieces = 4 if (n_local >= 4 * self.prefetch_min_rows and self.cfg.precision == PRECISIONS["fp64_i8"]) else 1
                step = -(-n_local // pieces)
                step = -(-step // 16384) * 16384          # chunk-aligned pieces (any execution plan)
                with torch.cuda.stream(self._copy):
                    y.copy_(yh, non_blocking=True)
                r0 = 0
                while r0 < n_local:
                    nr = min(step, n_local - r0)
                    with torch.cuda.stream(self._copy):
                        X[r0:r0 + nr].copy_(Xh[r0:r0 + nr], non_blocking=True)
                        cev = self._copy.record_event()
                    self._side.wait_event(cev)
                    check(self.lib.ggp_sgpr_prefetch_tiles_part(self.h, cfgp, ctypes.c_void_p(self._side.cuda_stream), _ptr(X), n_local,
                                                                r0, nr, _ptr(Z), _ptr(theta), m, d, batch), "ggp_sgpr_prefetch_tiles_part")
                    r0 += nr
                ev = self._side.record_event()
            else:
                X, y = _f64c(X, dev), _f64c(y, dev)
                # the k(X,Z) tiles do not depend on the Cholesky of Kzz: build them into the tile cache on a side stream while the
                # factorisation (latency-bound m x m kernels) runs on the caller's stream; pass 1 then finds them in place
            prefetch = None
            if use_side and not host_rows:
                # enqueued right AFTER the first factorisation attempt (Engine.factor calls it back): the factorisation is a single
                # cluster launch that takes its SMs first, the build then fills the others
                evbox = []

                def prefetch():
                    check(self.lib.ggp_sgpr_prefetch_tiles(self.h, cfgp, ctypes.c_void_p(self._side.cuda_stream), _ptr(X), n_local,
                                                           _ptr(Z), _ptr(theta), m, d, batch), "ggp_sgpr_prefetch_tiles")
                    evbox.append(self._side.record_event())
            jit, info1 = self.factor(Z, theta, jitter_policy, raise_on_fail, after_first_enqueue=prefetch)
            if prefetch is not None:
                ev = evbox[0]
            on_dmma = self.cfg.precision != PRECISIONS["fp64_i8"]
            if self.cfg.precision == PRECISIONS["fp64_i8"] and any(l > 0 for l in self.ladder_levels):
                # Kzz was numerically singular (exactly duplicated inducing rows: the with-replacement draw of
                # experiments/regression.py:83) and only the ladder's jitter made it factorable: cond(Kzz + jI) >= 1e8 / j.  The
                # sliced-integer products carry an error that is absolute w.r.t. the row scale (2e-16 of sum |a||b|) and cond(Kzz)
                # amplifies it on dF/dZ (measured 2.2e-7 against long double at the headline shape, jitter 1e-8), whereas the FP64
                # DMMA path, relative element by element, holds 2e-9 there: evaluate this call on the DMMA plan (same handle; the
                # prefetched FP64 tiles are reused).
                cfgp = ctypes.byref(self.cfg_dmma)
                on_dmma = True
            if ev is not None:
                torch.cuda.current_stream(dev).wait_event(ev)
            partial = torch.empty(batch, m * m + m + 3, dtype=torch.float64, device=dev)
            check(self.lib.ggp_sgpr_pass1(self.h, cfgp, _stream(), _ptr(X), _ptr(y), n_local, _ptr(Z), _ptr(theta), m, d, batch,
                                          _ptr(partial)), "ggp_sgpr_pass1")
            distributed, pg = self._resolve_group(group)
            if distributed:
                dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=pg)
            P = d + 2 + m * d
            bound = torch.empty(batch, dtype=torch.float64, device=dev)
            grad_mm = torch.empty(batch, P, dtype=torch.float64, device=dev) if need_grad else None
            info2 = torch.zeros(batch, dtype=torch.int32, device=dev)
            check(self.lib.ggp_sgpr_finish(self.h, cfgp, _stream(), _ptr(Z), _ptr(theta), m, d, batch, _ptr(partial),
                                           1 if need_grad else 0, _ptr(bound), _ptr(grad_mm), _ptr(info2)), "ggp_sgpr_finish")
            grad = None
            if need_grad:
                gp = torch.empty(batch, P, dtype=torch.float64, device=dev)
                check(self.lib.ggp_sgpr_pass2(self.h, cfgp, _stream(), _ptr(X), _ptr(y), n_local, _ptr(Z), _ptr(theta), m, d,
                                              batch, _ptr(gp)), "ggp_sgpr_pass2")
                if distributed:
                    dist.all_reduce(gp, op=dist.ReduceOp.SUM, group=pg)
                    if kg is not None:
                        dist.all_reduce(kg[1], op=dist.ReduceOp.SUM, group=pg)
                grad = grad_mm + gp
                if kg is not None:   # [d/d(kernel parameters), d_s2, d_Z]
                    grad = torch.cat([kg[0] + kg[1], grad[:, d + 1:]], dim=1)
            if raise_on_fail:
                info2_h = info2.cpu()
                if bool((info2_h != 0).any()):
                    raise NotPSDError(f"I + A A^T / s not positive definite; potrf info={info2_h.tolist()}")
            self._mark_state(Z_in, theta_in)
        return dict(bound=bound, grad=grad, jitter=jit, info=info1, info_b=info2, n_total=partial[:, -1], partial=partial,
                    path="fp64" if on_dmma else "fp64_i8")

    def sgpr_predict_state(self, X, y, Z, theta, jitter_policy="gpytorch", group=False, train_diag_correction=True):
        """Leave the handle in the state ggp_sgpr_predict reads, for the eval-mode predictive of models/sgpr.py:150-160.

        train_diag_correction=True (gpytorch's eval mode: the kernel is evaluated on the training inputs with the sgpr diagonal
        correction on): training row n carries Lambda_n = s2 + max(k_nn - q_nn, 0).  False: plain s2 on every training row (the
        state sgpr_eval leaves).  Returns the key sgpr_predict checks."""
        dev = self.device
        X, y, Z, theta = (_f64c(t, dev) for t in (X, y, Z, theta))
        if theta.dim() == 1:
            theta = theta.unsqueeze(0)
        n_local, d = X.shape
        m, batch = Z.shape[0], theta.shape[0]
        theta_ext = theta
        if self.prog is not None:
            theta, _, _ = self._composite_rows(theta, d, False)
        self.reserve(n_local, m, d, batch)
        if self.prog is not None:
            self._set_kparams()
        self._sgpr_state = None
        with torch.cuda.device(dev):
            cfgp = ctypes.byref(self.cfg_dmma)
            jit, _ = self.factor(Z, theta, jitter_policy, True)
            partial = torch.empty(batch, m * m + m + 3, dtype=torch.float64, device=dev)
            fn = self.lib.ggp_sgpr_predict_pass1 if train_diag_correction else self.lib.ggp_sgpr_pass1
            check(fn(self.h, cfgp, _stream(), _ptr(X), _ptr(y), n_local, _ptr(Z), _ptr(theta), m, d, batch, _ptr(partial)),
                  "ggp_sgpr_predict_pass1")
            distributed, pg = self._resolve_group(group)
            if distributed:
                import torch.distributed as dist
                dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=pg)
            bound = torch.empty(batch, dtype=torch.float64, device=dev)
            info2 = torch.zeros(batch, dtype=torch.int32, device=dev)
            check(self.lib.ggp_sgpr_finish(self.h, cfgp, _stream(), _ptr(Z), _ptr(theta), m, d, batch, _ptr(partial), 0, _ptr(bound),
                                           _ptr(None), _ptr(info2)), "ggp_sgpr_finish")
            if bool((info2 != 0).any()):
                raise NotPSDError(f"I + A W A^T not positive definite; potrf info={info2.tolist()}")
        self._mark_state(Z, theta_ext, by_value=True)
        return dict(jitter=jit)

    def sgpr_predict(self, Xs, Z, theta, full_cov=False, add_noise=True):
        """Predictive at the state left by the last sgpr_predict_state (or sgpr_eval) with the same (Z, theta)."""
        dev = self.device
        self._check_state(Z, theta)
        Xs, Z, theta = _f64c(Xs, dev), _f64c(Z, dev), _f64c(theta, dev)
        if theta.dim() == 1:
            theta = theta.unsqueeze(0)
        ns, d = Xs.shape
        m = Z.shape[0]
        batch = theta.shape[0]
        if self.prog is not None:
            theta, _, _ = self._composite_rows(theta, d, False)
            self._set_kparams()
        mean = torch.empty(batch, ns, dtype=torch.float64, device=dev)
        var = torch.empty(batch, ns, dtype=torch.float64, device=dev)
        cov = torch.empty(batch, ns, ns, dtype=torch.float64, device=dev) if full_cov else None
        with torch.cuda.device(dev):
            check(self.lib.ggp_sgpr_predict(self.h, ctypes.byref(self.cfg), _stream(), _ptr(Xs), ns, _ptr(Z), _ptr(theta), m, d,
                                            batch, 1 if add_noise else 0, _ptr(mean), _ptr(var), _ptr(cov)), "ggp_sgpr_predict")
        return mean, var, cov

    def svgp_eval(self, xb, yb, Z, qm, qLs, theta, num_data=None, likelihood="gaussian", jitter_policy="gpytorch",
                  base_jitter=1e-6, data_jitter=1e-4, lik_scale=None, kl_scale=None, need_grad=True, raise_on_fail=True):
        """Whitened SVGP ELBO (models/svgp.py:104-106) or, with qLs=None, the SGPMC conditional log-likelihood
        (models/sgp_hmc.py:63).  Returns dict(value[batch], grad[batch, d+2+m*d+m+m*m] | None, jitter, info).

        jitter: total Kzz jitter = base_jitter (gpytorch variational_cholesky_jitter 1e-6 / gpflow 1e-5) + ladder level.
        Defaults follow gpytorch: lik_scale = 1/nb, kl_scale = 1/num_data, data_jitter = 1e-4."""
        dev = self.device
        xb, yb, Z, qm, theta = (_f64c(t, dev) for t in (xb, yb, Z, qm, theta))
        qLs = _f64c(qLs, dev) if qLs is not None else None
        if theta.dim() == 1:
            theta = theta.unsqueeze(0)
        nb, d = xb.shape
        m = Z.shape[0]
        batch = theta.shape[0]
        qm_batched = qm.dim() == 2      # one whitened vector per batch element (SGPMC chains): [batch, m]
        assert qm.shape == ((batch, m) if qm_batched else (m,))
        self._sgpr_state = None         # the SVGP scratch aliases the SGPR m x m state
        if lik_scale is None:
            lik_scale = 1.0 / nb
        if kl_scale is None:
            kl_scale = 0.0 if (qLs is None or num_data is None) else 1.0 / float(num_data)
        self.reserve(min(nb, 4096), m, d, batch)
        ladder = [base_jitter + float(j) for j in jitter_ladder(jitter_policy)]
        level = [0] * batch
        P = d + 2 + m * d + m + m * m
        value = torch.empty(batch, dtype=torch.float64, device=dev)
        grad = torch.empty(batch, P, dtype=torch.float64, device=dev) if need_grad else None
        info = torch.zeros(batch, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            while True:
                jit = torch.tensor([ladder[l] for l in level], dtype=torch.float64, device=dev)
                check(self.lib.ggp_svgp_elbo(self.h, ctypes.byref(self.cfg), _stream(), _ptr(xb), _ptr(yb), nb, _ptr(Z), _ptr(qm),
                                             1 if qm_batched else 0, _ptr(qLs), _ptr(theta), _ptr(jit), m, d, batch, float(data_jitter), float(lik_scale),
                                             float(kl_scale), LIKELIHOODS[likelihood], 1 if need_grad else 0, _ptr(value),
                                             _ptr(grad), _ptr(info)), "ggp_svgp_elbo")
                if len(ladder) == 1 and not raise_on_fail:
                    # fixed jitter (gpflow's default) and the caller masks failed rows itself: no host read-back, the call stays
                    # asynchronous (HMC leapfrogs queue back to back, CUDA-graph capturable)
                    return dict(value=value, grad=grad, jitter=jit, info=info)
                info_h = info.cpu()
                bad = [b for b in range(batch) if int(info_h[b]) != 0]
                if not bad:
                    break
                if any(level[b] + 1 >= len(ladder) for b in bad):
                    if raise_on_fail:
                        raise NotPSDError(f"Kzz not positive definite after jitter ladder {ladder}; potrf info={info_h.tolist()}")
                    break
                for b in bad:
                    level[b] += 1
        return dict(value=value, grad=grad, jitter=jit, info=info_h)

    def svgp_predict(self, xs, Z, qm, qLs, theta, base_jitter=1e-6, data_jitter=1e-4, add_noise=True):
        """Predictive marginals of the whitened strategy (models/svgp.py:132-141): (mean[batch,ns], var[batch,ns])."""
        dev = self.device
        xs, Z, qm, theta = (_f64c(t, dev) for t in (xs, Z, qm, theta))
        qLs = _f64c(qLs, dev) if qLs is not None else None
        if theta.dim() == 1:
            theta = theta.unsqueeze(0)
        ns, d = xs.shape
        m, batch = Z.shape[0], theta.shape[0]
        qm_batched = qm.dim() == 2
        self._sgpr_state = None
        self.reserve(min(ns, 4096), m, d, batch)
        mean = torch.empty(batch, ns, dtype=torch.float64, device=dev)
        var = torch.empty(batch, ns, dtype=torch.float64, device=dev)
        info = torch.zeros(batch, dtype=torch.int32, device=dev)
        for j in jitter_ladder("gpytorch"):
            jit = torch.full((batch,), base_jitter + j, dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                check(self.lib.ggp_svgp_predict(self.h, ctypes.byref(self.cfg), _stream(), _ptr(xs), ns, _ptr(Z), _ptr(qm),
                                                1 if qm_batched else 0, _ptr(qLs),
                                                _ptr(theta), _ptr(jit), m, d, batch, float(data_jitter), 1 if add_noise else 0,
                                                _ptr(mean), _ptr(var), _ptr(info)), "ggp_svgp_predict")
            if not bool((info != 0).any()):
                return mean, var
        raise NotPSDError("Kzz not positive definite after the jitter ladder")

    # building blocks -----------------------------------------------------------------------------------------
    def chol(self, a, want_inverse=True):
        a = a.clone().contiguous()
        batch, m, _ = a.shape
        linv = torch.empty_like(a) if want_inverse else None
        info = torch.zeros(batch, dtype=torch.int32, device=a.device)
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            check(self.lib.ggp_chol_batched(self.h, _stream(), _ptr(a), _ptr(linv), m, batch, _ptr(info)), "ggp_chol_batched")
        self._shape = None  # chol may have re-reserved the handle for its own shape
        self._sgpr_state = None
        return a, linv, info

    def gemm_nt(self, A, B, C=None, alpha=1.0, beta=0.0):
        mm, kk = A.shape
        nn = B.shape[0]
        if C is None:
            C = torch.zeros(mm, nn, dtype=torch.float64, device=A.device)
        with torch.cuda.device(self.device):
            check(self.lib.ggp_gemm_nt(self.h, _stream(), _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(C), C.stride(0),
                                       mm, nn, kk, float(alpha), float(beta)), "ggp_gemm_nt")
        return C

    def gemm_nt_ex(self, A, B, C, alpha=1.0, beta=0.0, kmode=0, sym=0, splits=1, split_stride=0):
        mm, kk = A.shape
        nn = B.shape[0]
        with torch.cuda.device(self.device):
            check(self.lib.ggp_gemm_nt_ex(self.h, _stream(), _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(C), C.stride(-2),
                                          mm, nn, kk, float(alpha), float(beta), int(kmode), int(sym), int(splits),
                                          int(split_stride)), "ggp_gemm_nt_ex")
        return C

    def gemm_nt_i8(self, A, B):
        """C = A B^T through the sliced-integer tcgen05 path (FP64-class accuracy; tests / probes)."""
        A, B = _f64c(A, self.device), _f64c(B, self.device)
        C = torch.empty(A.shape[0], B.shape[0], dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.ggp_gemm_nt_i8(self.h, _stream(), _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(C), C.stride(0),
                                          A.shape[0], B.shape[0], A.shape[1]), "ggp_gemm_nt_i8")
        return C

    def kernel_matrix(self, X1, X2, theta):
        dev = self.device
        X1, X2, theta = _f64c(X1, dev), _f64c(X2, dev), _f64c(theta, dev)
        out = torch.empty(X1.shape[0], X2.shape[0], dtype=torch.float64, device=dev)
        if self.prog is not None:     # the program is registered per input dimension on a reserved handle
            d = X1.shape[1]
            if self._shape is None or self._shape[2] != d:
                self.reserve(64, 64, d, 1)
            theta, _, _ = self._composite_rows(theta.reshape(1, -1), d, False)
            self._set_kparams()
        with torch.cuda.device(dev):
            check(self.lib.ggp_kernel_matrix(self.h, ctypes.byref(self.cfg), _stream(), _ptr(X1), X1.shape[0], _ptr(X2),
                                             X2.shape[0], _ptr(theta), X1.shape[1], _ptr(out)), "ggp_kernel_matrix")
        return out

    PROFILE_CATEGORIES = ("build", "trmm", "syrk", "bwd", "mm", "other")

    def profile_enable(self, on=True):
        check(self.lib.ggp_profile_enable(self.h, 1 if on else 0), "ggp_profile_enable")

    def profile_read(self):
        """Synchronise and return ({category: ms}, {category: spans}, kernel launches) since the last read."""
        ms = (ctypes.c_double * 6)()
        cnt = (ctypes.c_int64 * 6)()
        launches = ctypes.c_int64()
        check(self.lib.ggp_profile_read(self.h, ms, cnt, ctypes.byref(launches)), "ggp_profile_read")
        cats = self.PROFILE_CATEGORIES
        return {c: ms[i] for i, c in enumerate(cats)}, {c: cnt[i] for i, c in enumerate(cats)}, launches.value

    def probe_dmma_peak(self, iters=20000):
        out = (ctypes.c_double * 4)()
        with torch.cuda.device(self.device):
            check(self.lib.ggp_probe_dmma_peak(self.h, _stream(), int(iters), out), "ggp_probe_dmma_peak")
        return dict(best=out[0], warps8=out[1], warps16=out[2], warps32=out[3])

    def probe_i8_peak(self, iters=4000):
        """Measured tcgen05 kind::i8 tensor-pipe rate [int8 TOP/s] with shared-memory-resident operands (no loads, no epilogue)."""
        out = (ctypes.c_double * 4)()
        with torch.cuda.device(self.device):
            check(self.lib.ggp_probe_i8_peak(self.h, _stream(), int(iters), out), "ggp_probe_i8_peak")
        return dict(best=out[0], uniform_n256=out[1], mix_depth2=out[2], mix_bn32=out[3])
