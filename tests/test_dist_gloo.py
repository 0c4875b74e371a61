"""world_size-2 gloo test of the row-sharding host logic: per-rank partial sums of pass 1 / pass 2 are summed
by ONE all-reduce each and every rank finishes redundantly.  On CPU the per-rank partials come from the oracle (tests may
use it as the checker); on the GPU box the same reduce is exercised with NCCL by bench.py --gpus 2."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _partial(X, y, Z, th, jit):
    """[A A^T | A y | y^T y, sum k_nn, n] for a row shard, from the oracle (same layout as ggp_sgpr_pass1)."""
    from oracle.kernels import ard_kernel
    D, M = X.shape[1], Z.shape[0]
    L = torch.linalg.cholesky(ard_kernel(Z, Z, th[:D], th[D]) + jit * torch.eye(M, dtype=torch.float64))
    A = torch.linalg.solve_triangular(L, ard_kernel(Z, X, th[:D], th[D]), upper=False)
    return torch.cat([(A @ A.T).reshape(-1), A @ y, torch.stack([y @ y, X.shape[0] * th[D], torch.tensor(float(X.shape[0]), dtype=torch.float64)])])


def _finish(partial, th, M, D):
    import math
    S, b = partial[:M * M].view(M, M), partial[M * M:M * M + M]
    yty, sumk, N = partial[-3], partial[-2], partial[-1]
    s2 = th[D + 1]
    LB = torch.linalg.cholesky(torch.eye(M, dtype=torch.float64) + S / s2)
    c = torch.linalg.solve_triangular(LB, b.unsqueeze(-1), upper=False).squeeze(-1) / s2
    return (-0.5 * N * math.log(2 * math.pi) - 0.5 * N * torch.log(s2) - torch.log(torch.diagonal(LB)).sum()
            - 0.5 * (yty / s2 - c @ c) - 0.5 * (sumk - torch.trace(S)) / s2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_problem
    import ggp_b200.dist as gd
    from oracle import sgpr as osgpr
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, M, D = 1001, 24, 3
    X, y, Z, th = make_problem(N, M, D, seed=3)
    lo, hi = gd.shard_rows(N, rank, world)
    part = _partial(X[lo:hi], y[lo:hi], Z, th, 1e-6)
    gd.allreduce_sum_(part)
    F = _finish(part, th, M, D)
    Fo = osgpr.sgpr_bound(X, y, Z, th[:D], th[D], th[D + 1], 1e-6, "none")
    ok = abs(F - Fo) < 1e-11 * abs(Fo) and int(part[-1].item()) == N
    # chains sharded across ranks: no collective in the sampling loop, only the final gather
    chains = gd.shard_chains(7, rank, world)
    gathered = gd.gather_chains(torch.tensor(chains, dtype=torch.float64), 7)
    ok = ok and gathered.tolist() == [float(i) for i in range(7)]
    q.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def test_row_sharded_bound_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(r[1] for r in res), res
    assert res[0][2][0] == 0 and res[0][2][1] == res[1][2][0] and res[1][2][1] == 1001


def test_shard_rows_covers_everything_exactly_once():
    import ggp_b200.dist as gd
    for N in [0, 1, 7, 1000, 1_000_000]:
        for world in [1, 2, 3, 8]:
            spans = [gd.shard_rows(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
