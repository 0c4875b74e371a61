"""Run under torchrun on N GPUs: the row-sharded evaluation (NCCL all-reduce of the partial sums) must agree with the
unsharded one to ~1e-12 relative (only the summation order differs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.distributed as dist
import ggp_b200, ggp_b200.dist as gd
from helpers import make_problem, relerr
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, M, D = 100_003, 512, 8
X, y, Z, th = make_problem(N, M, D, seed=5)
eng = ggp_b200.Engine.get(dev, precision=os.environ.get("GGP_CHECK_PRECISION", "fp64_i8"))
lo, hi = gd.shard_rows(N, rank, world)
out = eng.sgpr_eval(X[lo:hi], y[lo:hi], Z, th, jitter_policy=1e-6, group=True)   # row shard: default NCCL group
full = eng.sgpr_eval(X, y, Z, th, jitter_policy=1e-6, group=False)
eb, eg = relerr(out["bound"], full["bound"]), relerr(out["grad"], full["grad"])
gathered = [torch.zeros_like(out["bound"]) for _ in range(world)]
dist.all_gather(gathered, out["bound"])
same = all(torch.equal(g, gathered[0]) for g in gathered)
if rank == 0:
    print(f"world={world} rows/rank={hi-lo} bound relerr={eb:.2e} grad relerr={eg:.2e} identical across ranks={same} N_total={int(out['n_total'][0])}")
    assert eb < 1e-12 and eg < 1e-9 and same and int(out["n_total"][0]) == N
dist.destroy_process_group()
