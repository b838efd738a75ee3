"""Config 4's pairwise step on G GPUs: NCCL all-gather of positions + scp_b200_linearize_range on each rank's pair
share.  Launch: python -m torch.distributed.run --nproc-per-node G tools/bench_sharded_linearize.py [N K]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ba-path-planning_b200"))
import torch
import torch.distributed as dist
from path_planning.solvers.sharded import ShardedLinearizer

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sl = ShardedLinearizer(N, K, 0.8)
g = torch.Generator(device="cuda"); g.manual_seed(1)
pos_full = torch.rand((N, K, 2), dtype=torch.float64, device="cuda", generator=g) * 130
own = pos_full[sl.lo:sl.hi].contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for _ in range(3): sl.linearize(own)
barrier()
ts, tg = [], []
for r in range(10):
    flush.fill_(r)
    e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
    e0.record(); sl.gather_positions(own); e1.record(); torch.cuda.synchronize()
    tg.append(e0.elapsed_time(e1))
    barrier(); flush.fill_(r + 100)
    e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
    e0.record(); out = sl.linearize(own); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = torch.tensor([float(np.median(ts)), float(np.median(tg))], dtype=torch.float64, device="cuda")
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
P = N * (N - 1) // 2
if rank == 0:
    nbytes = 8 * (2 * N * K + 3 * P * K)
    print(json.dumps(dict(case="sharded linearize", N=N, K=K, n_gpus=world, ms_linearize_incl_allgather_and_minreduce=float(t[0]),
                          ms_allgather=float(t[1]), rows=P * K, algorithmic_bytes=nbytes, min_separation=sl.decode(out[2])[0], first_violation=sl.decode(out[2])[1],
                          agent_bounds=sl.bounds, aggregate_gbs=nbytes / float(t[0]) / 1e6)))
if world > 1: dist.destroy_process_group()
