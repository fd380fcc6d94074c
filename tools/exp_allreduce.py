"""All-reduce of [d|E|EQ] of the sharded 98k-atom box alone, and the sharded step (design experiment, under torchrun)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from q6_b200 import synth
from q6_b200.engine import Qnb
from q6_b200.system import shard_system
q, cuts, lam = synth.config("C5")
for mode in (sys.argv[1:] or ["p2p", "nccl"]):
    g = Qnb(shard_system(q, rank, world), device=lr)
    m = g.comm_connect(rank, world, dist, mode)
    g.make_pair_lists(q.xtop, **cuts)
    g.pot_energy_nonbonds(q.xtop, lam)
    dist.barrier()
    ar = g.bench_allreduce(200)
    dist.barrier()
    step = g.bench_nonbond(lam, 200) / 200 * 1e3
    dist.barrier()
    build = g.bench_build_lists(3) / 3
    g.comm_status()
    t = torch.tensor([ar * 1e3, step, build], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if m == "p2p":
        dist.barrier()
        g.bench_allreduce(20)       # every rank takes part
    if rank == 0 and m == "p2p":
        print("   phases of the last all-reduce on rank 0:", {k: round(v, 1) for k, v in g.allreduce_phases().items()}, flush=True)
    if rank == 0:
        print(f"N={world} {m}: allreduce {t[0].item():.1f} us, sharded step {t[1].item():.1f} us, build {t[2].item():.3f} ms", flush=True)
    g.close()
dist.destroy_process_group()
