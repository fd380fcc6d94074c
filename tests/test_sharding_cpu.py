"""i-range sharding (distribute_nonbonds) on CPU: the union of the shards' lists is the full list and the
all-reduced forces/energies equal the single-process result.  Runs world_size 2 over gloo."""
import os
import sys

import numpy as np
import pytest

import common


def test_distribute_nonbonds_ranges():
    from q6_b200.system import distribute_nonbonds
    c = np.array([5, 0, 3, 9, 1, 1, 7, 2])
    r = distribute_nonbonds(c, 3)
    assert r[0][0] == 1 and r[-1][1] == len(c)
    assert all(r[k][1] + 1 == r[k + 1][0] for k in range(2))
    sums = [c[a - 1:b].sum() for a, b in r]
    assert max(sums) <= c.sum() / 3 + c.max()
    assert distribute_nonbonds(np.ones(4), 8)[-1][1] == 4        # more ranks than items: empty tail ranges


def _worker(rank, world, port, q_path, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.pyoracle import Oracle
    from q6_b200.system import QSystem, shard_system
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    q = QSystem.load(q_path)
    s = shard_system(q, rank, world)
    o = Oracle(s)
    cuts = common.sph_cuts(8.0)
    counts = o.make_pair_lists(q.xtop, **cuts)
    # lrf_gather: all-reduce of the moments (centres are identical on every rank)
    lrf = torch.from_numpy(o.export_lrf().copy())
    cent = lrf[:, :3].clone()
    dist.all_reduce(lrf)
    lrf[:, :3] = cent
    lam = np.array([0.4, 0.6])
    d, E, EQ = o.pot_energy_nonbonds(q.xtop, lam, qq=(rank == 0))
    buf = torch.from_numpy(np.concatenate([d.reshape(-1), E, EQ.reshape(-1)]))
    dist.all_reduce(buf)          # gather_nonbond + sum (potene.f90:195-222)
    cnt = torch.from_numpy(counts.astype(np.int64).copy())
    dist.all_reduce(cnt)
    if rank == 0:
        np.savez(out_path, buf=buf.numpy(), cnt=cnt.numpy(), lrf=lrf.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single(tmp_path):
    import torch.multiprocessing as mp
    from oracle.pyoracle import Oracle
    from q6_b200 import synth
    q = synth.solvated_sphere(14.0, 8.0, 12, 2, 51, fep="evb")
    q.use_LRF = 0      # the Taylor term needs the reduced moments before the force step; covered on GPU with NCCL
    q_path = str(tmp_path / "q.npz")
    out_path = str(tmp_path / "out.npz")
    q.save(q_path)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, q_path, out_path), nprocs=2, join=True)
    z = np.load(out_path)
    o = Oracle(q)
    counts = o.make_pair_lists(q.xtop, **common.sph_cuts(8.0))
    d, E, EQ = o.pot_energy_nonbonds(q.xtop, np.array([0.4, 0.6]))
    ref = np.concatenate([d.reshape(-1), E, EQ.reshape(-1)])
    assert np.array_equal(z["cnt"][:5], counts[:5])         # union of the shards == full lists
    assert np.allclose(z["buf"], ref, rtol=1e-10, atol=1e-9)
