"""i-range sharding across GPUs with the NCCL all-reduce of [d|E|EQ] and of the LRF moments
(gather_nonbond / lrf_gather).  Needs >= 2 GPUs; the host logic is covered on CPU by test_sharding_cpu.py."""
import os
import sys

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q_path, out_path, cuts, lam, shard_rows=False, comm="nccl"):
    if not shard_rows:
        os.environ["QNB_SHARD_PAIRS"] = "1"     # read by qnb_init: the reference's pair partition instead of the row partition
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from q6_b200.engine import Qnb
    from q6_b200.system import QSystem, shard_system
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    q = QSystem.load(q_path)
    g = Qnb(shard_system(q, rank, world), device=rank)
    # MPI_Bcast of the 128-byte NCCL id / MPI_Allgather of the IPC descriptors in the Fortran host
    assert g.comm_connect(rank, world, dist, comm) == comm
    counts = g.make_pair_lists(q.xtop, **cuts)
    d, E, EQ = g.pot_energy_nonbonds(q.xtop, lam)
    for _ in range(3):      # the step's graph (it holds the peer-memory all-reduce) is replayed: same sums every time
        d2, E2, _ = g.pot_energy_nonbonds(q.xtop, lam)
        assert np.allclose(d2, d, rtol=1e-9, atol=1e-9) and np.allclose(E2, E, rtol=1e-9, atol=1e-9)
    g.comm_status()
    cnt = torch.from_numpy(counts.copy())
    dist.all_reduce(cnt)
    np.savez(out_path + f".{rank}.npz", d=d, E=E, EQ=EQ, cnt=cnt.numpy(), lrf=g.export_lrf())
    g.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("comm", ["nccl", "p2p"])
@pytest.mark.parametrize("shard_rows", [False, True], ids=["pair_partition", "row_partition"])
@pytest.mark.parametrize("case", ["sphere_q", "water_box"])
def test_sharded_allreduce_matches_oracle(case, shard_rows, comm, tmp_path):
    import torch
    import torch.multiprocessing as mp
    from oracle.pyoracle import Oracle
    from q6_b200 import synth
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    if case == "sphere_q":
        q = synth.solvated_sphere(16.0, 9.0, 20, 2, 61, fep="evb")
        cuts, lam = common.sph_cuts(8.0), np.array([0.4, 0.6])
    else:
        q = synth.water_box(10, 62)
        cuts = dict(Rq=-1.0, Rcq2=1.0, RcLRF2=13.0 ** 2, Rcpp2=81.0, Rcpw2=81.0, Rcww2=81.0, RcLRF=13.0)
        lam = np.array([1.0])
    q_path, out_path = str(tmp_path / "q.npz"), str(tmp_path / "out")
    q.save(q_path)
    port = 29600 + (os.getpid() % 2000)
    port += (7 if shard_rows else 0) + (13 if comm == "p2p" else 0)
    mp.spawn(_worker, args=(world, port, q_path, out_path, cuts, lam, shard_rows, comm), nprocs=world, join=True)
    o = Oracle(q)
    counts = o.make_pair_lists(q.xtop, **cuts)
    d, E, EQ = o.pot_energy_nonbonds(q.xtop, lam)
    lrf = o.export_lrf()
    for r in range(world):
        z = np.load(out_path + f".{r}.npz")
        assert np.array_equal(z["cnt"][:5], counts[:5])           # union of the shards == the reference lists
        assert common.rel_rms(z["d"], d) <= common.FORCE_REL_RMS  # every rank holds the summed gradient
        common.assert_energy("E", z["E"], E)
        common.assert_energy("EQ", z["EQ"], EQ)
        scale = np.abs(lrf).max(axis=0) + 1e-300
        tol = np.where(np.arange(43) < 16, 1e-9, 2e-5)
        assert np.all(np.abs(z["lrf"] - lrf) <= tol * scale + 1e-12)
