#!/usr/bin/env python
"""Benchmark of the Qdyn6 nonbonded hot path (BASELINE.json metric) -- see DESIGN.md "Measurement".

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2|C3|C4s|C5]

A "step" is one MD step's share of the hot path on one synthetic system: one nonbonded evaluation
(qnb_nonbond) plus a pair-list rebuild every NBcycle = 25 steps, as md_run does (md.f90:1661,1742).
`value`  = pair interactions/s with coordinates already resident in HBM (CUDA events, max over ranks, mean of
           `--repeats` windows of exactly K steps each; the MD step counter -- list build every 25 steps -- runs on
           across the windows);
`e2e`    = the same through the C ABI with HOST buffers (x up, d + energies down inside the timed region).
N > 1    = one independent replica / lambda window per GPU, no data-path collective (weak scaling); the same line
           carries `sharded_c5` (ONE 98k-atom periodic system, rows sharded over the N GPUs, NCCL all-reduce of
           forces and energies: strong scaling) and `fep_farm` (51 two-state lambda windows scheduled over the GPUs).
--impl reference times the CPU oracle (the reference's algorithm, decomposed over all host threads like Qdyn6p) on
the same workload -- N replicas at --gpus N, so both arms do equal work; the Fortran reference itself cannot be built
here (no Fortran compiler, profiles/r02c_fortran_probe.txt).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NBCYCLE = 25            # tests/basic_tests/*.inp: non_bond 25
DT_FS = 2.0             # dc*.inp stepsize 2.0
STEPS_PER_WINDOW = 5000  # exclude_tests/inputs/excl/gen_inps.pl
FEP_WINDOWS = 51        # gen_inps.pl: lambda 1.00 -> 0.00 in steps of 0.02
# algorithmic flop per unit (SURVEY.md 8d / BASELINE.md 3)
FLOP_PAIR, FLOP_WW_MOL, FLOP_LRF_UPDATE, FLOP_LRF_TAYLOR = 33.0, 209.0, 80.0, 125.0
WORKLOADS = {"C2": "C2: 32 A TIP3P sphere, protein-like core + 46 Q-atoms, LRF, 10 A cut-offs, 1 state",
             "C3": "C3: 25 A water sphere, 46-Q-atom ligand annihilation FEP, 2 states",
             "C4s": "C4s: 25 A sphere, 60 Q-atoms, two-state EVB", "C5": "C5: 32768-water periodic box"}


def flop_q_pair(nstates):
    return 21.0 + 20.0 * nstates


def pairs_per_step(counts, nstates, nqq_total):
    """Sum of list entries evaluated per step: nbpp+nbpw+nbww+(nbqp+nbqw)*nstates+nbqq+nbqqp."""
    return int(counts[0] + counts[1] + counts[2] + (counts[3] + counts[4]) * nstates + nqq_total)


def step_flop(counts, nstates, nqq, natom):
    """Algorithmic flop of one nonbonded evaluation, per kernel (every listed pair once; a pw pair is split half/half
    between the two row kernels that each accumulate one side of it)."""
    ns_q = flop_q_pair(nstates)
    return {"k_water_rows": counts[2] / 9.0 * FLOP_WW_MOL + 0.5 * counts[1] * FLOP_PAIR,
            "k_solute_rows": counts[0] * FLOP_PAIR + 0.5 * counts[1] * FLOP_PAIR,
            "k_q_partner": 0.5 * (counts[3] + counts[4]) * ns_q, "k_q_atom": 0.5 * (counts[3] + counts[4]) * ns_q,
            "k_qq_static": nqq * flop_q_pair(1), "k_lrf_taylor": natom * FLOP_LRF_TAYLOR,
            "k_pair_energy": 0.0,   # the energy sums are part of the 33 / 209 flop per pair counted with the row kernels
            "k_solvent_restraints": 0.0}


def lrf_interactions(q, cuts):
    """(source atom, target group) pairs lrf_update visits per list build: unit pairs outside the list cut-off and inside
    RcLRF, counted with a KD-tree over the switch atoms (periodic when the system is), weighted by the source unit's
    number of non-Q atoms.  An estimate for mixed systems (one cut-off per pair class is taken as Rcww), exact for
    water boxes; used only as the numerator of roofline.lrf."""
    try:
        from scipy.spatial import cKDTree
    except Exception:
        return None
    cgp = np.asarray(q.cgp).reshape(-1, 3)
    sw = cgp[:, 0] - 1
    x = np.asarray(q.xtop).reshape(-1, 3)
    pos = x[sw]
    nq = np.zeros(len(sw))
    iq = np.asarray(q.iqatom).reshape(-1) if q.iqatom is not None and q.nqat else np.zeros(q.natom, int)
    atoms = np.asarray(q.cgpatom).reshape(-1) - 1
    excl = np.asarray(q.excl).reshape(-1) if q.excl is not None else np.zeros(q.natom, int)
    keep = np.ones(len(sw), bool)
    for g in range(len(sw)):
        a = atoms[cgp[g, 1] - 1:cgp[g, 2]]
        nq[g] = float((iq[a] == 0).sum())
        keep[g] = not excl[sw[g]]
    pos, nq = pos[keep], nq[keep]
    box = None
    if q.use_PBC:
        box = np.asarray(q.boxlength, float)
        pos = pos - box * np.floor(pos / box)
    t = cKDTree(pos, boxsize=box)
    rl = float(cuts["RcLRF"]) if cuts["RcLRF"] > 0 else float(np.sqrt(max(cuts["RcLRF2"], 0.0)))
    rc = float(np.sqrt(cuts["Rcww2"]))
    w = (np.ones(len(pos)), nq)
    n = t.count_neighbors(t, [rc, rl], weights=w, cumulative=True)
    return float(n[1] - n[0])


def clocks_sampler(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        f = open(path, "w")
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                stdout=f, stderr=subprocess.DEVNULL), f
    except Exception:
        return None, None


def clocks_summary(path, devices):
    """Median SM clock under load and throttle reasons over the GPUs of this job (one sampler, on rank 0: every
    nvidia-smi query takes driver locks, so one process per rank would perturb the step it is watching)."""
    sm, mx, reasons = [], 0.0, set()
    try:
        for line in open(path):
            p = [t.strip() for t in line.split(",")]
            if len(p) < 9 or not p[0].isdigit() or int(p[0]) not in devices:
                continue
            try:
                sm.append(float(p[1]))
                mx = max(mx, float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    except OSError:
        pass
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
            "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def spread(v):
    v = sorted(float(t) for t in v)
    return {"n": len(v), "min": v[0], "median": float(np.median(v)), "max": v[-1]}


def cpu_reference(q, cuts, lam, threads, target_seconds=12.0):
    """The reference's algorithm on the host cores (oracle, decomposed like Qdyn6p)."""
    from oracle import pyoracle
    cut7 = [cuts[k] for k in ("Rq", "Rcq2", "RcLRF2", "Rcpp2", "Rcpw2", "Rcww2", "RcLRF")]
    r = pyoracle.time_decomposed(q, q.xtop, lam, cut7, threads, 2)          # calibrate
    per_step = max(r["seconds"] / 2, 1e-6)
    steps = int(max(4, min(400, (target_seconds - r["list_seconds"]) / per_step)))
    r = pyoracle.time_decomposed(q, q.xtop, lam, cut7, threads, steps)
    t_step = r["seconds"] / steps + r["list_seconds"] / NBCYCLE
    return t_step, steps, r


def cpu_reference_replicas(q, cuts, lam, threads, replicas, target_seconds):
    """`replicas` copies of the system advanced concurrently on the host, threads split evenly: the CPU arm of a run
    that advances one system per GPU.  Returns (seconds per step of the slowest replica, steps, threads per replica)."""
    if replicas <= 1:
        t, n, _ = cpu_reference(q, cuts, lam, threads, target_seconds)
        return t, n, threads
    from concurrent.futures import ThreadPoolExecutor
    per = max(1, threads // replicas)
    with ThreadPoolExecutor(replicas) as ex:       # ctypes releases the GIL inside the oracle
        res = list(ex.map(lambda _: cpu_reference(q, cuts, lam, per, target_seconds), range(replicas)))
    return max(r[0] for r in res), min(r[1] for r in res), per


def compiled_host_e2e(q, cuts, lam, steps, device):
    """The same end-to-end loop (host buffers, list build every NBCYCLE steps) driven by the compiled host
    q6_b200/host/qdyn_nb from written topology / FEP files: what a Fortran/C++ host pays, without Python in the loop."""
    import re
    import tempfile
    from q6_b200 import synth
    exe = os.path.join(ROOT, "q6_b200", "host", "qdyn_nb")
    with tempfile.TemporaryDirectory() as tmp:
        top, fep = os.path.join(tmp, "s.top"), os.path.join(tmp, "s.fep")
        synth.write_files(q, top, fep)
        cmd = [exe, top, fep if q.nqat else "-", "--no-shake", "--q_atom", repr(float(cuts["Rq"])),
               "--lrf", repr(float(cuts["RcLRF"])), "--solute_solute", repr(float(np.sqrt(cuts["Rcpp2"]))),
               "--solute_solvent", repr(float(np.sqrt(cuts["Rcpw2"]))), "--solvent_solvent", repr(float(np.sqrt(cuts["Rcww2"]))),
               "--lambda", ",".join(repr(float(v)) for v in lam), "--steps", str(int(steps)), "--non_bond", str(NBCYCLE),
               "--device", str(int(device))]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    m = re.search(r"([0-9.]+) ms per step", out.stdout)
    if out.returncode != 0 or not m:
        raise RuntimeError((out.stdout + out.stderr)[-300:])
    return float(m.group(1))


class Job:
    """Process-group plumbing of one bench process (torch.distributed over NCCL when launched by torchrun)."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        import torch
        self.torch = torch
        self.dist = None
        self.ctl = None
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            # NCCL is the job's process group (created on first use: the sharded block's QNB_COMM=nccl variant and the
            # object broadcasts); barriers and the max over ranks of the timed windows go through a gloo group on the host, so
            # that no NCCL communicator -- with its proxy and watchdog threads polling the driver -- exists while the
            # per-rank loops are timed (r03i: 0.058 ms per step per rank at N=2 and N=8 against 0.046 at N=1)
            dist.init_process_group("nccl")
            if os.environ.get("QNB_BENCH_CONTROL", "gloo") == "gloo":
                try:
                    self.ctl = dist.new_group(backend="gloo")
                except Exception:      # no usable host interface for gloo: the NCCL group serves the barriers as well
                    self.ctl = None
            self.dist = dist
        self.dev = self.local_rank if self.world > 1 else 0
        # one slice of the host cores per rank (all GPUs of the box hang off the same NUMA node, `nvidia-smi topo -m`): eight
        # Python ranks and the clock sampler otherwise migrate over each other's cores (r01: e2e efficiency 0.88 at N=8)
        self.cores = None
        if self.world > 1 and hasattr(os, "sched_setaffinity"):
            try:
                avail = sorted(os.sched_getaffinity(0))
                per = max(1, len(avail) // self.world)
                mine = avail[self.local_rank * per:(self.local_rank + 1) * per] or avail
                os.sched_setaffinity(0, mine)
                self.cores = [mine[0], mine[-1]]
            except OSError:
                pass

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            if self.ctl is not None:
                self.dist.barrier(group=self.ctl)
            else:
                self.dist.barrier()
            self.torch.cuda.synchronize()

    def reduce(self, v, op="max"):
        if self.dist is None:
            return v
        rop = {"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM}[op]
        if self.ctl is not None:
            t = self.torch.tensor([v], dtype=self.torch.float64)
            self.dist.all_reduce(t, op=rop, group=self.ctl)
        else:
            t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
            self.dist.all_reduce(t, op=rop)
        return float(t.item())

    def gather_objects(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.ctl)
        return out

    def close(self):
        if self.dist is not None:
            self.barrier()
            self.dist.destroy_process_group()


def timed_md(job, g, lam, steps, repeats):
    """`repeats` windows of exactly `steps` device-resident steps (list rebuild every NBCYCLE), each bracketed by a barrier
    and a synchronize; per window the max over ranks.  ms per window."""
    out = []
    for _ in range(repeats):
        job.barrier()
        ms = g.bench_md(lam, steps, NBCYCLE)
        job.barrier()
        out.append(job.reduce(ms))
    return out


def timed_e2e(job, g, q, x, lam, cuts, steps, repeats, phase=None):
    """`repeats` windows of `steps` end-to-end steps; the MD step counter (list build when it hits a multiple of NBCYCLE) runs
    on across the windows, as md_run's does."""
    phase = phase if phase is not None else [0]
    d = np.zeros((q.natom, 3))
    # what the Fortran host does once at set-up for its module arrays x and d (qnb_register_host_buffers)
    g.release_host_buffers()
    g.register_host_buffers(x.reshape(-1), d.reshape(-1))
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    step, _, _ = g.bind_step(x.reshape(-1), lam, d.reshape(-1), d_is_zero=True)   # pointers held once, as a compiled host does
    out = []
    for _ in range(repeats):
        job.barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            if phase[0] % NBCYCLE == 0:
                g.make_pair_lists(x, **cuts, counts=False)
            phase[0] += 1
            d[:] = 0                                           # d(:) = zero, potene.f90:109
            step()
        job.barrier()
        out.append(job.reduce(time.perf_counter() - t0))
    return out


def batched_windows(job, q, cuts, lam, nwin, steps, warmup, npairs, flop_step, fp32_peak):
    """nwin independent windows of the workload on THIS GPU through qnb_build_lists_batch / qnb_nonbond_batch (host
    buffers, own coordinates per window, list build every NBCYCLE steps): aggregate end-to-end throughput."""
    from q6_b200.engine import Qnb, QnbBatch
    hs = [Qnb(q, device=job.dev) for _ in range(nwin)]
    b = QnbBatch(hs)
    rng = np.random.default_rng(11 + job.rank)
    xs = [q.xtop + rng.normal(0.0, 0.01, q.xtop.shape) for _ in range(nwin)]
    lams = [lam for _ in range(nwin)]

    tb = [0.0, 0]

    phase = [0]

    def run(n):
        for k in range(n):
            if phase[0] % NBCYCLE == 0:
                t0 = time.perf_counter()
                b.make_pair_lists(xs, **cuts)
                tb[0] += time.perf_counter() - t0
                tb[1] += 1
            phase[0] += 1
            b.pot_energy_nonbonds(xs if k == 0 else None, lams if k == 0 else None)

    run(max(warmup, NBCYCLE + 1))
    tb[0], tb[1] = 0.0, 0
    times = []
    for _ in range(5):
        job.barrier()
        t0 = time.perf_counter()
        run(steps)
        job.barrier()
        times.append(job.reduce(time.perf_counter() - t0))
    host = b.last_timing()
    for g in hs:
        g.close()
    t = float(np.mean(times)) / steps            # seconds per batched step (nwin windows advance one step each)
    agg = job.world * nwin * npairs / t
    return {"windows_per_gpu": nwin, "ms_per_batched_step": t * 1e3, "ms_per_window_step": t / nwin * 1e3,
            "value": agg, "unit": "pairs/s", "steps_per_s": job.world * nwin / t,
            "fep_windows_per_hour": 3600.0 * job.world * nwin / (STEPS_PER_WINDOW * t),
            "roofline_frac_fp32_e2e": nwin * flop_step / t / 1e12 / fp32_peak,
            "spread_s": spread(times), "list_build_ms_per_batched_call": tb[0] / max(tb[1], 1) * 1e3,
            "host_breakdown_last_call_us": {k: round(v * 1e6, 1) for k, v in host.items()},
            "note": "qnb_build_lists_batch + qnb_nonbond_batch: one C-ABI call per step for all windows of the GPU, host "
                    f"buffers, own coordinates per window, list build every {NBCYCLE} steps; roofline_frac = algorithmic "
                    "flop of all step kernels of all windows / wall time / FP32 peak"}


def fep_farm(job, steps_per_window, batch):
    """BASELINE config 3: FEP_WINDOWS two-state lambda windows (C3) scheduled over the GPUs of the job, window w on rank
    w mod N, each rank advancing its windows `batch` at a time through the batched calls.  Wall clock from the first step
    to the last window finished; windows/hour scaled to the STEPS_PER_WINDOW steps of a production window."""
    from q6_b200 import synth
    from q6_b200.engine import Qnb, QnbBatch
    q, cuts, _ = synth.config("C3")
    mine = [w for w in range(FEP_WINDOWS) if w % job.world == job.rank]
    nb = min(batch, max(len(mine), 1))
    hs = [Qnb(q, device=job.dev) for _ in range(nb)]
    rng = np.random.default_rng(100 + job.rank)

    batches = {}

    def advance(ws, nsteps):
        if len(ws) not in batches:
            batches[len(ws)] = QnbBatch(hs[:len(ws)])
        b = batches[len(ws)]
        xs = [q.xtop + rng.normal(0.0, 0.01, q.xtop.shape) for _ in ws]
        lams = [np.array([1.0 - 0.02 * w, 0.02 * w]) for w in ws]
        for k in range(nsteps):
            if k % NBCYCLE == 0:
                b.make_pair_lists(xs, **cuts)
            b.pot_energy_nonbonds(xs if k == 0 else None, lams if k == 0 else None)

    groups = [mine[i:i + nb] for i in range(0, len(mine), nb)]
    for size in sorted({len(gr) for gr in groups}):
        advance(mine[:size], NBCYCLE + 3)   # warm-up of every batch size that will occur: allocations, page-locking, graphs
    job.barrier()
    t0 = time.perf_counter()
    for gr in groups:
        advance(gr, steps_per_window)
    job.torch.cuda.synchronize()
    t_rank = time.perf_counter() - t0
    job.barrier()
    wall = job.reduce(t_rank)
    for g in hs:
        g.close()
    return {"workload": WORKLOADS["C3"], "windows": FEP_WINDOWS, "windows_per_rank_max": -(-FEP_WINDOWS // job.world),
            "batch": nb, "steps_per_window_timed": steps_per_window, "wall_s": wall,
            "windows_per_hour": FEP_WINDOWS * 3600.0 / (wall * STEPS_PER_WINDOW / steps_per_window),
            "ms_per_window_step": wall / (-(-FEP_WINDOWS // job.world) * steps_per_window) * 1e3,
            "note": f"{FEP_WINDOWS} windows (lambda 1.00 -> 0.00), window w on rank w mod N, {nb} windows per batched call, "
                    f"host buffers, list build every {NBCYCLE} steps; wall clock of the slowest rank, scaled from "
                    f"{steps_per_window} to {STEPS_PER_WINDOW} steps per window"}


def sharded_c5(job, steps, warmup, repeats):
    """BASELINE config 5 at N > 1: ONE periodic 98k-atom system, the rows of its pair lists sharded over the N GPUs
    (calculation_assignment ranges, distribute_nonbonds nonbondene.f90:80-505), NCCL all-reduce of [d|E|EQ] per step
    (gather_nonbond L6760, potene.f90:195-222) and of the LRF moments per list build (lrf_gather L616).  The N = 1 time is
    measured in the same run (every rank runs the whole system on its own GPU, max over ranks)."""
    from q6_b200 import synth
    from q6_b200.engine import Qnb
    from q6_b200.system import shard_system
    q, cuts, lam = synth.config("C5")
    x = q.xtop.copy()
    # ---- N = 1 reference, same run, same GPUs
    g1 = Qnb(q, device=job.dev)
    c1 = g1.make_pair_lists(x, **cuts)
    npairs = pairs_per_step(c1, q.nstates, 0)
    g1.bench_md(lam, max(warmup, NBCYCLE), NBCYCLE)
    t1_md = float(np.mean(timed_md(job, g1, lam, steps, repeats))) / steps
    job.barrier()
    t1_step = job.reduce(g1.bench_nonbond(lam, 100) / 100)
    t1_build = job.reduce(g1.bench_build_lists(3) / 3)
    g1.close()
    # ---- sharded
    g = Qnb(shard_system(q, job.rank, job.world), device=job.dev)
    comm = g.comm_connect(job.rank, job.world, job.dist, os.environ.get("QNB_COMM", "auto"))
    cl = g.make_pair_lists(x, **cuts)
    tot = job.reduce(float(cl[2]), "sum")
    g.bench_md(lam, max(warmup, NBCYCLE), NBCYCLE)
    md = timed_md(job, g, lam, steps, repeats)
    tN_md = float(np.mean(md)) / steps
    job.barrier()
    tN_step = job.reduce(g.bench_nonbond(lam, 100) / 100)
    job.barrier()
    tN_build = job.reduce(g.bench_build_lists(3) / 3)
    job.barrier()
    ar = job.reduce(g.bench_allreduce(50))
    e2e = float(np.mean(timed_e2e(job, g, q, x, lam, cuts, steps, 3))) / steps
    kt = g.bench_kernels(lam, 10, flush_l2=True)
    g.comm_status()
    g.close()
    return {"workload": WORKLOADS["C5"], "natom": int(q.natom), "partition": "pairs" if os.environ.get("QNB_SHARD_PAIRS") else "rows",
            "pairs_per_step": npairs, "list_entries_all_ranks": int(tot),
            "ms_per_step": tN_md, "value": npairs / (tN_md * 1e-3), "unit": "pairs/s", "device_step_ms": tN_step,
            "list_build_ms": tN_build, "comm": comm, "allreduce_ms": ar, "allreduce_bytes": int((3 * q.natom + 7 + 6 * q.nstates) * 8),
            "e2e": {"ms_per_step": e2e * 1e3, "value": npairs / e2e, "unit": "pairs/s"},
            "n1": {"ms_per_step": t1_md, "device_step_ms": t1_step, "list_build_ms": t1_build},
            "efficiency_vs_n1": t1_md / (job.world * tN_md), "efficiency_device_step_vs_n1": t1_step / (job.world * tN_step),
            "speedup_vs_n1": t1_md / tN_md, "kernels_ms_rank0_l2_flushed": kt, "spread_ms_per_window": spread(md),
            "scaling": "strong",
            "note": "ms_per_step = device-resident MD loop (one evaluation + all-reduce per step, list build + LRF all-reduce "
                    f"every {NBCYCLE} steps), CUDA events, max over ranks, mean of {repeats} windows; n1 = the unsharded "
                    "system on one GPU in the same run; comm p2p = one kernel per rank over CUDA-IPC peer memory (owner of a "
                    "slice sums it over the ranks and writes the sum into every rank's buffer), captured in the step's CUDA "
                    "graph; comm nccl = ncclAllReduce issued eagerly after the step's kernels (QNB_COMM=nccl)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=25)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=list(WORKLOADS))
    ap.add_argument("--repeats", type=int, default=25, help="timed windows of --steps steps each (mean reported, min / median / max under `repeats`)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sharded", action="store_true",
                    help="N>1: ONE system as the HEADLINE, rows of the pair lists sharded over the GPUs, NCCL all-reduce of forces "
                         "and energies (strong scaling). Default for --workload C5; otherwise one independent window per GPU")
    ap.add_argument("--batch-windows", type=int, default=7, help="windows per GPU of the batched calls (0: skip); 7 = ceil(51/8)")
    ap.add_argument("--farm-steps", type=int, default=100, help="steps per window timed in the 51-window farm (0: skip)")
    ap.add_argument("--no-sharded-c5", action="store_true", help="N>1: skip the sharded C5 block")
    args = ap.parse_args()
    steps, warmup, repeats = args.steps, max(args.warmup, 3), max(args.repeats, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from q6_b200 import synth
    q, cuts, lam = synth.config(args.workload)
    config = {"workload": WORKLOADS[args.workload], "natom": int(q.natom), "nat_solute": int(q.nat_solute), "nwat": int(q.nwat),
              "nqat": int(q.nqat), "nstates": int(q.nstates), "nbcycle": NBCYCLE,
              "replicas": "one independent system (lambda window) per GPU, each rank pinned to its own slice of the host cores"
                          if world > 1 else "single system",
              "cache": "coordinates, rows and gradient (< 3 MB) are L2-resident by nature of the workload; "
                       "per-kernel times in roofline are taken with an L2 flush before every launch"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        from oracle.pyoracle import Oracle
        o = Oracle(q)
        counts = o.make_pair_lists(q.xtop, **cuts)
        nqq = sum(o.list_count(5, s + 1) + o.list_count(6, s + 1) for s in range(q.nstates))
        npairs = pairs_per_step(counts, q.nstates, nqq)
        o.close()
        nrep = max(1, args.gpus) if not (args.sharded or args.workload == "C5") else 1
        t_step, nsteps, per = cpu_reference_replicas(q, cuts, lam, threads, nrep, target_seconds=max(5.0, min(30.0, 0.1 * steps)))
        val = nrep * npairs / t_step
        config["replicas"] = (f"{nrep} independent systems advanced concurrently on the host, {per} threads each "
                              "(the GPU arm advances one system per GPU)") if nrep > 1 else "single system"
        line = {"impl": "reference", "metric": "nonbonded pair interactions/s", "value": val, "unit": "pairs/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t_step * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port",
                                 "sample": f"{nsteps} nonbonded evaluations + 1 list build (amortised over {NBCYCLE} steps) "
                                           f"of {nrep} system(s) on {threads} threads in all ({per} per system), i-range "
                                           "decomposition as Qdyn6p"},
                "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "per_system_value": npairs / t_step, "systems": nrep,
                "ns_per_day": 86400.0 / t_step * DT_FS * 1e-6, "pairs_per_step": npairs}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    job = Job()
    from q6_b200.engine import Qnb, bench_peak
    dev = job.dev
    if world > 1 and q.nstates > 1:
        # independent lambda windows: rank r runs window r of 51 (gen_inps.pl: 1.00 -> 0.00 step 0.02)
        l1 = 1.0 - 0.02 * (rank % FEP_WINDOWS)
        lam = np.array([l1, 1.0 - l1])
    sharded = world > 1 and (args.sharded or args.workload == "C5")
    if sharded:
        # distribute_nonbonds (nonbondene.f90:80-505): contiguous i-ranges per rank, everything else replicated
        from q6_b200.system import shard_system
        g = Qnb(shard_system(q, rank, world), device=dev)
        comm = g.comm_connect(rank, world, job.dist, os.environ.get("QNB_COMM", "auto"))
        config["replicas"] = f"one system, rows of the pair lists sharded over {world} GPUs, all-reduce of [d|E|EQ] per step ({comm})"
    else:
        g = Qnb(q, device=dev)
    x = q.xtop.copy()
    counts = g.make_pair_lists(x, **cuts)
    nqq = sum(g.list_count(5, s + 1) + g.list_count(6, s + 1) for s in range(q.nstates))
    counts_local, nqq_local = np.array(counts).copy(), nqq     # what THIS rank's kernels process (roofline)
    if sharded:
        t = job.torch.tensor([float(c) for c in counts[:5]] + [float(nqq)], dtype=job.torch.float64, device="cuda")
        job.dist.all_reduce(t)
        counts = np.array([int(v) for v in t[:5].tolist()] + [0] * (len(counts) - 5))
        nqq = int(t[5].item())
    npairs = pairs_per_step(counts, q.nstates, nqq)

    # ---- device-resident throughput: `repeats` windows of exactly `steps` steps
    g.bench_md(lam, warmup, NBCYCLE)
    clk_path = os.path.join(ROOT, f".bench_clocks_{rank}.csv")
    proc, fh = clocks_sampler(clk_path) if rank == 0 else (None, None)
    l0 = g.launch_count()
    md = timed_md(job, g, lam, steps, repeats)
    launches = (g.launch_count() - l0) // repeats
    # mean over the windows: the list build falls into 4 of 5 windows of 20 steps, the mean weighs it as the MD loop does
    t_step = float(np.mean(md)) * 1e-3 / steps
    # ---- end to end through the C ABI with host buffers
    e2e_phase = [0]
    timed_e2e(job, g, q, x, lam, cuts, max(warmup, NBCYCLE + 1), 1, e2e_phase)
    e2e_runs = timed_e2e(job, g, q, x, lam, cuts, steps, 10, e2e_phase)
    e2e_step = float(np.mean(e2e_runs)) / steps
    host_breakdown = g.last_timing()
    h2d, d2h = g.last_copy_bytes()
    h2d_step = h2d + (3 * q.natom * 8) / NBCYCLE   # + the list build's coordinate upload, amortised
    if proc is not None:
        proc.terminate()
        fh.close()
    clocks = clocks_summary(clk_path, set(range(world)) if world > 1 else {dev}) if rank == 0 else None
    try:
        os.remove(clk_path)
    except OSError:
        pass

    # ---- per-kernel times (L2 flushed before each launch), roofline of the dominant kernel, of the step and of the build
    list_ms_all = g.bench_build_lists(3) / 3 if sharded else None   # the sharded build all-reduces the LRF moments
    roof, ext = None, {}
    fp32_meas = bench_peak(0, dev)
    fp64_meas = bench_peak(1, dev)
    alg = step_flop(counts_local, q.nstates, nqq_local, q.natom)
    flop_step = float(sum(alg.values()))
    if rank == 0:
        kt = g.bench_kernels(lam, 20, flush_l2=True)
        kt_warm = g.bench_kernels(lam, 20, flush_l2=False)
        list_ms = list_ms_all if sharded else g.bench_build_lists(5) / 5
        bt = g.last_build_timing()
        device_step_ms = None if sharded else g.bench_nonbond(lam, 200) / 200
        peaks = measured_peaks()
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic, traffic_src = tj["bytes_per_launch"], tj.get("source")
        except Exception:
            traffic, traffic_src = {}, None
        dom = max((k for k in kt if alg.get(k, 0.0) > 0.0), key=lambda k: kt[k])
        ach = alg[dom] / (kt[dom] * 1e-3) / 1e12
        nlrf = lrf_interactions(q, cuts) if q.use_LRF else None
        # the pipe a kernel computes on: the row and Q-partner gradient kernels are FP32, the Q-atom / static-list / LRF Taylor
        # kernels FP64 throughout (FEP observables)
        fp64_kernels = {"k_q_atom", "k_qq_static", "k_lrf_taylor"}
        dom_peak = fp64_meas if dom in fp64_kernels else fp32_meas
        roof = {"bound": "fp64" if dom in fp64_kernels else "fp32", "kernel": dom, "achieved": ach, "peak": dom_peak, "unit": "TFLOP/s",
                "frac": ach / dom_peak,
                "dominant_by": "largest CUDA-event time of one launch with the L2 flushed; at 12 k atoms every step kernel is a "
                               "single latency-bound wave of 8-17 us, see roofline.step for the whole evaluation and the MD loop",
                "traffic": traffic.get(dom),
                "traffic_source": f"archived: ncu --set full capture {traffic_src} (dram__bytes_read.sum + dram__bytes_write.sum per "
                                  "launch), not measured in this run",
                "peak_source": "FP32 FMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP32 figure); "
                               f"nominal 148 SM x 128 x 2 x {peaks.get('sm_max_mhz', 1965.0)} MHz = "
                               f"{148 * 128 * 2 * peaks.get('sm_max_mhz', 1965.0) * 1e-6:.1f} TFLOP/s; FP64 pipe measured {fp64_meas:.1f}",
                "algorithmic_flop_per_launch": alg[dom], "ms_per_launch_l2_flushed": kt[dom],
                "ms_per_launch_warm": kt_warm.get(dom),
                "kernels_ms_l2_flushed": kt, "kernels_ms_warm": kt_warm,
                "kernels_frac_fp32": {k: alg[k] / (kt[k] * 1e-3) / 1e12 / fp32_meas for k in kt if alg.get(k, 0.0) > 0.0},
                # the whole evaluation (its kernels overlap on several streams inside one CUDA graph) and the MD loop
                "step": {"algorithmic_flop": flop_step, "device_step_ms": device_step_ms,
                         "frac_fp32_device_step": (flop_step / (device_step_ms * 1e-3) / 1e12 / fp32_meas) if device_step_ms else None,
                         "md_loop_ms_per_step": t_step * 1e3,
                         "lrf_update_flop_per_step": (nlrf * FLOP_LRF_UPDATE / NBCYCLE) if nlrf else None,
                         "frac_fp32_md_loop": (flop_step + (nlrf * FLOP_LRF_UPDATE / NBCYCLE if nlrf else 0.0)) / t_step / 1e12 / fp32_meas},
                "lrf": {"kernel": "k_lrf_allpairs" if not q.use_PBC else "k_lrf_accumulate", "interactions": nlrf,
                        "flop": (nlrf * FLOP_LRF_UPDATE) if nlrf else None, "ms": bt["lrf_ms"],
                        "frac_fp32": (nlrf * FLOP_LRF_UPDATE / (bt["lrf_ms"] * 1e-3) / 1e12 / fp32_meas) if nlrf and bt["lrf_ms"] > 0 else None,
                        "frac_fp64": (nlrf * FLOP_LRF_UPDATE / (bt["lrf_ms"] * 1e-3) / 1e12 / fp64_meas) if nlrf and bt["lrf_ms"] > 0 else None,
                        "note": "lrf_update per list build (phi0-phi2 FP64, phi3 FP32); interactions counted with a KD-tree over "
                                "the switch atoms; CUDA events on the side stream the kernels run on, next to the row scan"},
                "rows": {"kernel": "k_rows_scan (count pass + fill pass)", "ms": bt["rows_count_ms"] + bt["rows_fill_ms"],
                         "ms_count": bt["rows_count_ms"], "ms_fill": bt["rows_fill_ms"],
                         "algorithmic_bytes": 16.0 * 2 * float(g_scanned_candidates(q, cuts)) + 4.0 * float(g_total_row_entries(g)),
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
                "list_build": {"ms": list_ms, "hbm_peak_gbs": peaks.get("hbm_gbs"),
                               "algorithmic_bytes": 24.0 * q.ncgp + 8.0 * float(g_total_unit_pairs(g, q)) + 320.0 * q.ncgp}}
        for blk in (roof["rows"], roof["list_build"]):
            blk["achieved_gbs"] = blk["algorithmic_bytes"] / max(blk["ms"] * 1e-3, 1e-12) / 1e9
            if blk["hbm_peak_gbs"]:
                blk["frac_hbm"] = blk["achieved_gbs"] / blk["hbm_peak_gbs"]
        roof["list_build"]["frac"] = roof["list_build"].get("frac_hbm")
        # ---- optional paths, reported next to the headline (never part of it)
        if not sharded:
            ext = {"device_step_ms": device_step_ms,
                   "device_step_ms_no_pp_pw_ww_energies": g.bench_nonbond(lam, 200, energies=False) / 200}
            if not q.use_PBC and q.nwat > 0:
                from q6_b200.engine import wat_shells
                rw = float(np.linalg.norm(q.xtop[q.nat_solute:] - np.asarray(q.xpcent), axis=1).max())
                g.set_solvent_restraints(wat_shells(q.xpcent, rw, crgQtot=-1.0), np.zeros(8))
                ext["device_step_ms_with_solvent_restraints"] = g.bench_nonbond(lam, 200, restraints=True) / 200
            if world == 1 and q.nwat > 0:
                try:     # N2: solvent SHAKE on the device (qnb_shake); reported, never fatal
                    cons, starts, winv = synth.water_constraints(q)
                    g.set_constraints(cons, starts, winv)
                    g.pot_energy_nonbonds(x, lam)
                    moved = x + np.random.default_rng(3).normal(0, 0.01, x.shape)
                    _, sweeps = g.shake(moved)
                    t0 = time.perf_counter()
                    for _ in range(50):
                        g.shake(moved)
                    ext["solvent_shake"] = {"ms_per_call_host_buffers": (time.perf_counter() - t0) / 50 * 1e3,
                                            "molecules": int(q.nwat), "sweeps_per_molecule": sweeps / max(int(q.nwat), 1)}
                except Exception as e:
                    ext["solvent_shake"] = {"error": str(e)[-200:]}
    g.close()

    # ---- several windows per GPU through the batched calls; the 51-window farm; the sharded 98k-atom system (all ranks)
    bw = farm = sh = None
    if not sharded and args.batch_windows > 1:
        bw = batched_windows(job, q, cuts, lam, args.batch_windows, steps, warmup, npairs, flop_step, fp32_meas)
    if not sharded and args.farm_steps > 0:
        farm = fep_farm(job, args.farm_steps, max(args.batch_windows, 1))
    if world > 1 and not sharded and not args.no_sharded_c5:
        try:
            sh = sharded_c5(job, 50, 25, 7)
        except Exception as e:   # reported, never fatal for the headline
            sh = {"error": str(e)[-300:]}

    if rank == 0:
        nsys = 1 if sharded else world          # systems advanced per step over the whole job
        value = npairs * nsys / t_step
        line = {"metric": "nonbonded pair interactions/s", "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
                "scaling": "strong" if sharded else "weak", "vs_baseline": None,
                "dtype": "f32 pair math, f64 accumulation and energies",
                "data": "synthetic", "config": config, "pairs_per_step": npairs,
                "repeats": {"windows": repeats, "steps_per_window": steps, "ms_per_window": spread(md),
                            "note": "ms_per_step = mean window / steps (the MD step counter runs on across the windows, so the list "
                                    f"build every {NBCYCLE} steps falls into some windows and not others: min / median / max show it); every "
                                    "window is bracketed by barrier + synchronize and reduced with max over ranks"},
                "ns_per_day": 86400.0 / e2e_step * DT_FS * 1e-6,
                "ns_per_day_device_resident": 86400.0 / t_step * DT_FS * 1e-6,
                "fep_windows_per_hour": nsys * 3600.0 / (STEPS_PER_WINDOW * e2e_step),
                "e2e": {"value": npairs * nsys / e2e_step, "unit": "pairs/s", "ms_per_step": e2e_step * 1e3,
                        "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h),
                        "spread_s_per_window": spread(e2e_runs),
                        "host_breakdown_last_call_us": {k: round(v * 1e6, 1) for k, v in host_breakdown.items()}},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
        if bw is not None:
            bw["vs_single_window_e2e"] = bw["value"] / line["e2e"]["value"]
            line["batched_windows"] = bw
        if farm is not None:
            line["fep_farm"] = farm
        if sh is not None:
            line["sharded_c5"] = sh
        if world == 1:
            try:
                ms_c = compiled_host_e2e(q, cuts, lam, max(steps, 10 * NBCYCLE), dev)
                line["e2e_compiled_host"] = {"ms_per_step": ms_c, "value": npairs / (ms_c * 1e-3), "unit": "pairs/s",
                                             "note": "q6_b200/host/qdyn_nb (C++ host over the C ABI, input files written by "
                                                     "synth.write_files, host buffers, list build every "
                                                     f"{NBCYCLE} steps); separate process, timed by its own host clock"}
            except Exception as e:  # reported, never fatal: the headline e2e above does not depend on it
                line["e2e_compiled_host"] = {"error": str(e)[-200:]}
        line["extensions"] = ext
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            t_cpu, nsteps, _ = cpu_reference(q, cuts, lam, threads, target_seconds=12.0)
            t_cpu1, nsteps1, _ = cpu_reference(q, cuts, lam, 1, target_seconds=8.0)
            line["cpu_baseline"] = {"value": npairs / t_cpu, "unit": "pairs/s", "cores": threads, "kind": "port",
                                    "sample": f"{nsteps} nonbonded evaluations + 1 list build (amortised over {NBCYCLE} "
                                              f"steps) of the same system on {threads} threads (Qdyn6p-style i-range "
                                              "decomposition of the C oracle)",
                                    "serial_value": npairs / t_cpu1, "serial_ms_per_step": t_cpu1 * 1e3,
                                    "ms_per_step": t_cpu * 1e3}
        print(json.dumps(line))
    job.close()


def g_total_unit_pairs(g, q):
    """Charge-group pairs emitted by the list build (each listed pair once)."""
    ww = g.list_count(2) // 9
    # pp/pw group pairs are not exported separately; atom pairs / mean group size is a lower bound that the
    # roofline of the (tiny) byte figure does not depend on in any visible digit
    return ww + g.list_count(1) // 9 + g.list_count(0) // 9


def g_total_row_entries(g):
    """Row entries written by the fill pass (both sides of every pair)."""
    return 2 * (g.list_count(2) // 9) + 2 * g.list_count(1) // 3 + 2 * g.list_count(0)


def g_scanned_candidates(q, cuts):
    """Candidates the row scan reads (16-byte screening record each, both passes): units in the 5 x 5 x 5 cells of half
    the cut-off around every unit ~ (5/2 Rc)^3 / (4/3 pi Rc^3) = 3.73 x the listed neighbours; estimated from the density."""
    cgp = np.asarray(q.cgp).reshape(-1, 3)
    x = np.asarray(q.xtop).reshape(-1, 3)[cgp[:, 0] - 1]
    rc = float(np.sqrt(max(cuts["Rcpp2"], cuts["Rcpw2"], cuts["Rcww2"])))
    if q.use_PBC:
        vol = float(np.prod(np.asarray(q.boxlength, float)))
    else:
        r = np.linalg.norm(x - x.mean(axis=0), axis=1).max()
        vol = 4.0 / 3.0 * np.pi * r ** 3
    return len(x) * min(len(x), len(x) / vol * (2.5 * rc) ** 3)


if __name__ == "__main__":
    main()
