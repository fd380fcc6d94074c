#!/usr/bin/env python
"""Benchmark of the Qdyn6 nonbonded hot path (BASELINE.json metric) -- see DESIGN.md "Measurement".

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2|C3|C4s|C5]

A "step" is one MD step's share of the hot path on one synthetic system: one nonbonded evaluation
(qnb_nonbond) plus a pair-list rebuild every NBcycle = 25 steps, as md_run does (md.f90:1661,1742).
`value`  = pair interactions/s with coordinates already resident in HBM (CUDA events, max over ranks);
`e2e`    = the same through the C ABI with HOST buffers (x up, d + energies down inside the timed region).
N > 1    = one independent replica / lambda window per GPU, no data-path collective (weak scaling).
--impl reference times the CPU oracle (the reference's algorithm, decomposed over all host threads like
Qdyn6p) on the same workload; the Fortran reference itself cannot be built here (no Fortran compiler).
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
# algorithmic flop per unit (SURVEY.md 8d / BASELINE.md 3)
FLOP_PAIR, FLOP_WW_MOL, FLOP_LRF_UPDATE, FLOP_LRF_TAYLOR = 33.0, 209.0, 80.0, 125.0


def flop_q_pair(nstates):
    return 21.0 + 20.0 * nstates


def pairs_per_step(counts, nstates, nqq_total):
    """Sum of list entries evaluated per step: nbpp+nbpw+nbww+(nbqp+nbqw)*nstates+nbqq+nbqqp."""
    return int(counts[0] + counts[1] + counts[2] + (counts[3] + counts[4]) * nstates + nqq_total)


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=25)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C3", "C4s", "C5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sharded", action="store_true",
                    help="N>1: ONE system, i-ranges of the pair lists sharded over the GPUs, NCCL all-reduce of forces and "
                         "energies (strong scaling). Default for --workload C5; otherwise one independent window per GPU")
    ap.add_argument("--windows-per-gpu", type=int, default=4,
                    help="also time this many independent windows sharing each GPU (0/1: skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = args.steps, max(args.warmup, 3)

    from q6_b200 import synth
    q, cuts, lam = synth.config(args.workload)
    workload = {"C2": "C2: 32 A TIP3P sphere, protein-like core + 46 Q-atoms, LRF, 10 A cut-offs, 1 state",
                "C3": "C3: 25 A water sphere, 46-Q-atom ligand annihilation FEP, 2 states",
                "C4s": "C4s: 25 A sphere, 60 Q-atoms, two-state EVB", "C5": "C5: 32768-water periodic box"}[args.workload]
    config = {"workload": workload, "natom": int(q.natom), "nat_solute": int(q.nat_solute), "nwat": int(q.nwat),
              "nqat": int(q.nqat), "nstates": int(q.nstates), "nbcycle": NBCYCLE,
              "replicas": "one independent system (lambda window) per GPU" if world > 1 else "single system",
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
        t_step, nsteps, _ = cpu_reference(q, cuts, lam, threads, target_seconds=max(5.0, min(30.0, 0.1 * steps)))
        val = npairs / t_step
        line = {"impl": "reference", "metric": "nonbonded pair interactions/s", "value": val, "unit": "pairs/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t_step * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port",
                                 "sample": f"{nsteps} nonbonded evaluations + 1 list build (amortised over {NBCYCLE} steps) "
                                           f"on {threads} threads, i-range decomposition as Qdyn6p"},
                "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "ns_per_day": 86400.0 / t_step * DT_FS * 1e-6, "pairs_per_step": npairs}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from q6_b200.engine import Qnb, bench_peak
    dev = local_rank if world > 1 else 0
    if world > 1 and q.nstates > 1:
        # independent lambda windows: rank r runs window r of 51 (gen_inps.pl: 1.00 -> 0.00 step 0.02)
        l1 = 1.0 - 0.02 * (rank % 51)
        lam = np.array([l1, 1.0 - l1])
    sharded = world > 1 and (args.sharded or args.workload == "C5")
    if sharded:
        # distribute_nonbonds (nonbondene.f90:80-505): contiguous i-ranges per rank, everything else replicated
        from q6_b200.system import shard_system
        g = Qnb(shard_system(q, rank, world), device=dev)
        uid = [g.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(rank, world, uid[0])
        config["replicas"] = f"one system, pair lists sharded over {world} GPUs, NCCL all-reduce of [d|E|EQ] per step"
        args.windows_per_gpu = 0
    else:
        g = Qnb(q, device=dev)
    x = q.xtop.copy()
    counts = g.make_pair_lists(x, **cuts)
    nqq = sum(g.list_count(5, s + 1) + g.list_count(6, s + 1) for s in range(q.nstates))
    counts_local, nqq_local = np.array(counts).copy(), nqq     # what THIS rank's kernels process (roofline)
    if sharded:
        t = torch.tensor([float(c) for c in counts[:5]] + [float(nqq)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        counts = np.array([int(v) for v in t[:5].tolist()] + [0] * (len(counts) - 5))
        nqq = int(t[5].item())
    npairs = pairs_per_step(counts, q.nstates, nqq)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput
    g.bench_md(lam, warmup, NBCYCLE)
    clk_path = os.path.join(ROOT, f".bench_clocks_{rank}.csv")
    proc, fh = clocks_sampler(clk_path) if rank == 0 else (None, None)
    l0 = g.launch_count()
    barrier()
    ms = g.bench_md(lam, steps, NBCYCLE)
    barrier()
    launches = g.launch_count() - l0
    ms = max_over_ranks(ms)
    t_step = ms * 1e-3 / steps
    # ---- end to end through the C ABI with host buffers
    d = np.zeros((q.natom, 3))
    for k in range(warmup):
        if k % NBCYCLE == 0:
            g.make_pair_lists(x, **cuts, counts=False)
        d[:] = 0
        g.pot_energy_nonbonds(x, lam, d=d)
    barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        if k % NBCYCLE == 0:
            g.make_pair_lists(x, **cuts, counts=False)
        d[:] = 0
        _, E, EQ = g.pot_energy_nonbonds(x, lam, d=d)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    host_breakdown = g.last_timing()
    h2d, d2h = g.last_copy_bytes()
    # ---- k independent windows per GPU (FEP production: the windows do not talk to each other; one host thread and
    # one handle each, same C-ABI calls with host buffers).  Reported next to the single-window e2e, not instead of it.
    kwin = args.windows_per_gpu
    win_s = None
    if kwin > 1:
        import threading
        extra = [Qnb(q, device=dev) for _ in range(kwin - 1)]
        handles = [g] + extra

        def window(hd, nst):
            xw = q.xtop.copy()
            dw = np.zeros((q.natom, 3))
            for k in range(nst):
                if k % NBCYCLE == 0:
                    hd.make_pair_lists(xw, **cuts, counts=False)
                dw[:] = 0
                hd.pot_energy_nonbonds(xw, lam, d=dw)

        for hd in handles:
            window(hd, NBCYCLE + 3)
        th = [threading.Thread(target=window, args=(hd, steps)) for hd in handles]
        barrier()
        t0 = time.perf_counter()
        for a in th:
            a.start()
        for a in th:
            a.join()
        barrier()
        win_s = max_over_ranks(time.perf_counter() - t0)
        for hd in extra:
            hd.close()
    h2d_step = h2d + (3 * q.natom * 8) / NBCYCLE   # + the list build's coordinate upload, amortised
    if proc is not None:
        proc.terminate()
        fh.close()
    clocks = clocks_summary(clk_path, set(range(world)) if world > 1 else {dev}) if rank == 0 else None
    try:
        os.remove(clk_path)
    except OSError:
        pass

    # ---- per-kernel times (L2 flushed before each launch) and the roofline of the dominant kernel
    line = None
    list_ms_all = g.bench_build_lists(3) / 3 if sharded else None   # the sharded build all-reduces the LRF moments
    if rank == 0:
        kt = g.bench_kernels(lam, 20, flush_l2=True)
        kt_warm = g.bench_kernels(lam, 20, flush_l2=False)
        list_ms = list_ms_all if sharded else g.bench_build_lists(3) / 3
        counts, nqq = counts_local, nqq_local
        peaks = measured_peaks()
        fp32_meas = bench_peak(0, dev)
        fp64_meas = bench_peak(1, dev)
        ns_q = flop_q_pair(q.nstates)
        # algorithmic work per launch of each kernel: every listed pair once; a pw pair is split half/half between
        # the two kernels that each accumulate one side of it (DESIGN.md "Kernels")
        n_ww_mol = counts[2] / 9.0
        alg = {
            "k_water_rows": n_ww_mol * FLOP_WW_MOL + 0.5 * counts[1] * FLOP_PAIR,
            "k_solute_rows": counts[0] * FLOP_PAIR + 0.5 * counts[1] * FLOP_PAIR,
            "k_q_partner": 0.5 * (counts[3] + counts[4]) * ns_q,
            "k_q_atom": 0.5 * (counts[3] + counts[4]) * ns_q,
            "k_qq_static": nqq * flop_q_pair(1),
            "k_lrf_taylor": q.natom * FLOP_LRF_TAYLOR,
            "k_pair_energy": 0.0,   # the energy sums are part of the 33 / 209 flop per pair counted with the row kernels
            "k_solvent_restraints": 0.0,
        }
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["bytes_per_launch"]
        except Exception:
            traffic = {}
        dom = max(kt, key=kt.get)
        peak_tf = fp32_meas
        ach = alg[dom] / (kt[dom] * 1e-3) / 1e12
        roof = {"bound": "fp32", "kernel": dom, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic.get(dom),   # dram bytes per launch from the committed ncu --set full capture
                "peak_source": "FP32 FMA micro-benchmark run in this process (MEASURED_PEAKS.json has no FP32 figure); "
                               f"nominal 148 SM x 128 x 2 x {peaks.get('sm_max_mhz', 1965.0)} MHz = "
                               f"{148 * 128 * 2 * peaks.get('sm_max_mhz', 1965.0) * 1e-6:.1f} TFLOP/s; FP64 pipe measured {fp64_meas:.1f}",
                "algorithmic_flop_per_launch": alg[dom], "ms_per_launch_l2_flushed": kt[dom],
                "ms_per_launch_warm": kt_warm.get(dom),
                "kernels_ms_l2_flushed": kt, "kernels_ms_warm": kt_warm,
                "step_fraction_all_kernels": sum(alg[k] for k in kt) / (sum(kt_warm.values()) * 1e-3) / 1e12 / peak_tf,
                "list_build": {"ms": list_ms,
                               "algorithmic_bytes": 24.0 * q.ncgp + 8.0 * float(g_total_unit_pairs(g, q)) + 320.0 * q.ncgp,
                               "hbm_peak_gbs": peaks.get("hbm_gbs")}}
        lb = roof["list_build"]
        lb["achieved_gbs"] = lb["algorithmic_bytes"] / (lb["ms"] * 1e-3) / 1e9
        if lb["hbm_peak_gbs"]:
            lb["frac"] = lb["achieved_gbs"] / lb["hbm_peak_gbs"]
        nsys = 1 if sharded else world          # systems advanced per step over the whole job
        value = npairs * nsys / t_step
        e2e_step = e2e_s / steps
        line = {"metric": "nonbonded pair interactions/s", "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
                "scaling": "strong" if sharded else "weak", "vs_baseline": None,
                "dtype": "f32 pair math, f64 accumulation and energies",
                "data": "synthetic", "config": config, "pairs_per_step": npairs,
                "ns_per_day": 86400.0 / e2e_step * DT_FS * 1e-6,
                "ns_per_day_device_resident": 86400.0 / t_step * DT_FS * 1e-6,
                "fep_windows_per_hour": nsys * 3600.0 / (STEPS_PER_WINDOW * e2e_step),
                "e2e": {"value": npairs * nsys / e2e_step, "unit": "pairs/s", "ms_per_step": e2e_step * 1e3,
                        "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": int(d2h),
                        "host_breakdown_last_call_us": {k: round(v * 1e6, 1) for k, v in host_breakdown.items()}},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
        if win_s is not None:
            agg = world * kwin * steps / win_s      # MD steps per second over all windows of all GPUs
            line["concurrent_windows"] = {"windows_per_gpu": kwin, "steps_per_s": agg, "value": npairs * agg, "unit": "pairs/s",
                                          "ms_per_step_per_window": win_s / steps * 1e3,
                                          "fep_windows_per_hour": 3600.0 * agg / STEPS_PER_WINDOW,
                                          "note": "same end-to-end path as e2e (host buffers, list build every "
                                                  f"{NBCYCLE} steps), {kwin} handles and host threads per GPU"}
        if world == 1:
            try:
                ms_c = compiled_host_e2e(q, cuts, lam, max(steps, 10 * NBCYCLE), dev)
                line["e2e_compiled_host"] = {"ms_per_step": ms_c, "value": npairs / (ms_c * 1e-3), "unit": "pairs/s",
                                             "note": "q6_b200/host/qdyn_nb (C++ host over the C ABI, input files written by "
                                                     "synth.write_files, host buffers, list build every "
                                                     f"{NBCYCLE} steps); separate process, timed by its own host clock"}
            except Exception as e:  # reported, never fatal: the headline e2e above does not depend on it
                line["e2e_compiled_host"] = {"error": str(e)[-200:]}
        # ---- optional paths, reported next to the headline (never part of it)
        ext = {} if sharded else {"device_step_ms": g.bench_nonbond(lam, 200) / 200,
                                  "device_step_ms_no_pp_pw_ww_energies": g.bench_nonbond(lam, 200, energies=False) / 200}
        if not sharded and not q.use_PBC and q.nwat > 0:
            from q6_b200.engine import wat_shells
            rw = float(np.linalg.norm(q.xtop[q.nat_solute:] - np.asarray(q.xpcent), axis=1).max())
            g.set_solvent_restraints(wat_shells(q.xpcent, rw, crgQtot=-1.0), np.zeros(8))
            ext["device_step_ms_with_solvent_restraints"] = g.bench_nonbond(lam, 200, restraints=True) / 200
        if world == 1 and q.nwat > 0:
            # N2: solvent SHAKE on the device (qnb_shake; xx = the coordinates resident from the step).  Reported, never
            # fatal: the kernel has not run on hardware before this bench.
            try:
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
    g.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def g_total_unit_pairs(g, q):
    """Charge-group pairs emitted by the list build (each listed pair once)."""
    ww = g.list_count(2) // 9
    # pp/pw group pairs are not exported separately; atom pairs / mean group size is a lower bound that the
    # roofline of the (tiny) byte figure does not depend on in any visible digit
    return ww + g.list_count(1) // 9 + g.list_count(0) // 9


if __name__ == "__main__":
    main()
