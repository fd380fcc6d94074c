"""k independent lambda windows per GPU, one host thread + handle each (design experiment):
aggregate MD steps/s (list build every 25 steps, host buffers in and out) against one window alone."""
import os, sys, time, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q6_b200 import synth, engine
wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
q, cuts, lam = synth.config(wl)
nsteps = 500

def run(handle, lam_w, out, k):
    x = q.xtop.copy()
    d = np.zeros((q.natom, 3))
    for it in range(nsteps):
        if it % 25 == 0: handle.make_pair_lists(x, **cuts, counts=False)
        d[:] = 0
        handle.pot_energy_nonbonds(x, lam_w, d=d)
    out[k] = 1

for nw in (1, 2, 4, 8):
    hs = [engine.Qnb(q) for _ in range(nw)]
    lams = [np.array([1.0 - 0.02 * k, 0.02 * k])[:q.nstates] if q.nstates == 2 else np.array(lam) for k in range(nw)]
    for h, l in zip(hs, lams):
        h.make_pair_lists(q.xtop, **cuts, counts=False); h.pot_energy_nonbonds(q.xtop, l)
    out = [0] * nw
    th = [threading.Thread(target=run, args=(hs[k], lams[k], out, k)) for k in range(nw)]
    t = time.perf_counter()
    for a in th: a.start()
    for a in th: a.join()
    dt = time.perf_counter() - t
    print(f"{wl}: {nw} windows on one GPU: {nw * nsteps / dt:9.0f} steps/s aggregate, {dt / nsteps * 1e6:7.1f} us per step of each window")
    for h in hs: h.close()
