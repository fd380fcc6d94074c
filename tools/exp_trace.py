"""Per-warp timeline of the persistent force kernels (library built with -DQNB_TRACE; design experiment).
build: nvcc <flags> -DQNB_TRACE -o tools/exp/libqnb_trace.so q6_b200/csrc/qnb.cu -ldl"""
import os, sys, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from q6_b200 import synth, engine
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
q, cuts, lam = synth.config(wl)
lib = engine.load_library(os.path.abspath("tools/exp/libqnb_trace.so"))
g = engine.Qnb(q, lib=lib)
g.make_pair_lists(q.xtop, **cuts, counts=False)
for _ in range(5): g.pot_energy_nonbonds(q.xtop, lam)
lib.qnb_trace_clear()
g.pot_energy_nonbonds(q.xtop, lam)
buf = np.zeros((2, 8192, 10), dtype=np.uint64)
lib.qnb_trace_read(buf.ctypes.data_as(ctypes.c_void_p))
t_all0 = min(int(buf[k, :, 0][buf[k, :, 0] > 0].min()) for k in range(2) if (buf[k, :, 0] > 0).any())
for k, name in enumerate(["water", "solute"]):
    b = buf[k]; m = b[:, 0] > 0
    if not m.any(): continue
    b = b[m].astype(np.int64)
    done = b[:, 3] > 0
    t0 = b[:, 0].min()
    start = (b[:, 0] - t0) / 1e3
    end = (b[done, 3] - t0) / 1e3
    dur_cyc = (b[done, 2] - b[done, 1])
    nchunk = b[done, 4]
    print(f"{name}: warps {m.sum()} working {done.sum()}  kernel first-warp offset vs step {(t0 - t_all0)/1e3:.1f} us")
    if k == 1: print("   block prologue (smem tables) us: mean %.2f max %.2f" % (((b[:, 0] - b[:, 5]) / 1e3).mean(), ((b[:, 0] - b[:, 5]) / 1e3).max()))
    print("   warp start us  : min %.1f p50 %.1f p90 %.1f max %.1f" % (start.min(), np.median(start), np.percentile(start, 90), start.max()))
    print("   warp end us    : min %.1f p50 %.1f p90 %.1f max %.1f" % (end.min(), np.median(end), np.percentile(end, 90), end.max()))
    print("   warp busy cyc  : min %d p50 %d p90 %d max %d   chunks/warp min %d max %d" % (dur_cyc.min(), np.median(dur_cyc), np.percentile(dur_cyc, 90), dur_cyc.max(), nchunk.min(), nchunk.max()))
    per = dur_cyc / np.maximum(nchunk, 1)
    print("   cycles per chunk: p10 %d p50 %d p90 %d max %d" % (np.percentile(per, 10), np.median(per), np.percentile(per, 90), per.max()))
    # least-squares cost of a tile load and of the three chunk kinds (cycles), for chunk_cost() in qnb_lists.cuh
    A = b[done][:, 6:10].astype(float)
    A = np.hstack([A, np.ones((A.shape[0], 1))])
    coef, res, rk, sv = np.linalg.lstsq(A, dur_cyc.astype(float), rcond=None)
    pred = A @ coef
    print("   fit cycles: tile %.0f own %.0f mirror %.0f other %.0f const %.0f   rms residual %.0f (busy mean %.0f)" % (*coef, np.sqrt(np.mean((pred - dur_cyc) ** 2)), dur_cyc.mean()))
