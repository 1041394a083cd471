"""Saturated 16x16 chi=32 state: time one SU colour (profiled families) — for kernel tuning via env vars."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqs_b200 as tq
L, chi = 16, 32
g = tq.named_grid((L, L))
layer = [("Rx", [v], 0.5) for v in g.vertices()] + [("Rz", [v], 0.4) for v in g.vertices()]
groups = tq.edge_color(g, 4)
for grp in groups:
    layer += [("Rzz", list(p), 0.25) for p in grp]
seq = tq.bipartite_edge_sequence(g)
psi = tq.BeliefPropagationCache(tq.zerostate(np.complex64, g))
kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
bp = dict(maxiter=25, tolerance=1e-5, edge_sequence=seq)
for l in range(15):
    psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=True)
psi.set_profiling(True)
if os.environ.get("PROBE_DEBUG"):
    os.environ["TNQS_JACOBI_DEBUG"] = "1"
for ci in range(2):
    two = [("Rzz", list(p), 0.25) for p in groups[ci]]
    psi.stats(reset=True)
    psi, _ = tq.apply_gates(two, psi, apply_kwargs=kw, update_cache=False, inplace=True)
    st = psi.stats()
    print("colour %d: su %.1f ms mode %.1f gram %.1f small %.1f" % (ci, st["su_ms"], st["mode_ms"], st["gram_ms"], st["small_ms"]), flush=True)
    os.environ.pop("TNQS_JACOBI_DEBUG", None)
    tq.update(psi, inplace=True, maxiter=1, edge_sequence=seq)
