"""Saturated 16x16 chi=32: 12 layers with TNQS_SLOWLOG=1 to catch sporadic host-side stalls."""
import sys, os, time
os.environ["TNQS_SLOWLOG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqs_b200 as tq
L, chi = 16, 32
g = tq.named_grid((L, L))
layer = [("Rx", [v], 0.5) for v in g.vertices()] + [("Rz", [v], 0.4) for v in g.vertices()]
for grp in tq.edge_color(g, 4):
    layer += [("Rzz", list(p), 0.25) for p in grp]
seq = tq.bipartite_edge_sequence(g)
psi = tq.BeliefPropagationCache(tq.zerostate(np.complex64, g))
kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
bp = dict(maxiter=25, tolerance=1e-5, edge_sequence=seq)
for l in range(15):
    psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=True)
print("---- saturated, timing", flush=True)
for l in range(14):
    inplace = l < 7
    psi.stats(reset=True)
    t0 = time.perf_counter()
    psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=inplace)
    t1 = time.perf_counter()
    st = psi.stats()
    print("layer %2d %s wall %.1f dev %.1f (bp %.1f su %.1f)" % (l, "inplace" if inplace else "clone  ", (t1 - t0) * 1e3, st["bp_ms"] + st["su_ms"], st["bp_ms"], st["su_ms"]), flush=True)
