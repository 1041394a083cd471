"""Saturated 8x8 chi=32 state, one SU colour with TNQS_JACOBI_DEBUG=1: prints sweeps per Jacobi launch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqs_b200 as tq
L, chi = 8, 32
g = tq.named_grid((L, L))
layer = [("Rx", [v], 0.5) for v in g.vertices()] + [("Rz", [v], 0.4) for v in g.vertices()]
groups = tq.edge_color(g, 4)
for grp in groups:
    layer += [("Rzz", list(p), 0.25) for p in grp]
seq = tq.bipartite_edge_sequence(g)
psi = tq.BeliefPropagationCache(tq.zerostate(np.complex64, g))
kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
bp = dict(maxiter=25, tolerance=1e-5, edge_sequence=seq)
os.environ.pop("TNQS_JACOBI_DEBUG", None)
for l in range(14):
    psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=True)
print("bond dims", psi.bond_dims().min(), psi.bond_dims().max(), flush=True)
psi.set_profiling(True)
os.environ["TNQS_JACOBI_DEBUG"] = "1"
two = [("Rzz", list(p), 0.25) for p in groups[0]]
psi.stats(reset=True)
psi, _ = tq.apply_gates(two, psi, apply_kwargs=kw, update_cache=False, inplace=True)
print(psi.stats())
