import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqs_b200 as tq
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 32
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 50
sched = sys.argv[4] if len(sys.argv) > 4 else "bipartite"
g = tq.named_grid((L, L))
layer = [("Rx", [v], 0.5) for v in g.vertices()] + [("Rz", [v], 0.4) for v in g.vertices()]
for grp in tq.edge_color(g, 4):
    layer += [("Rzz", list(p), 0.25) for p in grp]
seq = tq.bipartite_edge_sequence(g) if sched == "bipartite" else tq.forest_cover_edge_sequence(g)
psi = tq.BeliefPropagationCache(tq.zerostate(np.complex64, g))
kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
bp = dict(maxiter=25, tolerance=1e-5, edge_sequence=seq)
tot = 0
for l in range(nl):
    psi.stats(reset=True)
    t0 = time.perf_counter()
    psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=True)
    dt = time.perf_counter() - t0
    st = psi.stats()
    bd = psi.bond_dims()
    z = tq.expect(psi, ("Z", [(L // 2 + 1, L // 2 + 1)]))
    tot += dt
    print(json.dumps(dict(layer=l + 1, wall_s=round(dt, 3), dev_ms=round(st["bp_ms"] + st["su_ms"], 1), bp_ms=round(st["bp_ms"], 1),
                          su_ms=round(st["su_ms"], 1), sweeps=[r["niter"] for r in psi.last_bp_reports], bond=[int(bd.min()), round(float(bd.mean()), 1), int(bd.max())],
                          maxerr=float(errs.max()), z=round(float(np.real(z)), 6), launches=st["kernel_launches"])), flush=True)
    if tot > 400: break
