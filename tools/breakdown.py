import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqs_b200 as tq
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 32
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 16
g = tq.named_grid((L, L))
layer = [("Rx", [v], 0.5) for v in g.vertices()] + [("Rz", [v], 0.4) for v in g.vertices()]
for grp in tq.edge_color(g, 4):
    layer += [("Rzz", list(p), 0.25) for p in grp]
seq = tq.bipartite_edge_sequence(g)
psi = tq.BeliefPropagationCache(tq.zerostate(np.complex64, g))
kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
bp = dict(maxiter=25, tolerance=1e-5, edge_sequence=seq)
for l in range(nl):
    psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=True)
fam = ["mode_ms", "gram_ms", "small_ms", "mode_flops", "gram_flops", "mode_launches", "gram_launches", "kernel_launches", "bp_ms", "su_ms"]
# un-profiled layer
psi.stats(reset=True)
t0 = time.perf_counter(); psi, _ = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=True); t1 = time.perf_counter()
st = psi.stats(); print("layer unprofiled wall %.1f ms, bp %.1f su %.1f" % ((t1 - t0) * 1e3, st["bp_ms"], st["su_ms"]))
psi.set_profiling(True)
psi.stats(reset=True)
t0 = time.perf_counter(); tq.update(psi, inplace=True, maxiter=1, edge_sequence=seq); t1 = time.perf_counter()
st = psi.stats(); print("BP sweep (profiled) wall %.1f ms:" % ((t1 - t0) * 1e3), {k: round(st[k], 2) for k in fam})
print("   mode TF/s %.1f gram TF/s %.1f" % (st["mode_flops"] / st["mode_ms"] / 1e9, st["gram_flops"] / st["gram_ms"] / 1e9))
psi.stats(reset=True)
two = [gt for gt in layer if len(gt[1]) == 2][:120]
t0 = time.perf_counter(); psi, _ = tq.apply_gates(two, psi, apply_kwargs=kw, update_cache=False, inplace=True); t1 = time.perf_counter()
st = psi.stats(); print("SU segment of %d gates (profiled) wall %.1f ms:" % (len(two), (t1 - t0) * 1e3), {k: round(st[k], 2) for k in fam})
print("   mode TF/s %.1f gram TF/s %.1f" % (st["mode_flops"] / st["mode_ms"] / 1e9, st["gram_flops"] / st["gram_ms"] / 1e9))
psi.set_profiling(False)
psi.stats(reset=True)
t0 = time.perf_counter(); tq.update(psi, inplace=True, maxiter=1, edge_sequence=seq); t1 = time.perf_counter()
st = psi.stats(); print("BP sweep (unprofiled) wall %.1f ms bp_ms %.1f" % ((t1 - t0) * 1e3, st["bp_ms"]))
two = [gt for gt in layer if len(gt[1]) == 2][120:240]
psi.stats(reset=True)
t0 = time.perf_counter(); psi, _ = tq.apply_gates(two, psi, apply_kwargs=kw, update_cache=False, inplace=True); t1 = time.perf_counter()
st = psi.stats(); print("SU segment (unprofiled) wall %.1f ms su_ms %.1f" % ((t1 - t0) * 1e3, st["su_ms"]))
one = [gt for gt in layer if len(gt[1]) == 1]
psi.stats(reset=True)
t0 = time.perf_counter(); psi, _ = tq.apply_gates(one, psi, apply_kwargs=kw, update_cache=False, inplace=True); t1 = time.perf_counter()
st = psi.stats(); print("one-site gates (unprofiled) wall %.1f ms su_ms %.1f" % ((t1 - t0) * 1e3, st["su_ms"]))
