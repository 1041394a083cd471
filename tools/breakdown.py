"""Per-phase timing of one saturated layer: device time (CUDA events on the engine stream), host wall
time inside the library (stats wall_ms) and per-kernel-family time in profiling mode."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tnqs_b200 as tq
L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rand = len(sys.argv) > 3 and sys.argv[3] == "random"  # synthetic random TNS with every bond = chi (generated on the device)
nl = 0 if rand else (int(sys.argv[3]) if len(sys.argv) > 3 else 16)
g = tq.named_grid((L, L))
layer = [("Rx", [v], 0.5) for v in g.vertices()] + [("Rz", [v], 0.4) for v in g.vertices()]
groups = tq.edge_color(g, 4)
for grp in groups:
    layer += [("Rzz", list(p), 0.25) for p in grp]
seq = tq.bipartite_edge_sequence(g)
kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True)
bp = dict(maxiter=25, tolerance=1e-5, edge_sequence=seq)
if rand:
    psi = tq.update(tq.random_bpc_on_device(np.complex64, g, bond_dimension=chi, seed=1234), inplace=True, **bp)
else:
    psi = tq.BeliefPropagationCache(tq.zerostate(np.complex64, g))
for l in range(nl):
    psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=True)
fam = ["mode_ms", "gram_ms", "small_ms", "mode_launches", "gram_launches", "kernel_launches", "bp_ms", "su_ms", "wall_ms"]


def show(tag, t0, t1, st):
    print("%-34s wall %7.1f ms | lib wall %7.1f (host busy %6.1f) dev bp %7.1f su %7.1f | mode %6.1f gram %6.1f small %6.1f | launches %d" % (
        tag, (t1 - t0) * 1e3, st["wall_ms"], st["wall_ms"] - st["sync_ms"], st["bp_ms"], st["su_ms"], st["mode_ms"], st["gram_ms"], st["small_ms"], st["kernel_launches"]))


for rep in range(1 if rand else 2):
    psi.stats(reset=True)
    t0 = time.perf_counter(); psi, _ = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp, inplace=True); t1 = time.perf_counter()
    show("layer (unprofiled)", t0, t1, psi.stats())
for prof in (False, True, False):
    psi.set_profiling(prof)
    tagp = "profiled" if prof else "unprofiled"
    psi.stats(reset=True)
    t0 = time.perf_counter(); tq.update(psi, inplace=True, maxiter=1, edge_sequence=seq); t1 = time.perf_counter()
    st = psi.stats(); show("BP sweep (%s)" % tagp, t0, t1, st)
    if prof:
        print("   mode TF/s %.1f GB/s %.0f | gram TF/s %.1f GB/s %.0f" % (st["mode_flops"] / st["mode_ms"] / 1e9, st["mode_bytes"] / st["mode_ms"] / 1e6,
                                                                   st["gram_flops"] / st["gram_ms"] / 1e9, st["gram_bytes"] / st["gram_ms"] / 1e6))
    for ci, grp in enumerate(groups[:2]):
        two = [("Rzz", list(p), 0.25) for p in grp]
        psi.stats(reset=True)
        t0 = time.perf_counter(); psi, _ = tq.apply_gates(two, psi, apply_kwargs=kw, update_cache=False, inplace=True); t1 = time.perf_counter()
        st = psi.stats(); show("SU colour %d, %d gates (%s)" % (ci, len(two), tagp), t0, t1, st)
        if prof:
            print("   mode TF/s %.1f GB/s %.0f | gram TF/s %.1f GB/s %.0f" % (st["mode_flops"] / st["mode_ms"] / 1e9, st["mode_bytes"] / st["mode_ms"] / 1e6,
                                                                       st["gram_flops"] / st["gram_ms"] / 1e9, st["gram_bytes"] / st["gram_ms"] / 1e6))
        psi.stats(reset=True)
        t0 = time.perf_counter(); tq.update(psi, inplace=True, maxiter=1, edge_sequence=seq); t1 = time.perf_counter()
        show("  BP sweep after it (%s)" % tagp, t0, t1, psi.stats())
psi.set_profiling(False)
one = [gt for gt in layer if len(gt[1]) == 1]
psi.stats(reset=True)
t0 = time.perf_counter(); psi, _ = tq.apply_gates(one, psi, apply_kwargs=kw, update_cache=False, inplace=True); t1 = time.perf_counter()
show("one-site gates", t0, t1, psi.stats())
