"""Python mirror of /root/reference/examples/2dIsing_dynamics.jl on the B200 engine: real-time TFIM dynamics on
a 5×5 square lattice with belief-propagation simple update, ⟨Z⟩ on the centre vertex after every Trotter layer.

    python examples/2dIsing_dynamics.py [nx ny [maxdim [nlayers]]]

Same constants as the reference example (:6-41): dt = 0.25, hx = 1.0, hz = 0.8, J = 0.5, ComplexF32,
apply_kwargs = (maxdim = 5, cutoff = 1e-10, normalize_tensors = false).  The boundary-MPS cross-check of the
original (:50,63-64) is host-side ITensor code and out of scope here; `tq.network(psi_bpc)` exports the state for it.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tnqs_b200 as tq  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    chi = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    no_trotter_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    g = tq.named_grid((nx, ny))
    dt, hx, hz, J = 0.25, 1.0, 0.8, 0.5
    # one Trotter layer: Rx, Rz on every vertex, then Rzz per edge-colour group (2dIsing_dynamics.jl:12-28)
    layer = [("Rx", [v], 2 * hx * dt) for v in g.vertices()]
    layer += [("Rz", [v], 2 * hz * dt) for v in g.vertices()]
    for colored_edges in tq.edge_color(g, 4):
        layer += [("Rzz", list(pair), 2 * J * dt) for pair in colored_edges]
    v_measure = ((nx + 1) // 2, (ny + 1) // 2)
    obs = ("Z", [v_measure])
    psi0 = tq.tensornetworkstate(np.complex64, lambda v: "↑", g, "S=1/2")
    psi_bpc = tq.BeliefPropagationCache(psi0)
    apply_kwargs = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=False)
    print(f"Max bond dimension of the TNS will be {chi}; measuring Z on {v_measure}")
    t0 = time.perf_counter()
    for l in range(1, no_trotter_steps + 1):
        psi_bpc, errors = tq.apply_gates(layer, psi_bpc, apply_kwargs=apply_kwargs, verbose=False)
        sz_bp = tq.expect(psi_bpc, obs)
        print(f"Layer {l}: maxvirtualdim {tq.maxvirtualdim(psi_bpc)}, max truncation error {errors.max():.3e}, "
              f"BP measured magnetisation {np.real(sz_bp):.8f}, BP norm² {abs(tq.norm_sqr(psi_bpc, alg='bp')):.6f}")
    print(f"Total time {time.perf_counter() - t0:.2f} s; bond entropy across {g.edges[0]}: "
          f"{tq.renyi_entropy(psi_bpc, g.edges[0], 1.0):.6f}")


if __name__ == "__main__":
    main()
