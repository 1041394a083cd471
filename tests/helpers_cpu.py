"""CPU-only helpers (no device calls): circuits as dense matrices for the oracle."""
import numpy as np

import tnqs_b200 as tq


def tfim_layer_cpu(g, dt=0.25, hx=1.0, hz=0.8, J=0.5, ncol=4):
    layer = [("Rx", [v], 2 * hx * dt) for v in g.vertices()]
    layer += [("Rz", [v], 2 * hz * dt) for v in g.vertices()]
    for grp in tq.edge_color(g, ncol):
        layer += [("Rzz", list(pair), 2 * J * dt) for pair in grp]
    nverts, verts, mats = tq.circuit_arrays(layer, g)
    mc = mats.view(np.complex128)
    gm, off = [], 0
    for n in nverts:
        k = 4 ** int(n)
        gm.append(mc[off:off + k].reshape(2 ** int(n), 2 ** int(n)))
        off += k
    gv = [[int(x) for x in v[:n]] for v, n in zip(verts, nverts)]
    return layer, gm, gv
