"""Shared helpers for the parity tests: build the same inputs for the device path and the oracle."""
import numpy as np

import tnqs_b200 as tq
from oracle import tnqs_oracle as orc

Z = np.diag([1.0, -1.0]).astype(complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)


def ragged_state(g, dims, dtype, seed=0, d=2):
    """Random TNS with bond dimension dims[e] on edge e (ragged on purpose: catches stride mix-ups)."""
    rng = np.random.default_rng(seed)
    tensors = {}
    for i, v in enumerate(g.vertices()):
        shp = (d,) + tuple(int(dims[e]) for e, _ in g.incident[i])
        t = rng.standard_normal(shp) + 1j * rng.standard_normal(shp)
        tensors[v] = (t / np.linalg.norm(t)).astype(dtype)
    return tq.TensorNetworkState(g, tensors, dtype)


def oracle_from_tns(psi):
    g = psi.graph
    return orc.OracleCache(g.nv, g.edge_uv(), [psi.tensors[v] for v in g.vertices()], psi.dtype)


def oracle_from_bpc(bpc):
    """Download a device cache (tensors + set messages) into an OracleCache."""
    g = bpc.graph
    c = orc.OracleCache(g.nv, g.edge_uv(), [bpc.site(v) for v in g.vertices()], bpc.dtype)
    for (a, b), m in bpc.messages().items():
        c.msg[(g.index[a], g.index[b])] = m
    return c


def random_psd_messages(g, dims, dtype, seed=1):
    rng = np.random.default_rng(seed)
    out = {}
    for e, (a, b) in enumerate(g.edges):
        for edge in ((a, b), (b, a)):
            n = int(dims[e])
            w = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
            m = w @ w.conj().T + 0.1 * np.eye(n)
            out[edge] = (m / m.sum()).astype(dtype)
    return out


def circuit_for_oracle(g, circuit):
    nverts, verts, mats = tq.circuit_arrays(circuit, g)
    mc = mats.view(np.complex128)
    gm, off = [], 0
    for n in nverts:
        k = 4 ** int(n)
        gm.append(mc[off:off + k].reshape(2 ** int(n), 2 ** int(n)))
        off += k
    gv = [[int(x) for x in v[:n]] for v, n in zip(verts, nverts)]
    return gm, gv


def seq_idx(g, seq):
    return [(g.index[a], g.index[b]) for a, b in seq]


def tfim_layer(g, dt=0.25, hx=1.0, hz=0.8, J=0.5, ncol=4):
    """The layer of /root/reference/examples/2dIsing_dynamics.jl:20-28."""
    layer = [("Rx", [v], 2 * hx * dt) for v in g.vertices()]
    layer += [("Rz", [v], 2 * hz * dt) for v in g.vertices()]
    for grp in tq.edge_color(g, ncol):
        layer += [("Rzz", list(pair), 2 * J * dt) for pair in grp]
    return layer


def state_overlap(c1, c2):
    """|⟨ψ1|ψ2⟩| / (‖ψ1‖‖ψ2‖) and the two norms, by dense contraction (small systems)."""
    p1, p2 = orc.to_statevector(c1), orc.to_statevector(c2)
    n1, n2 = np.linalg.norm(p1), np.linalg.norm(p2)
    return abs(np.vdot(p1, p2)) / (n1 * n2), n1, n2


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
