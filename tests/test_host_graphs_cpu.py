"""Host-side graph logic that feeds the C-ABI integer tables (no device calls): the lattices of the BASELINE configs, the edge
colourings that define the gate batches, and the BP edge schedules (NamedGraphs / SimpleGraphAlgorithms
on the reference side)."""
import collections

import numpy as np
import pytest

import tnqs_b200 as tq


def _degrees(g):
    deg = collections.Counter()
    for a, b in g.edges:
        deg[a] += 1
        deg[b] += 1
    return deg


def _connected(g):
    adj = collections.defaultdict(list)
    for a, b in g.edges:
        adj[a].append(b)
        adj[b].append(a)
    vs = g.vertices()
    seen, stack = {vs[0]}, [vs[0]]
    while stack:
        x = stack.pop()
        for y in adj[x]:
            if y not in seen:
                seen.add(y)
                stack.append(y)
    return len(seen) == g.nv


def test_eagle_heavy_hex_is_the_127_qubit_graph():
    g = tq.eagle_heavy_hex()
    assert g.nv == 127 and g.ne == 144 and _connected(g)
    deg = _degrees(g)
    assert max(deg.values()) == 3
    kinds = collections.Counter(tuple(sorted((deg[a], deg[b]))) for a, b in g.edges)
    assert kinds == {(2, 3): 106, (2, 2): 36, (1, 3): 2}
    assert g.bipartition() is not None


@pytest.mark.parametrize("name,g,k,sizes", [
    ("16x16", lambda: tq.named_grid((16, 16)), 4, [120, 120, 120, 120]),
    ("eagle", tq.eagle_heavy_hex, 3, [48, 48, 48]),
    ("6x6x6 periodic", lambda: tq.named_grid((6, 6, 6), periodic=True), 6, [108] * 6),
])
def test_edge_colourings_are_proper_and_cover_every_edge(name, g, k, sizes):
    g = g()
    groups = tq.edge_color(g, k)
    assert len(groups) == k
    seen = set()
    for grp in groups:
        touched = set()
        for a, b in grp:
            assert a not in touched and b not in touched  # vertex-disjoint: one batched launch per colour
            touched.update((a, b))
            seen.add(frozenset((a, b)))
    assert seen == {frozenset(e) for e in g.edges}
    assert sum(len(grp) for grp in groups) == g.ne == sum(sizes)


def test_periodic_cubic_lattice_of_config_4():
    g = tq.named_grid((6, 6, 6), periodic=True)
    assert g.nv == 216 and g.ne == 648 and set(_degrees(g).values()) == {6}


@pytest.mark.parametrize("g", [tq.named_grid((5, 5)), tq.named_grid((4, 3, 2)), tq.eagle_heavy_hex(), tq.named_comb_tree((3, 3))])
def test_bp_schedules_visit_every_directed_edge_once(g):
    want = collections.Counter()
    for a, b in g.edges:
        want[(a, b)] += 1
        want[(b, a)] += 1
    for seq in (tq.forest_cover_edge_sequence(g), tq.bipartite_edge_sequence(g)):
        assert collections.Counter(tuple(e) for e in seq) == want
    # bipartite schedule: all messages leaving one colour class, then the other (two dependency levels on the device)
    col = g.bipartition()
    seq = tq.bipartite_edge_sequence(g)
    classes = [col[g.index[a]] for a, _ in seq]
    assert classes == sorted(classes)


def test_circuit_arrays_memo_gives_the_same_matrices():
    g = tq.named_grid((3, 3))
    vs = g.vertices()
    circuit = [("Rx", [vs[0]], 0.3), ("Rx", [vs[1]], 0.3), ("Rx", [vs[2]], 0.7), ("Rzz", [vs[0], vs[1]], 0.3),
               ("Rzz", [vs[1], vs[2]], 0.3), ("CPHASE", [vs[3], vs[4]], 0.2), ("X", [vs[5]])]
    nverts, verts, mats = tq.circuit_arrays(circuit, g)
    mc = mats.view(np.complex128)
    off = 0
    for gate, n in zip(circuit, nverts):
        k = 4 ** int(n)
        ref = tq.gate_matrix(gate[0], int(n), gate[2] if len(gate) > 2 else None)
        assert np.array_equal(mc[off:off + k].reshape(ref.shape), ref)
        off += k
    assert off == mc.size and list(nverts) == [1, 1, 1, 2, 2, 2, 1]
