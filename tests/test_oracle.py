"""Pins the CPU oracle (oracle/tnqs_oracle.py) to the reference's own test invariants
(the reference holds no numeric goldens for this path) and to an independent
dense state-vector simulator."""
import math

import numpy as np
import pytest

import tnqs_b200 as tq
from oracle import tnqs_oracle as orc
from oracle.statevector import StateVector

UP, DN = (1.0, 0.0), (0.0, 1.0)
Z = np.diag([1.0, -1.0])
X = np.array([[0, 1], [1, 0.0]])
Y = np.array([[0, -1j], [1j, 0]])


def circuit_to_arrays(g, circuit):
    mats, verts = [], []
    for gate in circuit:
        vs = gate[1] if isinstance(gate[1], list) else [gate[1]]
        mats.append(tq.gate_matrix(gate[0], len(vs), gate[2] if len(gate) > 2 else None))
        verts.append([g.index[v] for v in vs])
    return mats, verts


def seq_idx(g, seq):
    return [(g.index[a], g.index[b]) for a, b in seq]


def tfim_layer(g, dt=0.25, hx=1.0, hz=0.8, J=0.5, ncol=4):
    layer = [("Rx", [v], 2 * hx * dt) for v in g.vertices()]
    layer += [("Rz", [v], 2 * hz * dt) for v in g.vertices()]
    for grp in tq.edge_color(g, ncol):
        layer += [("Rzz", list(pair), 2 * J * dt) for pair in grp]
    return layer


def test_two_qubit_circuit_invariants():
    # /root/reference/test/test_apply.jl:11-20
    circuit = [("Rx", [(1, 1)], 0.5), ("Rx", [(2, 1)], 0.2), ("CPHASE", [(1, 1), (2, 1)], -0.3)]
    g = tq.build_graph_from_circuit(circuit)
    c = orc.product_state(g.nv, g.edge_uv(), [DN, DN], np.complex64)
    seq = seq_idx(g, tq.forest_cover_edge_sequence(g))
    c, _ = orc.bp_update(c, seq, **orc.default_bp_update_kwargs(c))
    mats, verts = circuit_to_arrays(g, circuit)
    c2, errs, _ = orc.apply_gates(c, mats, verts, seq,
                                  dict(maxdim=2, cutoff=1e-10, normalize_tensors=False))
    assert c2.dtype == np.complex64 and c2.T[0].dtype == np.complex64
    assert c2.maxvirtualdim() <= 2
    psi = orc.to_statevector(c2)
    assert abs(np.vdot(psi, psi) - 1.0) < 1e-5
    sv = StateVector([DN, DN])
    for m, vs in zip(mats, verts):
        sv.apply(m, vs)
    assert abs(abs(np.vdot(sv.psi, psi)) - 1.0) < 1e-5


@pytest.mark.parametrize("dtype,tol", [(np.complex64, 2e-5), (np.complex128, 1e-11)])
def test_grid_tfim_layer_norm_and_statevector(dtype, tol):
    # /root/reference/test/test_apply.jl:23-53 — untruncated SU on a loopy graph is exact
    g = tq.named_grid((3, 3))
    rng = np.random.default_rng(123)
    locs = []
    for _ in range(g.nv):
        a = rng.standard_normal(2) + 1j * rng.standard_normal(2)
        locs.append(a / np.linalg.norm(a))
    c = orc.product_state(g.nv, g.edge_uv(), locs, dtype)
    seq = seq_idx(g, tq.forest_cover_edge_sequence(g))
    mats, verts = circuit_to_arrays(g, tfim_layer(g))
    c2, errs, reps = orc.apply_gates(c, mats, verts, seq, dict(cutoff=1e-10, normalize_tensors=False))
    assert c2.maxvirtualdim() <= 2
    assert len(reps) == 5  # 4 colour groups + final refresh
    assert np.all(errs < 1e-9)
    psi = orc.to_statevector(c2)
    assert abs(np.vdot(psi, psi) - 1.0) < 50 * tol
    sv = StateVector(locs)
    for m, vs in zip(mats, verts):
        sv.apply(m, vs)
    assert abs(abs(np.vdot(sv.psi, psi)) - 1.0) < 50 * tol


def test_segmentation_rule():
    # /root/reference/src/Apply/apply_gates.jl:60-90
    g = tq.named_grid((4, 4))
    mats, verts = circuit_to_arrays(g, tfim_layer(g))
    fires = orc.segment_gates(verts)
    assert len(fires) == 4
    n1 = 2 * g.nv
    assert fires[0] == n1  # first Rzz fires: every vertex was touched by the one-site gates
    # inside a segment all two-site gates are vertex disjoint
    bounds = fires + [len(verts)]
    for a, b in zip(bounds[:-1], bounds[1:]):
        vs = [v for k in range(a, b) for v in verts[k]]
        assert len(vs) == len(set(vs))
    # one-site gate after a two-site gate on the same vertex does not fire
    assert orc.segment_gates([[0, 1], [0], [2, 3], [1, 2]]) == [3]


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_bp_exact_on_tree(dtype):
    # /root/reference/test/test_beliefpropagation.jl:31-55
    g = tq.named_comb_tree((3, 3))
    c = orc.random_state(g.nv, g.edge_uv(), 2, 2, dtype, seed=123)
    seq = seq_idx(g, tq.forest_cover_edge_sequence(g))
    assert len(seq) == 2 * g.ne and len(set(seq)) == 2 * g.ne
    assert not c.msg
    c, rep = orc.bp_update(c, seq, **orc.default_bp_update_kwargs(c))
    assert rep["niter"] == 1 and len(c.msg) == 2 * g.ne
    psi = orc.to_statevector(c)
    vc = g.index[g.center()[0]]
    rho_exact = np.moveaxis(psi, vc, 0).reshape(2, -1)
    rho_exact = rho_exact @ rho_exact.conj().T
    rho_exact /= np.trace(rho_exact)
    rho_bp = orc.rdm_local(c, vc).astype(np.complex128)
    rho_bp /= np.trace(rho_bp)
    eps = np.finfo(np.float32 if dtype == np.complex64 else np.float64).eps
    assert np.linalg.norm(rho_bp - rho_exact) <= 10 * eps


def test_expect_bp_vs_exact_tree_and_loopy():
    # /root/reference/test/test_expect.jl:10-45
    for g, is_tree in ((tq.named_path_graph(6), True), (tq.named_grid((3, 3)), False)):
        c = orc.random_state(g.nv, g.edge_uv(), 2, 2, np.complex128, seed=7)
        seq = seq_idx(g, tq.forest_cover_edge_sequence(g))
        c, _ = orc.bp_update(c, seq, maxiter=100, tolerance=1e-14)
        psi = orc.to_statevector(c)
        sv = StateVector([UP] * g.nv)
        sv.psi = psi
        v = g.nv // 2
        e_bp = orc.expect_local(c, v, Z)
        e_ex = sv.expect(Z, [v])
        if is_tree:
            assert abs(e_bp - e_ex) < 1e-12
            w = c.incident[v][0][1]
            assert abs(orc.expect_two_site(c, v, w, X, Y) - sv.expect(np.kron(X, Y), [v, w])) < 1e-12
        else:
            assert abs(e_bp - e_ex) > 1e-6


def test_tree_dynamics_matches_statevector_with_messages_fixed_point():
    # SU without truncation on a tree is exact, and the messages written by
    # apply_gate! (diag σ) equal freshly recomputed BP messages up to normalisation.
    g = tq.named_comb_tree((2, 3))
    c = orc.product_state(g.nv, g.edge_uv(), [UP] * g.nv, np.complex128)
    seq = seq_idx(g, tq.forest_cover_edge_sequence(g))
    layer = [("Rx", [v], 0.7) for v in g.vertices()]
    for grp in tq.edge_color(g, 3):
        layer += [("Rxx", list(p), 0.9) for p in grp]
    mats, verts = circuit_to_arrays(g, layer)
    sv = StateVector([UP] * g.nv)
    for _ in range(3):
        c, errs, _ = orc.apply_gates(c, mats, verts, seq, dict(cutoff=1e-14, normalize_tensors=True))
        for m, vs in zip(mats, verts):
            sv.apply(m, vs)
    for v in range(g.nv):
        assert abs(orc.expect_local(c, v, Z) - sv.expect(Z, [v])) < 1e-10
    # gate then compare its written messages with a recomputed BP fixed point
    v1, v2 = c.edges[0]
    c3 = c.copy()
    orc.apply_gate(c3, tq.gate_matrix("Rzz", 2, 0.4), [v1, v2], cutoff=1e-14)
    m_written = c3.msg[(v1, v2)]
    m_bp = orc.updated_message(c3, v1, v2)
    assert orc.message_diff(m_written, m_bp) < 1e-12


def test_truncate_spectrum_semantics():
    p = np.array([0.5, 0.3, 0.15, 0.04, 0.01])
    assert orc.truncate_spectrum(p, None, None) == (5, 0.0)
    n, e = orc.truncate_spectrum(p, 3, None)
    assert n == 3 and abs(e - 0.05) < 1e-15
    n, e = orc.truncate_spectrum(p, None, 0.02)  # 0.01 ≤ 0.02 dropped; 0.01+0.04 > 0.02 kept
    assert n == 4 and abs(e - 0.01) < 1e-15
    n, e = orc.truncate_spectrum(p, 4, 0.2)      # maxdim first, then cutoff continues: .01+.04+.15=.2 ≤ .2
    assert n == 2 and abs(e - 0.2) < 1e-15
    n, e = orc.truncate_spectrum(np.array([1.0, 0.0, 0.0]), None, 1e-10)
    assert n == 1 and e == 0.0
    n, e = orc.truncate_spectrum(np.array([0.0, 0.0]), None, 1e-10)  # mindim = 1
    assert n == 1


def test_truncation_maxdim_respected_and_error_positive():
    # /root/reference/test/test_truncate.jl:29-33 flavour: χ ≤ maxdim, 0 ≤ err ≤ 1
    g = tq.named_grid((3, 3))
    c = orc.product_state(g.nv, g.edge_uv(), [UP] * g.nv, np.complex128)
    seq = seq_idx(g, tq.forest_cover_edge_sequence(g))
    mats, verts = circuit_to_arrays(g, tfim_layer(g))
    allerrs = []
    for _ in range(3):
        c, errs, _ = orc.apply_gates(c, mats, verts, seq, dict(maxdim=2, cutoff=1e-12))
        allerrs.append(errs)
    assert c.maxvirtualdim() <= 2
    allerrs = np.concatenate(allerrs)
    assert np.all(allerrs >= 0) and np.all(allerrs <= 1) and allerrs.max() > 1e-8


def test_pseudo_sqrt_inv_sqrt():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((6, 3)) + 1j * rng.standard_normal((6, 3))
    m = (a @ a.conj().T).astype(np.complex64)  # rank 3 PSD
    A, B = orc.pseudo_sqrt_inv_sqrt(m, 10 * np.finfo(np.float32).eps * np.trace(m).real)
    P = A.astype(complex) @ B.astype(complex)
    assert np.allclose(P @ P, P, atol=1e-4) and abs(np.trace(P).real - 3) < 1e-3
    assert np.allclose(A.astype(complex) @ A.astype(complex), m, atol=1e-4 * np.trace(m).real)


def test_gate_conventions():
    # Rzz(θ) = exp(-iθ/2 ZZ) after the registry's θ→θ/2 (gate_definitions.jl:49-51)
    th = 0.37
    zz = np.kron(Z, Z)
    assert np.allclose(tq.gate_matrix("Rzz", 2, th), np.diag(np.exp(-0.5j * th * np.diag(zz))))
    assert np.allclose(tq.gate_matrix("rzz", 2, th), tq.gate_matrix("Rzz", 2, th))
    assert np.allclose(tq.gate_matrix("Rx", 1, th),
                       math.cos(th / 2) * np.eye(2) - 1j * math.sin(th / 2) * X)
    assert np.allclose(tq.gate_matrix("cp", 2, th), np.diag([1, 1, 1, np.exp(1j * th)]))
    assert np.allclose(tq.gate_matrix("XZ", 2), np.kron(X, Z))
    with pytest.raises(tq.ArgumentError):
        tq.gate_matrix("Rzx", 2, 0.1)
    with pytest.raises(tq.ArgumentError):
        tq.register_gate("Rx", lambda t: np.eye(2), 1, 1)
    with pytest.raises(tq.ArgumentError):
        tq.gate_matrix("xx_plus_yy", 2, 0.1)
    tq.register_gate("MyZRot", lambda t: tq.gate_matrix("Rz", 1, t), 1, 1)
    assert np.allclose(tq.gate_matrix("MyZRot", 1, 0.3), tq.gate_matrix("Rz", 1, 0.3))
    tq.unregister_gate("MyZRot")
    with pytest.raises(tq.ArgumentError):
        tq.gate_matrix("MyZRot", 1, 0.3)


def test_truncate_reduces_bond_dimension_and_keeps_fidelity():
    """/root/reference/test/test_truncate.jl:29-33 flavour: after `truncate(bpc; maxdim)` every bond is
    ≤ maxdim and the fidelity with the untruncated state is in [0, 1] (and close to 1 for a mild cut)."""
    import tnqs_b200 as tq
    from helpers_cpu import tfim_layer_cpu
    g = tq.named_grid((3, 2))
    layer, gm, gv = tfim_layer_cpu(g)
    seq = [(g.index[a], g.index[b]) for a, b in tq.bipartite_edge_sequence(g)]
    c = orc.product_state(g.nv, g.edge_uv(), [(1.0, 0.0)] * g.nv, np.complex128)
    for _ in range(3):
        c, _, _ = orc.apply_gates(c, gm, gv, seq, dict(maxdim=8, cutoff=1e-14), dict(maxiter=100, tolerance=1e-12))
    assert c.maxvirtualdim() > 2
    groups = [[(g.index[a], g.index[b]) for a, b in grp] for grp in tq.edge_color(g, 3)]
    t = orc.truncate(c, groups, seq, maxdim=2, bp_update_kwargs=dict(maxiter=100, tolerance=1e-12))
    assert t.maxvirtualdim() <= 2
    p1, p2 = orc.to_statevector(c), orc.to_statevector(t)
    f = abs(np.vdot(p1, p2)) ** 2 / (np.vdot(p1, p1).real * np.vdot(p2, p2).real)
    assert 0.0 <= f <= 1.0 + 1e-12
    assert f > 0.9
    same = orc.truncate(c, groups, seq, maxdim=64, bp_update_kwargs=dict(maxiter=100, tolerance=1e-12))
    p3 = orc.to_statevector(same)
    f3 = abs(np.vdot(p1, p3)) ** 2 / (np.vdot(p1, p1).real * np.vdot(p3, p3).real)
    assert abs(f3 - 1) < 1e-10  # nothing to cut: an identity gate through the simple update changes nothing


def test_partitionfunction_exact_on_tree_and_rescale():
    """/root/reference/test/test_beliefpropagation.jl:24-29 (Z_bp = exact on a tree) and the rescale! contract
    (abstractbeliefpropagationcache.jl:318-322: afterwards every vertex and edge scalar is 1)."""
    g = tq.named_comb_tree((3, 2))
    c = orc.random_state(g.nv, g.edge_uv(), 2, 3, np.complex128, seed=5)
    seq = [(g.index[a], g.index[b]) for a, b in tq.forest_cover_edge_sequence(g)]
    c, _ = orc.bp_update(c, seq, maxiter=1, tolerance=None)
    psi = orc.to_statevector(c)
    z = orc.partitionfunction(c)
    assert abs(z - np.vdot(psi, psi)) < 1e-10 * abs(z)
    r = orc.rescale(c)
    assert np.allclose([orc.vertex_scalar(r, v) for v in range(r.nv)], 1.0, atol=1e-12)
    assert np.allclose([orc.edge_scalar(r, u, v) for (u, v) in r.edges], 1.0, atol=1e-12)
    assert abs(orc.partitionfunction(r) - 1) < 1e-12
    p2 = orc.to_statevector(r)
    assert abs(np.vdot(p2, p2) - 1) < 1e-10  # exact on a tree: the rescaled state is normalised


def _ghz_tensors(g, dtype=np.complex128):
    """|0…0⟩ + |1…1⟩ as a bond-dimension-2 TNS: T_v[s, a, b, …] = 1 iff s = a = b = …
    (the ψ1 + ψ2 of /root/reference/test/test_constructors.jl:69-70)."""
    ts = []
    for i in range(g.nv):
        z = len(g.incident[i])
        t = np.zeros((2,) + (2,) * z, dtype=dtype)
        t[(0,) * (z + 1)] = 1
        t[(1,) * (z + 1)] = 1
        ts.append(t)
    return ts


@pytest.mark.parametrize("graph", ["path", "grid"])
def test_ghz_bond_entropy_is_log2(graph):
    """Known answer of /root/reference/test/test_constructors.jl:69-74:
    von_neumann_entanglement_entropy(ψGHZ, e; alg="bp") ≈ log 2."""
    g = tq.named_path_graph(4) if graph == "path" else tq.named_grid((3, 3))
    c = orc.OracleCache(g.nv, g.edge_uv(), _ghz_tensors(g), np.complex128)
    seq = [(g.index[a], g.index[b]) for a, b in tq.forest_cover_edge_sequence(g)]
    c, _ = orc.bp_update(c, seq, maxiter=50, tolerance=1e-12)
    u, v = g.edge_uv()[0]
    assert abs(orc.renyi_entropy(c, u, v, 1.0) - math.log(2)) < 1e-10
    assert abs(orc.renyi_entropy(c, u, v, 2.0) - math.log(2)) < 1e-10


def test_symmetric_gauge_keeps_state_and_symmetrises_messages():
    """symmetric_gauge (src/symmetric_gauge.jl:1-56): a pure gauge transformation — the physical state is
    unchanged, both messages of every edge become the same diagonal matrix, and they stay a BP fixed point."""
    g = tq.named_grid((3, 2))
    c = orc.random_state(g.nv, g.edge_uv(), 2, 3, np.complex128, seed=9)
    seq = [(g.index[a], g.index[b]) for a, b in tq.bipartite_edge_sequence(g)]
    c, rep = orc.bp_update(c, seq, maxiter=500, tolerance=1e-14)
    s = orc.symmetric_gauge(c)
    p1, p2 = orc.to_statevector(c), orc.to_statevector(s)
    ov = abs(np.vdot(p1, p2)) / (np.linalg.norm(p1) * np.linalg.norm(p2))
    assert abs(ov - 1) < 1e-9
    for (a, b) in s.edges:
        m1, m2 = s.message(a, b), s.message(b, a)
        assert np.allclose(m1, m2) and np.allclose(m1, np.diag(np.diag(m1)))
    s2, rep2 = orc.bp_update(s, seq, maxiter=1, tolerance=None)
    for (a, b) in s.edges:  # still the fixed point (up to normalisation)
        x, y = s2.message(a, b), s.message(a, b)
        assert np.allclose(x / np.trace(x), y / np.trace(y), atol=1e-7)


def test_oracle_reproduces_golden_config1():
    """tests/golden/config1_…json (minted by tests/golden/make_golden.py) pins the oracle against drift; layer 1
    has an analytic answer: ⟨Z⟩ = cos(2·hx·dt) = cos(0.5) (Rx on |↑⟩, Rz and Rzz commute with Z)."""
    import json
    import os
    sys_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = json.load(open(os.path.join(sys_path, "config1_5x5_tfim_maxdim4_c128.json")))
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(sys_path, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    now = mg.config1(nlayers=3)
    assert abs(gold["layers"][0]["sz_center"] - math.cos(0.5)) < 1e-12
    for a, b in zip(now["layers"], gold["layers"]):
        assert a["maxvirtualdim"] == b["maxvirtualdim"]
        assert np.allclose(a["sz_all"], b["sz_all"], atol=1e-10)
        assert np.allclose(a["trunc_err"], b["trunc_err"], rtol=1e-6, atol=1e-13)


def test_multi_site_expect_exact_on_tree():
    """expect(alg="bp") for non-adjacent vertices (src/expect.jl:59-82, Steiner-tree region) is exact on a tree:
    compared with the dense state vector (the reference's test_expect.jl:26-30 flavour)."""
    g = tq.named_comb_tree((3, 2))
    c = orc.random_state(g.nv, g.edge_uv(), 2, 2, np.complex128, seed=13)
    seq = [(g.index[a], g.index[b]) for a, b in tq.forest_cover_edge_sequence(g)]
    c, _ = orc.bp_update(c, seq, maxiter=1, tolerance=None)
    psi = orc.to_statevector(c).reshape(-1)
    nrm = np.vdot(psi, psi)
    Zm, Xm = np.diag([1.0, -1.0]).astype(complex), np.array([[0, 1], [1, 0]], dtype=complex)
    for (u, v, o1, o2) in ((0, g.nv - 1, Zm, Xm), (1, 4, Xm, Xm), (2, 3, Zm, Zm)):
        region = orc.steiner_path(c, u, v)
        assert region[0] == u and region[-1] == v
        got = orc.expect_region(c, region, {u: o1, v: o2})
        full = np.ones((1, 1), dtype=complex)
        for i in range(g.nv):
            full = np.kron(full, o1 if i == u else (o2 if i == v else np.eye(2)))
        want = np.vdot(psi, full @ psi) / nrm
        assert abs(got - want) < 1e-10, (u, v)
    # adjacent pair: agrees with the dedicated two-site routine
    (a, b) = g.edge_uv()[0]
    assert abs(orc.expect_region(c, [a, b], {a: Zm, b: Zm}) - orc.expect_two_site(c, a, b, Zm, Zm)) < 1e-12


def test_truncate_spectrum_follows_ndtensors_truncate():
    """NDTensors `truncate!` as summarised in SURVEY.md §8a: maxdim first (whatever mindim says), then the summed
    relative test, or the per-weight absolute test with an unscaled truncation error."""
    p = np.array([4.0, 2.0, 1.0, 0.5, 0.25, 0.125])
    tot = p.sum()
    assert orc.truncate_spectrum(p, 3, None) == (3, (0.5 + 0.25 + 0.125) / tot)
    assert orc.truncate_spectrum(p, 2, None, mindim=4) == (2, (1.0 + 0.5 + 0.25 + 0.125) / tot)  # maxdim wins
    n, e = orc.truncate_spectrum(p, None, 0.4 / tot)  # summed relative test: 0.125 + 0.25 = 0.375 ≤ 0.4 < 0.875
    assert n == 4 and abs(e - 0.375 / tot) < 1e-15
    assert orc.truncate_spectrum(p, None, 0.4 / tot, mindim=5)[0] == 5
    n, e = orc.truncate_spectrum(p, None, 0.5, use_absolute_cutoff=True)  # per-weight test, error unscaled
    assert n == 3 and abs(e - 0.875) < 1e-15
    n, e = orc.truncate_spectrum(p, None, 0.4, use_relative_cutoff=False)  # summed test against 1
    assert n == 4 and abs(e - 0.375) < 1e-15
    assert orc.truncate_spectrum(np.array([3.0]), 1, 1.0) == (1, 0.0)
