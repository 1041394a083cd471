"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same inputs.

Tolerances (stated per north_star): ComplexF64 states must agree with the complex128 oracle to
1e-9 relative or better; ComplexF32 states are compared with the complex128 oracle at 2e-4 on
short circuits (fp32 round-off of a different but equivalent operation order; the oracle run in
complex64 differs from its own complex128 run by the same order) and at 1e-5 on single kernels."""
import numpy as np
import pytest

import tnqs_b200 as tq
from oracle import tnqs_oracle as orc
from helpers import (X, Y, Z, circuit_for_oracle, oracle_from_bpc, oracle_from_tns, ragged_state,
                     random_psd_messages, rel, seq_idx, state_overlap, tfim_layer)

pytestmark = pytest.mark.gpu

DT = [(np.complex128, 1e-10), (np.complex64, 2e-5)]


def small_graphs():
    return [tq.named_grid((3, 3)), tq.named_comb_tree((3, 2)), tq.named_path_graph(2), tq.named_grid((2, 2, 2))]


@pytest.mark.parametrize("dtype,tol", DT)
def test_site_and_message_roundtrip(dtype, tol):
    g = tq.named_grid((3, 2))
    dims = [2, 3, 4, 5, 2, 3, 4][:g.ne]
    psi = ragged_state(g, dims, dtype, seed=3)
    bpc = tq.BeliefPropagationCache(psi)
    assert not bpc.messages()  # empty right after construction (test_beliefpropagation.jl:18)
    for v in g.vertices():
        assert np.array_equal(bpc.site(v), psi[v])
    assert list(bpc.bond_dims()) == dims
    ms = random_psd_messages(g, dims, dtype)
    bpc.setmessages(list(ms), list(ms.values()))  # test_beliefpropagation.jl:72-82
    got = bpc.messages()
    assert set(got) == set(ms)
    for e in ms:
        assert np.array_equal(got[e], ms[e])
    e0 = g.edges[0]
    assert np.array_equal(tq.BeliefPropagationCache(psi).message(e0), np.eye(dims[0], dtype=dtype))
    c2 = bpc.copy()
    assert np.array_equal(c2.site(g.vertices()[1]), psi[g.vertices()[1]])
    assert np.array_equal(c2.message(e0), ms[e0])


@pytest.mark.parametrize("dtype,tol", DT)
def test_expect_with_given_messages(dtype, tol):
    # exercises the mode-product kernel on every leg position and the Gram kernel (ragged dims)
    for g in small_graphs():
        dims = [2 + (3 * e) % 4 for e in range(g.ne)]
        psi = ragged_state(g, dims, dtype, seed=5)
        bpc = tq.BeliefPropagationCache(psi)
        ms = random_psd_messages(g, dims, dtype)
        bpc.setmessages(list(ms), list(ms.values()))
        c = oracle_from_tns(psi)
        for (a, b), m in ms.items():
            c.msg[(g.index[a], g.index[b])] = m
        obs = [("Z", [v]) for v in g.vertices()] + [("X", [v], 0.5) for v in g.vertices()]
        got = tq.expect(bpc, obs)
        want = [orc.expect_local(c, g.index[v], Z) for v in g.vertices()] + \
               [orc.expect_local(c, g.index[v], X, 0.5) for v in g.vertices()]
        assert rel(got, want) < 50 * tol
        a, b = g.edges[0]
        got2 = tq.expect(bpc, ("XY", [a, b]))
        want2 = orc.expect_two_site(c, g.index[a], g.index[b], X, Y)
        assert abs(got2 - want2) < 50 * tol


@pytest.mark.parametrize("dtype,tol", DT)
@pytest.mark.parametrize("schedule", ["forest", "bipartite"])
def test_bp_update_matches_oracle(dtype, tol, schedule):
    for g in small_graphs():
        dims = [2 + (e % 3) for e in range(g.ne)]
        psi = ragged_state(g, dims, dtype, seed=7)
        seq = tq.forest_cover_edge_sequence(g) if schedule == "forest" else tq.bipartite_edge_sequence(g)
        bpc = tq.BeliefPropagationCache(psi)
        out = tq.update(bpc, maxiter=3, edge_sequence=seq)
        assert not bpc.messages()  # input not mutated (abstractbeliefpropagationcache.jl:228)
        c = oracle_from_tns(psi)
        c, _ = orc.bp_update(c, seq_idx(g, seq), maxiter=3, tolerance=None)
        got = out.messages()
        assert len(got) == 2 * g.ne
        for (a, b), m in got.items():
            assert rel(m, c.msg[(g.index[a], g.index[b])]) < 100 * tol, (schedule, a, b)
        # convergence loop + report
        out2 = tq.update(bpc, maxiter=200, tolerance=1e-7 if dtype == np.complex64 else 1e-13, edge_sequence=seq)
        c2, rep = orc.bp_update(oracle_from_tns(psi), seq_idx(g, seq), maxiter=200,
                                tolerance=1e-7 if dtype == np.complex64 else 1e-13)
        assert out2.last_bp_report["converged"]
        assert abs(out2.last_bp_report["niter"] - rep["niter"]) <= (2 if dtype == np.complex64 else 0)


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_bp_exact_on_tree(dtype):
    # /root/reference/test/test_beliefpropagation.jl:31-55
    g = tq.named_comb_tree((3, 3))
    psi = tq.random_tensornetworkstate(dtype, g, bond_dimension=2, seed=123)
    bpc = tq.update(tq.BeliefPropagationCache(psi))
    assert len(bpc.messages()) == 2 * g.ne
    assert bpc.last_bp_report["niter"] == 1
    c = oracle_from_tns(psi)
    full = orc.to_statevector(c)
    vc = g.center()[0]
    i = g.index[vc]
    rho = np.moveaxis(full, i, 0).reshape(2, -1)
    rho = rho @ rho.conj().T
    rho /= np.trace(rho)
    eps = np.finfo(np.float32 if dtype == np.complex64 else np.float64).eps
    for name, op in (("Z", Z), ("X", X), ("Y", Y)):
        assert abs(tq.expect(bpc, (name, [vc])) - np.trace(op @ rho)) <= 20 * eps


@pytest.mark.parametrize("dtype,tol", DT)
def test_one_site_gates(dtype, tol):
    g = tq.named_grid((2, 3))
    dims = [2, 3, 2, 4, 3, 2, 3][:g.ne]
    psi = ragged_state(g, dims, dtype, seed=9)
    circ = [("Rx", [v], 0.3 + 0.1 * i) for i, v in enumerate(g.vertices())] + [("H", [g.vertices()[0]]), ("Ry", g.vertices()[1], 0.7)]
    for norm in (False, True):
        bpc = tq.BeliefPropagationCache(psi)
        out, errs = tq.apply_gates(circ, bpc, apply_kwargs=dict(normalize_tensors=norm), update_cache=False)
        assert np.all(errs == 0)
        c = oracle_from_tns(psi)
        gm, gv = circuit_for_oracle(g, circ)
        c, _, _ = orc.apply_gates(c, gm, gv, [], dict(normalize_tensors=norm), update_cache=False)
        for i, v in enumerate(g.vertices()):
            assert rel(out.site(v), c.T[i]) < 20 * tol
        assert np.array_equal(bpc.site(g.vertices()[0]), psi[g.vertices()[0]])  # input untouched


@pytest.mark.parametrize("dtype,tol", DT)
@pytest.mark.parametrize("gate", ["Rzz", "CNOT", "Rxxyy", "SWAP"])
def test_single_two_site_gate(dtype, tol, gate):
    # one simple update with non-trivial environments; compare gauge-invariant outputs
    g = tq.named_grid((3, 2))
    dims = [2, 3, 2, 3, 2, 3, 2][:g.ne]
    psi = ragged_state(g, dims, dtype, seed=11)
    ms = random_psd_messages(g, dims, dtype, seed=12)
    for e_id in (0, 3, g.ne - 1):
        a, b = g.edges[e_id]
        for (maxdim, cutoff) in ((None, None), (3, 1e-12), (2, None)):
            bpc = tq.BeliefPropagationCache(psi)
            bpc.setmessages(list(ms), list(ms.values()))
            circ = [(gate, [a, b], 0.37)] if gate.startswith("R") else [(gate, [a, b])]
            kw = dict(normalize_tensors=True)
            if maxdim:
                kw["maxdim"] = maxdim
            if cutoff:
                kw["cutoff"] = cutoff
            out, errs = tq.apply_gates(circ, bpc, apply_kwargs=kw, update_cache=False)
            c = oracle_from_tns(psi)
            for (x, y), m in ms.items():
                c.msg[(g.index[x], g.index[y])] = m
            gm, gv = circuit_for_oracle(g, circ)
            c, oerrs, _ = orc.apply_gates(c, gm, gv, [], kw, update_cache=False)
            assert out.bond_dims()[e_id] == c.bond_dims()[e_id]
            assert abs(errs[0] - oerrs[0]) <= 200 * tol * max(oerrs[0], 1e-3)
            assert rel(np.diag(out.message((a, b))), np.diag(c.msg[(g.index[a], g.index[b])])) < 100 * tol
            assert np.array_equal(out.message((a, b)), out.message((b, a)))
            ov, n1, n2 = state_overlap(oracle_from_bpc(out), c)
            assert abs(ov - 1) < 100 * tol and abs(n1 / n2 - 1) < 100 * tol


@pytest.mark.parametrize("dtype,tol", DT)
@pytest.mark.parametrize("kw", [dict(maxdim=3, cutoff=1e-3, use_absolute_cutoff=True),
                                dict(cutoff=3e-2, use_relative_cutoff=False),
                                dict(cutoff=1e-3, use_absolute_cutoff=True, mindim=3),
                                dict(maxdim=2, mindim=3, cutoff=1e-12),
                                dict(maxdim=3, cutoff=1e-12, alg="qr_iteration")])
def test_factorize_svd_keywords(dtype, tol, kw):
    """The keywords `simple_update` forwards to `factorize_svd` (simple_update.jl:53-59): `use_absolute_cutoff`,
    `use_relative_cutoff`, `mindim` (maxdim wins: NDTensors `truncate!` drops to maxdim before it looks at mindim)
    and `alg`, against the oracle's restatement of NDTensors `truncate!`."""
    g = tq.named_grid((3, 2))
    dims = [2, 3, 2, 3, 2, 3, 2][:g.ne]
    psi = ragged_state(g, dims, dtype, seed=61)
    ms = random_psd_messages(g, dims, dtype, seed=62)
    a, b = g.edges[3]
    bpc = tq.BeliefPropagationCache(psi)
    bpc.setmessages(list(ms), list(ms.values()))
    kw = dict(kw, normalize_tensors=True)
    out, errs = tq.apply_gates([("Rxxyy", [a, b], 0.9)], bpc, apply_kwargs=kw, update_cache=False)
    c = oracle_from_tns(psi)
    for (x, y), m in ms.items():
        c.msg[(g.index[x], g.index[y])] = m
    gm, gv = circuit_for_oracle(g, [("Rxxyy", [a, b], 0.9)])
    c, oerrs, _ = orc.apply_gates(c, gm, gv, [], kw, update_cache=False)
    assert out.bond_dims()[3] == c.bond_dims()[3]
    assert abs(errs[0] - oerrs[0]) <= 200 * tol * max(oerrs[0], 1e-3)
    assert rel(np.diag(out.message((a, b))), np.diag(c.msg[(g.index[a], g.index[b])])) < 100 * tol
    with pytest.raises(tq.ArgumentError):
        tq.apply_gates([("Rzz", [a, b], 0.1)], bpc, apply_kwargs=dict(alg="bogus"))
    with pytest.raises(tq.ArgumentError):
        tq.apply_gates([("Rzz", [a, b], 0.1)], bpc, apply_kwargs=dict(ortho="left"))


@pytest.mark.parametrize("dtype,tol", DT)
def test_tfim_layers_small_grid(dtype, tol):
    # test_apply.jl:23-53 flavour with truncation: per-gate truncerr, ⟨Z⟩, bond dims, messages
    g = tq.named_grid((3, 3))
    layer = tfim_layer(g)
    seq = tq.bipartite_edge_sequence(g)
    bptol = 1e-13 if dtype == np.complex128 else 1e-10
    kw = dict(maxdim=3, cutoff=1e-12, normalize_tensors=True)
    bp = dict(maxiter=300, tolerance=bptol, edge_sequence=seq)
    psi = tq.BeliefPropagationCache(tq.zerostate(dtype, g))
    c = orc.product_state(g.nv, g.edge_uv(), [(1.0, 0.0)] * g.nv, np.complex128)
    gm, gv = circuit_for_oracle(g, layer)
    ftol = tol if dtype == np.complex128 else 10 * tol
    for layer_i in range(4):
        psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp)
        c, oerrs, reps = orc.apply_gates(c, gm, gv, seq_idx(g, seq), kw, dict(maxiter=300, tolerance=1e-13))
        assert len(psi.last_bp_reports) == len(reps) == 5
        assert list(psi.bond_dims()) == c.bond_dims()
        assert np.max(np.abs(errs - oerrs)) <= 100 * ftol * max(np.max(oerrs), 1e-6), layer_i
        zs = tq.expect(psi, [("Z", [v]) for v in g.vertices()])
        zo = [orc.expect_local(c, i, Z) for i in range(g.nv)]
        assert np.max(np.abs(np.array(zs) - np.array(zo))) < 100 * ftol, layer_i
    assert psi.maxvirtualdim() <= 3


def test_example_2d_ising_dynamics_plumbing_config():
    """BASELINE config 1: examples/2dIsing_dynamics.jl (5×5, dt=.25, hx=1, hz=.8, J=.5) with
    maxdim=4 / ComplexF64 per BASELINE.json, default BP kwargs, 6 layers, vs the oracle."""
    g = tq.named_grid((5, 5))
    layer = tfim_layer(g)
    kw = dict(maxdim=4, cutoff=1e-10, normalize_tensors=False)
    psi = tq.BeliefPropagationCache(tq.tensornetworkstate(np.complex128, lambda v: "↑", g, "S=1/2"))
    c = orc.product_state(g.nv, g.edge_uv(), [(1.0, 0.0)] * g.nv, np.complex128)
    seq = tq.forest_cover_edge_sequence(g)
    gm, gv = circuit_for_oracle(g, layer)
    for l in range(6):
        psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw)
        c, oerrs, _ = orc.apply_gates(c, gm, gv, seq_idx(g, seq), kw)
        sz = tq.expect(psi, ("Z", [(3, 3)]))
        so = orc.expect_local(c, g.index[(3, 3)], Z)
        assert psi.maxvirtualdim() == c.maxvirtualdim() <= 4
        assert abs(sz - so) < 1e-7, (l, sz, so)
        assert np.max(np.abs(errs - oerrs)) < 1e-7 * max(1.0, np.max(oerrs) / 1e-3), l


def test_error_conventions():
    g = tq.named_grid((2, 2))
    psi = tq.BeliefPropagationCache(tq.zerostate(np.complex64, g))
    with pytest.raises(tq.TnqsError) as ei:  # apply_gates.jl:114-120
        tq.apply_gates([("Rzz", [(1, 1), (2, 2)], 0.1)], psi)
    assert ei.value.code == 2
    with pytest.raises(tq.ArgumentError):    # gate_definitions.jl:131-140
        tq.apply_gates([("Rqq", [(1, 1), (2, 1)], 0.1)], psi)
    assert tq.expect(psi, ("Z", [(1, 1)], 0.0)) == 0
    assert abs(tq.expect(psi, ("Z", [(1, 1)])) - 1) < 1e-6


def test_oversized_factorisation_is_refused_up_front():
    # θ of a two-site gate may have at most 512 rows (d²·χ ≤ 512): refused with EINVAL before anything is
    # touched, not in the middle of the call after earlier gates of the same call were committed
    g = tq.named_path_graph(3)
    a, b, c = g.vertices()
    psi = tq.BeliefPropagationCache(ragged_state(g, [129, 300], np.complex64, seed=2))
    before = psi.site(b).copy()
    with pytest.raises(tq.TnqsError) as ei:
        tq.apply_gates([("Rzz", [b, c], 0.1), ("Rzz", [a, b], 0.1)], psi, apply_kwargs=dict(maxdim=300), inplace=True)
    assert "512" in str(ei.value)
    assert np.array_equal(psi.site(b), before) and list(psi.bond_dims()) == [129, 300]


def test_two_qubit_circuit_invariants():
    # /root/reference/test/test_apply.jl:11-20
    circuit = [("Rx", [(1, 1)], 0.5), ("Rx", [(2, 1)], 0.2), ("CPHASE", [(1, 1), (2, 1)], -0.3)]
    g = tq.build_graph_from_circuit(circuit)
    psi0 = tq.tensornetworkstate(np.complex64, lambda v: "↓", g)
    psi, _ = tq.apply_circuit(circuit, psi0, apply_kwargs=dict(maxdim=2, cutoff=1e-10, normalize_tensors=False))
    assert isinstance(psi, tq.TensorNetworkState) and psi.scalartype() == np.complex64
    assert psi.maxvirtualdim() <= 2
    full = orc.to_statevector(oracle_from_tns(psi))
    assert abs(np.vdot(full, full) - 1) < 1e-5


@pytest.mark.parametrize("chi", [8, 12, 6, 16])
def test_tensor_core_mode_products(chi):
    """ComplexF32 mode products on the tcgen05 path (3×TF32): every leg position (MN-major and
    K-major operand variants), padded shapes (χ not a multiple of 8), against the complex128 oracle.
    Tolerance 2e-5 relative: 3×TF32 carries ~2^-21 per product, fp32 accumulation in TMEM."""
    g = tq.named_grid((3, 3))
    dims = [chi] * g.ne
    psi = ragged_state(g, dims, np.complex64, seed=21)
    bpc = tq.BeliefPropagationCache(psi)
    ms = random_psd_messages(g, dims, np.complex64, seed=22)
    bpc.setmessages(list(ms), list(ms.values()))
    c = oracle_from_tns(psi)
    c.dtype = np.dtype(np.complex128)
    c.T = [t.astype(np.complex128) for t in c.T]
    for (a, b), m in ms.items():
        c.msg[(g.index[a], g.index[b])] = m.astype(np.complex128)
    bpc.stats(reset=True)
    got = tq.expect(bpc, [("Z", [v]) for v in g.vertices()] + [("X", [v]) for v in g.vertices()])
    want = [orc.expect_local(c, i, Z) for i in range(g.nv)] + [orc.expect_local(c, i, X) for i in range(g.nv)]
    assert np.max(np.abs(np.array(got) - np.array(want))) < 2e-5
    if chi >= 8:
        assert bpc.stats()["tc_launches"] > 0  # the tensor-core kernels really ran (tiny shapes stay on SIMT)
    seq = tq.bipartite_edge_sequence(g)
    out = tq.update(bpc, maxiter=2, edge_sequence=seq)
    c2, _ = orc.bp_update(c, seq_idx(g, seq), maxiter=2, tolerance=None)
    for (a, b), m in out.messages().items():
        assert rel(m, c2.msg[(g.index[a], g.index[b])]) < 5e-5, (a, b)
    # one two-site gate (final plane-mixing product on the tensor cores as well)
    a, b = (2, 2), (2, 3)
    kw = dict(maxdim=chi, cutoff=1e-12, normalize_tensors=True)
    out2, errs = tq.apply_gates([("Rzz", [a, b], 0.4)], bpc, apply_kwargs=kw, update_cache=False)
    gm, gv = circuit_for_oracle(g, [("Rzz", [a, b], 0.4)])
    c3, oerrs, _ = orc.apply_gates(c, gm, gv, [], kw, update_cache=False)
    assert abs(errs[0] - oerrs[0]) < 2e-4 * max(oerrs[0], 1e-3)
    assert rel(np.diag(out2.message((a, b))), np.diag(c3.msg[(g.index[a], g.index[b])])) < 2e-4
    zs = tq.expect(out2, [("Z", [a]), ("Z", [b])])
    zo = [orc.expect_local(c3, g.index[a], Z), orc.expect_local(c3, g.index[b], Z)]
    assert np.max(np.abs(np.array(zs) - np.array(zo))) < 1e-4


def _evolve(dtype, g, layer, seq, kw, nlayers, env):
    """Run `nlayers` layers with the engine's kernel-selection environment variables set to `env`
    (they are read when a cache is created)."""
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        psi = tq.BeliefPropagationCache(tq.zerostate(dtype, g))
        bp = dict(maxiter=100, tolerance=1e-12 if dtype == np.complex128 else 1e-9, edge_sequence=seq)
        errs_all = []
        for _ in range(nlayers):
            psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp)
            errs_all.append(errs)
        zs = np.array(tq.expect(psi, [("Z", [v]) for v in g.vertices()]))
        return zs, np.concatenate(errs_all), list(psi.bond_dims())
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-9), (np.complex64, 2e-4)])
@pytest.mark.parametrize("variant", [{"TNQS_FAST_SVD": "0"}, {"TNQS_CHOL": "0"}, {"TNQS_DMMA": "0"},
                                     {"TNQS_CLUSTER_JACOBI": "0", "TNQS_FAST_SVD": "0", "TNQS_CHOL": "0"}])
def test_kernel_variants_agree(dtype, tol, variant):
    """The accelerated factorisation paths (Cholesky-preconditioned θ-SVD with Jacobi polish,
    Cholesky-preconditioned eigendecomposition of the reduced-factor Gram, fp64 tensor-core Gram,
    shared-memory cluster Jacobi) against the plain ones they replace: same truncation errors, bond
    dimensions and ⟨Z⟩ on a truncating 4×4 TFIM evolution (bond dimension up to 8 ⇒ θ is 32×32)."""
    g = tq.named_grid((4, 4))
    layer = tfim_layer(g)
    seq = tq.bipartite_edge_sequence(g)
    kw = dict(maxdim=8, cutoff=1e-12, normalize_tensors=True)
    base = {"TNQS_FAST_SVD": "1", "TNQS_CHOL": "1", "TNQS_DMMA": "1", "TNQS_CLUSTER_JACOBI": "1"}
    z0, e0, b0 = _evolve(dtype, g, layer, seq, kw, 5, base)
    z1, e1, b1 = _evolve(dtype, g, layer, seq, kw, 5, {**base, **variant})
    assert b0 == b1
    assert max(b0) == 8
    assert np.max(np.abs(z0 - z1)) < tol
    assert np.max(np.abs(e0 - e1)) < tol * max(1.0, float(np.max(e1)) / 1e-3)


@pytest.mark.parametrize("name", ["heavy_hex", "cubic_periodic"])
def test_other_lattices_match_oracle(name):
    """BASELINE configs 3 and 4 at oracle-sized bond dimension: the IBM-Eagle heavy-hex graph
    (degrees 1–3, 3 edge colours, kicked-Ising layer of examples/heavyhexIsing_dynamics.jl:12-26) and
    a periodic cubic lattice (degree 6, layer of examples/3dIsing_dynamics.jl:15-26), complex128."""
    if name == "heavy_hex":
        g = tq.eagle_heavy_hex()
        layer = [("Rx", [v], 0.4) for v in g.vertices()]
        for grp in tq.edge_color(g, 3):
            layer += [("Rzz", list(p), np.pi / 2) for p in grp]
        kw = dict(maxdim=4, cutoff=1e-12, normalize_tensors=True)
        nl, probe = 3, g.vertices()[:12]
    else:
        g = tq.named_grid((2, 2, 3), periodic=False)
        layer = [("Rz", [v], -0.04) for v in g.vertices()]
        for grp in tq.edge_color(g, 6):
            layer += [("Rxx", list(p), -0.08) for p in grp]
        layer += [("Rz", [v], -0.04) for v in g.vertices()]
        kw = dict(maxdim=2, cutoff=1e-10, normalize_tensors=True)
        nl, probe = 2, g.vertices()
    seq = tq.bipartite_edge_sequence(g)
    bp = dict(maxiter=200, tolerance=1e-13, edge_sequence=seq)
    psi = tq.BeliefPropagationCache(tq.zerostate(np.complex128, g))
    c = orc.product_state(g.nv, g.edge_uv(), [(1.0, 0.0)] * g.nv, np.complex128)
    gm, gv = circuit_for_oracle(g, layer)
    for _ in range(nl):
        psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp)
        c, oerrs, _ = orc.apply_gates(c, gm, gv, seq_idx(g, seq), kw, dict(maxiter=200, tolerance=1e-13))
        assert list(psi.bond_dims()) == c.bond_dims()
        assert np.max(np.abs(errs - oerrs)) < 1e-9 * max(1.0, float(np.max(oerrs)) / 1e-3)
    zs = np.array(tq.expect(psi, [("Z", [v]) for v in probe]))
    zo = np.array([orc.expect_local(c, g.index[v], Z) for v in probe])
    assert np.max(np.abs(zs - zo)) < 1e-8


@pytest.mark.parametrize("dtype,tol", DT)
def test_truncate_matches_oracle(dtype, tol):
    """`truncate(bpc; maxdim)` (src/truncate.jl:12-30; reference test test_truncate.jl:29-33): identity gates
    through the batched simple update per colour group + BP update, against the oracle with the same groups."""
    g = tq.named_grid((3, 3))
    layer = tfim_layer(g)
    seq = tq.bipartite_edge_sequence(g)
    bptol = 1e-13 if dtype == np.complex128 else 1e-10
    bp = dict(maxiter=300, tolerance=bptol, edge_sequence=seq)
    psi = tq.BeliefPropagationCache(tq.zerostate(dtype, g))
    for _ in range(3):
        psi, _ = tq.apply_gates(layer, psi, apply_kwargs=dict(maxdim=6, cutoff=1e-12), bp_update_kwargs=bp)
    assert psi.maxvirtualdim() > 3
    c = oracle_from_bpc(psi)
    groups = tq.edge_color(g, 4)
    out = tq.truncate(psi, maxdim=3, bp_update_kwargs=bp, edge_groups=groups)
    assert psi.maxvirtualdim() > 3 and out.maxvirtualdim() <= 3  # functional copy, bound respected
    ogroups = [[(g.index[a], g.index[b]) for a, b in grp] for grp in groups]
    co = orc.truncate(c, ogroups, seq_idx(g, seq), maxdim=3, bp_update_kwargs=dict(maxiter=300, tolerance=1e-13))
    assert list(out.bond_dims()) == co.bond_dims()
    ftol = tol if dtype == np.complex128 else 10 * tol
    zs = np.array(tq.expect(out, [("Z", [v]) for v in g.vertices()]))
    zo = np.array([orc.expect_local(co, i, Z) for i in range(g.nv)])
    assert np.max(np.abs(zs - zo)) < 100 * ftol
    ov, n1, n2 = state_overlap(oracle_from_bpc(out), co)
    assert abs(ov - 1) < 100 * ftol


@pytest.mark.parametrize("dtype,tol", DT)
def test_partitionfunction_rescale_normalize(dtype, tol):
    """vertex/edge scalars, freenergy / partitionfunction (abstractbeliefpropagationcache.jl:22-28,289-304,
    beliefpropagationcache.jl:47-49), rescale (:82-101,127-140) and normalize(alg="bp") (normalize.jl:1-6)
    against the oracle on a loopy graph, and the exactness on a tree of test_beliefpropagation.jl:24-29."""
    ftol = tol if dtype == np.complex128 else 10 * tol
    g = tq.named_grid((3, 3))
    dims = [2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3][:g.ne]
    psi = ragged_state(g, dims, dtype, seed=31)
    seq = tq.bipartite_edge_sequence(g)
    bpc = tq.update(tq.BeliefPropagationCache(psi), maxiter=200, tolerance=1e-13 if dtype == np.complex128 else 1e-9,
                    edge_sequence=seq)
    c = oracle_from_bpc(bpc)
    vs = tq.vertex_scalars(bpc)
    vo = np.array([orc.vertex_scalar(c, i) for i in range(g.nv)])
    assert rel(vs, vo) < 100 * ftol
    es = tq.edge_scalars(bpc)
    eo = np.array([orc.edge_scalar(c, u, v) for (u, v) in c.edges])
    assert rel(es, eo) < 100 * ftol
    z, zo = tq.partitionfunction(bpc), orc.partitionfunction(c)
    assert abs(z - zo) < 100 * ftol * abs(zo)
    assert abs(tq.norm_sqr(bpc, alg="bp") - z) < 1e-12 * abs(z)
    r = tq.rescale(bpc)
    assert abs(tq.partitionfunction(bpc) - z) < 1e-12 * abs(z)  # functional copy: the input is untouched
    assert np.max(np.abs(tq.vertex_scalars(r) - 1)) < 100 * ftol
    assert np.max(np.abs(tq.edge_scalars(r) - 1)) < 100 * ftol
    ro = orc.rescale(c)
    for i, v in enumerate(g.vertices()):
        assert rel(r.site(v), ro.T[i]) < 100 * ftol
    # tree: BP is exact, so normalize(alg="bp") returns a state of unit norm
    gt = tq.named_comb_tree((3, 2))
    pt = ragged_state(gt, [2, 3, 2, 3, 2][:gt.ne], dtype, seed=32)
    nt = tq.normalize(pt, alg="bp")
    full = orc.to_statevector(oracle_from_tns(nt))
    assert abs(np.vdot(full, full) - 1) < 100 * ftol
    full0 = orc.to_statevector(oracle_from_tns(pt))
    assert abs(tq.norm_sqr(pt, alg="bp") - np.vdot(full0, full0)) < 100 * ftol * abs(np.vdot(full0, full0))


@pytest.mark.parametrize("dtype,tol", DT)
def test_bond_entropy(dtype, tol):
    """renyi_entropy(bp_cache, e; α) (src/entanglement.jl:73-86): the GHZ known answer log 2 of
    /root/reference/test/test_constructors.jl:69-74 on the device path, and a random loopy state vs the oracle."""
    g = tq.named_grid((3, 3))
    ts = {}
    for i, v in enumerate(g.vertices()):
        z = len(g.incident[i])
        t = np.zeros((2,) + (2,) * z, dtype=dtype)
        t[(0,) * (z + 1)] = 1
        t[(1,) * (z + 1)] = 1
        ts[v] = t
    ghz = tq.TensorNetworkState(g, ts, dtype)
    assert ghz.maxvirtualdim() == 2
    e0 = g.edges[0]
    s = tq.von_neumann_entanglement_entropy(ghz, e0, alg="bp")
    assert abs(s - np.log(2)) < (1e-9 if dtype == np.complex128 else 1e-4)
    dims = [2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3][:g.ne]
    psi = ragged_state(g, dims, dtype, seed=41)
    seq = tq.bipartite_edge_sequence(g)
    bpc = tq.update(tq.BeliefPropagationCache(psi), maxiter=200, tolerance=1e-13 if dtype == np.complex128 else 1e-9,
                    edge_sequence=seq)
    c = oracle_from_bpc(bpc)
    for e in g.edges[:4]:
        for alpha in (1.0, 2.0):
            got = tq.renyi_entropy(bpc, e, alpha)
            want = orc.renyi_entropy(c, g.index[e[0]], g.index[e[1]], alpha)
            assert abs(got - want) < (1e-9 if dtype == np.complex128 else 1e-3), (e, alpha)


@pytest.mark.parametrize("dtype,tol", DT)
def test_symmetric_gauge(dtype, tol):
    """symmetric_gauge (src/symmetric_gauge.jl:1-56) on the device path: the state is unchanged, both messages of
    every bond become the same diagonal matrix, equal to the oracle's, and remain the BP fixed point."""
    ftol = tol if dtype == np.complex128 else 10 * tol
    g = tq.named_grid((3, 2))
    dims = [2, 3, 2, 3, 2, 3, 2][:g.ne]
    psi = ragged_state(g, dims, dtype, seed=51)
    seq = tq.bipartite_edge_sequence(g)
    bpc = tq.update(tq.BeliefPropagationCache(psi), maxiter=500, tolerance=1e-14 if dtype == np.complex128 else 1e-10,
                    edge_sequence=seq)
    c = oracle_from_bpc(bpc)
    sg = tq.symmetric_gauge(bpc)
    so = orc.symmetric_gauge(c)
    ov, n1, n2 = state_overlap(oracle_from_bpc(sg), c)
    assert abs(ov - 1) < 100 * ftol and abs(n1 / n2 - 1) < 100 * ftol
    for (a, b) in g.edges:
        m1, m2 = sg.message((a, b)), sg.message((b, a))
        assert np.array_equal(m1, m2)
        assert rel(np.diag(m1), np.diag(so.msg[(g.index[a], g.index[b])])) < 100 * ftol
    again = tq.update(sg, maxiter=1, tolerance=None, edge_sequence=seq)
    for (a, b) in g.edges:
        x, y = again.message((a, b)).astype(np.complex128), sg.message((a, b)).astype(np.complex128)
        assert rel(x / np.trace(x), y / np.trace(y)) < (1e-6 if dtype == np.complex128 else 1e-3)


def test_device_matches_golden_config1():
    """BASELINE config 1 against the committed golden vectors (tests/golden/, minted by the oracle): per-gate
    truncation errors, bond dimension and ⟨Z⟩ on every vertex for 4 layers; layer 1 also has the analytic
    value ⟨Z⟩ = cos(0.5)."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                       "config1_5x5_tfim_maxdim4_c128.json")))
    g = tq.named_grid((5, 5))
    layer = tfim_layer(g)
    kw = dict(maxdim=4, cutoff=1e-10, normalize_tensors=False)
    psi = tq.BeliefPropagationCache(tq.tensornetworkstate(np.complex128, lambda v: "↑", g, "S=1/2"))
    for l, ref in enumerate(gold["layers"]):
        psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw)
        assert psi.maxvirtualdim() == ref["maxvirtualdim"]
        zs = np.real(np.array(tq.expect(psi, [("Z", [v]) for v in g.vertices()])))
        assert np.max(np.abs(zs - np.array(ref["sz_all"]))) < 1e-7, l
        assert np.max(np.abs(errs - np.array(ref["trunc_err"]))) < 1e-7 * max(1.0, ref["max_trunc_err"] / 1e-3), l
        if l == 0:
            assert abs(zs[g.index[(3, 3)]] - np.cos(0.5)) < 1e-12


@pytest.mark.parametrize("dtype,tol", DT)
def test_multi_site_expect_and_rdm(dtype, tol):
    """Multi-site BP `expect` over a Steiner path (src/expect.jl:67-81) and `reduced_density_matrix(alg="bp")`
    (src/rdm.jl:52-73) on the device path (`tnqs_site_contract`), against the oracle's dense region contraction on a
    loopy graph and against the exact state vector on a tree (BP is exact there, test_beliefpropagation.jl:44-54)."""
    ftol = tol if dtype == np.complex128 else 10 * tol
    # loopy graph, converged BP messages
    g = tq.named_grid((3, 3))
    dims = [2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3][:g.ne]
    psi = ragged_state(g, dims, dtype, seed=71)
    seq = tq.bipartite_edge_sequence(g)
    bpc = tq.update(tq.BeliefPropagationCache(psi), maxiter=300, tolerance=1e-13 if dtype == np.complex128 else 1e-9, edge_sequence=seq)
    c = oracle_from_bpc(bpc)
    ix = g.index
    cases = [("ZZ", [(1, 1), (3, 1)]), ("XY", [(1, 1), (3, 3)], 0.5), ("ZXZ", [(1, 2), (2, 2), (3, 2)]), ("YZ", [(2, 1), (2, 3)])]
    PA = {"X": X, "Y": Y, "Z": Z}
    for obs in cases:
        got = tq.expect(bpc, obs)
        path = tq.steiner_path(g, obs[1])
        ops = {ix[v]: PA[o] for v, o in zip(obs[1], obs[0])}
        want = orc.expect_region(c, path, ops, obs[2] if len(obs) > 2 else 1.0)
        assert abs(got - want) < 100 * ftol, (obs, got, want)
    for vs in ([(2, 2)], [(1, 1), (3, 1)], [(1, 3), (3, 3)], [(1, 1), (1, 2)]):
        got = tq.reduced_density_matrix(bpc, vs)
        want = orc.rdm_region(c, [ix[v] for v in vs])
        assert abs(np.trace(got) - 1) < 1e-12
        assert np.max(np.abs(got - want)) < 100 * ftol, vs
    # tree: BP reduced density matrices are exact
    gt = tq.named_comb_tree((3, 2))
    pt = ragged_state(gt, [2, 3, 2, 3, 2][:gt.ne], dtype, seed=72)
    bt = tq.update(tq.BeliefPropagationCache(pt))
    full = orc.to_statevector(oracle_from_tns(pt))
    va, vb = gt.vertices()[0], gt.vertices()[-1]
    a, b = gt.index[va], gt.index[vb]
    f = np.moveaxis(full, (a, b), (0, 1)).reshape(4, -1)
    ex = f @ f.conj().T
    ex /= np.trace(ex)
    assert np.max(np.abs(tq.reduced_density_matrix(bt, [va, vb]) - ex)) < 100 * ftol
    zz = np.trace(np.kron(Z, Z) @ ex)
    assert abs(tq.expect(bt, ("ZZ", [va, vb])) - zz) < 100 * ftol
