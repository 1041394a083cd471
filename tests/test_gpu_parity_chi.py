"""GPU parity at the BENCHMARKED shapes (BASELINE configs 2 and the χ=64 target): random χ=32 TNS on a 4×4
patch (interior vertices of full degree 4: θ is 128×128, site tensors 16.8 MB — the kernel branches
`bench.py` runs), one colour group of Rzz + one BP sweep + ⟨Z⟩/⟨ZZ⟩, and one interior χ=64 gate
(θ 256×256, site tensors 268 MB), against the complex128 oracle ON THE SAME INPUTS.

Tolerances are the north star's: ComplexF32 ≤ 1e-5 relative on per-gate truncation error, singular
values and expectation values (the oracle is evaluated in complex128 on the upcast ComplexF32 inputs
with the ComplexF32 `sqrt_cutoff`, so the comparison measures the device's arithmetic, not the
oracle's own fp32 round-off); ComplexF64 ≤ 1e-9.  The measured deviations are printed (pytest -s)."""
import json
import os

import numpy as np
import pytest

import tnqs_b200 as tq
from oracle import tnqs_oracle as orc
from helpers import X, Z, circuit_for_oracle, seq_idx

pytestmark = pytest.mark.gpu

EPS32 = float(np.finfo(np.float32).eps)
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.jsonl")


def _report(**kw):
    print("PARITY", json.dumps(kw))
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def _random_state(g, chi, dtype, seed):
    rng = np.random.default_rng(seed)
    ts = {}
    for i, v in enumerate(g.vertices()):
        shp = (2,) + (chi,) * len(g.incident[i])
        t = rng.standard_normal(shp, dtype=np.float32) + 1j * rng.standard_normal(shp, dtype=np.float32)
        ts[v] = (t / np.linalg.norm(t)).astype(dtype)
    return tq.TensorNetworkState(g, ts, dtype)


def _psd_messages(g, chi, dtype, seed, null_dirs=0):
    """Random PSD messages, sum-normalised like BP's (bench.py CPU sample uses the same construction).
    `null_dirs` > 0 plants that many eigenvalues at 1e-12 of the trace: they sit far below the absolute
    pseudo-inverse cutoff 10·eps(Float32) on both sides, so the projector P ≠ 1 branch runs."""
    rng = np.random.default_rng(seed)
    out = {}
    for (a, b) in g.edges:
        for edge in ((a, b), (b, a)):
            w = rng.standard_normal((chi, chi)) + 1j * rng.standard_normal((chi, chi))
            q, _ = np.linalg.qr(w)
            lam = rng.uniform(0.2, 1.0, chi)
            if null_dirs:
                lam[:null_dirs] = 1e-12
            m = (q * lam) @ q.conj().T
            m = 0.5 * (m + m.conj().T)
            out[edge] = (m / np.trace(m).real).astype(dtype)
    return out


def _oracle(g, psi, ms):
    """complex128 oracle cache holding exactly the (possibly ComplexF32) values the device received."""
    c = orc.OracleCache(g.nv, g.edge_uv(), [np.asarray(psi.tensors[v], dtype=np.complex128) for v in g.vertices()],
                        np.complex128)
    for (a, b), m in ms.items():
        c.msg[(g.index[a], g.index[b])] = np.asarray(m, dtype=np.complex128)
    return c


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


@pytest.mark.parametrize("dtype,tol", [(np.complex64, 1e-5), (np.complex128, 1e-9)])
@pytest.mark.parametrize("null_dirs", [0, 5])
def test_chi32_patch_colour_bp_expect(dtype, tol, null_dirs):
    chi = 32
    g = tq.named_grid((4, 4))
    psi = _random_state(g, chi, dtype, seed=101)
    ms = _psd_messages(g, chi, dtype, seed=102, null_dirs=null_dirs)
    bpc = tq.BeliefPropagationCache(psi)
    bpc.setmessages(list(ms), list(ms.values()))
    c = _oracle(g, psi, ms)
    # the colour group that holds the interior-interior edge (2,2)-(3,2): 8 gates, two of them between
    # degree-4 sites (θ 128×128), the others corner / boundary sites (mixed shapes in one batch)
    groups = tq.edge_color(g, 4)
    grp = next(gr for gr in groups if any(set(p) == {(2, 2), (3, 2)} for p in gr))
    assert len(grp) == 8
    circ = [("Rzz", list(p), 0.7) for p in grp]
    kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True, sqrt_cutoff=10 * EPS32)
    bpc.stats(reset=True)
    out, errs = tq.apply_gates(circ, bpc, apply_kwargs=kw, update_cache=False)
    st = out.stats()
    gm, gv = circuit_for_oracle(g, circ)
    c, oerrs, _ = orc.apply_gates(c, gm, gv, [], kw, update_cache=False)
    assert list(out.bond_dims()) == c.bond_dims()
    d_err = float(np.max(np.abs(errs - oerrs) / np.maximum(oerrs, 1e-3)))
    d_sig = max(_rel(np.diag(out.message(tuple(p))), np.diag(c.msg[(g.index[p[0]], g.index[p[1]])])) for p in grp)
    # expectation values right after the gates (new tensors + diag(σ) messages + old outer messages)
    obs = [("Z", [v]) for p in grp for v in p] + [("X", [v]) for p in grp for v in p]
    got = np.array(tq.expect(out, obs))
    want = np.array([orc.expect_local(c, g.index[v], Z) for p in grp for v in p] +
                    [orc.expect_local(c, g.index[v], X) for p in grp for v in p])
    d_exp = float(np.max(np.abs(got - want)))
    got2 = np.array(tq.expect(out, [("ZZ", list(p)) for p in grp]))
    want2 = np.array([orc.expect_two_site(c, g.index[p[0]], g.index[p[1]], Z, Z) for p in grp])
    d_zz = float(np.max(np.abs(got2 - want2)))
    # one BP sweep (bipartite schedule, two fully parallel levels), then ⟨Z⟩ everywhere
    seq = tq.bipartite_edge_sequence(g)
    out2 = tq.update(out, maxiter=1, tolerance=None, edge_sequence=seq)
    c2, _ = orc.bp_update(c, seq_idx(g, seq), maxiter=1, tolerance=None)
    gate_edges = {frozenset(p) for p in grp}
    d_msg = 0.0
    for (a, b), m in out2.messages().items():
        mo = c2.msg[(g.index[a], g.index[b])]
        if frozenset((a, b)) in gate_edges:
            # the new bond's basis is fixed only up to a phase per singular vector: m[c,c'] carries D_c·conj(D_c'), and
            # the reference's normalisation by sum(m) (abstractbeliefpropagationcache.jl:182-187) is not phase invariant
            # either, so compare |m| / tr m
            m, mo = m.astype(np.complex128), mo.astype(np.complex128)
            d_msg = max(d_msg, _rel(np.abs(m / np.trace(m)), np.abs(mo / np.trace(mo))))
        else:
            d_msg = max(d_msg, _rel(m, mo))
    zs = np.array(tq.expect(out2, [("Z", [v]) for v in g.vertices()]))
    zo = np.array([orc.expect_local(c2, i, Z) for i in range(g.nv)])
    d_z = float(np.max(np.abs(zs - zo)))
    _report(test="chi32_patch", dtype=np.dtype(dtype).name, null_dirs=null_dirs, truncerr_rel=d_err, sigma_rel=d_sig,
            expect_abs=d_exp, zz_abs=d_zz, bp_message_rel=d_msg, z_after_bp_abs=d_z, max_truncerr=float(np.max(oerrs)),
            tc_launches=int(st["tc_launches"]), tma_launches=int(st["tma_launches"]), kernel_launches=int(st["kernel_launches"]))
    assert d_err <= tol and d_sig <= tol and d_exp <= tol and d_zz <= tol and d_z <= tol
    assert d_msg <= 5 * tol
    if dtype == np.complex64:
        assert st["tc_launches"] > 0  # tcgen05 mode products / final plane-mixing product really ran
        assert st["tma_launches"] > 0  # … through the TMA-fed warp-specialised kernel (kernels_tc2.cuh)


def _hub_graph():
    """Two degree-4 hubs joined by an edge, three leaves each: the smallest graph with an interior
    square-lattice gate (both sites z = 4)."""
    vs = ["a", "b"] + [f"a{i}" for i in range(3)] + [f"b{i}" for i in range(3)]
    es = [("a", "b")] + [("a", f"a{i}") for i in range(3)] + [("b", f"b{i}") for i in range(3)]
    return tq.NamedGraph(vs, es)


@pytest.mark.parametrize("chi,dtype,tol", [(64, np.complex64, 1e-5), (32, np.complex128, 1e-9)])
def test_interior_gate_hub_graph(chi, dtype, tol):
    """One interior two-site gate at the target bond dimension χ=64 (ComplexF32: site tensors 268 MB, θ 256×256,
    Gram matrices 128×128) and its BP message update, against the complex128 oracle."""
    g = _hub_graph()
    psi = _random_state(g, chi, dtype, seed=201)
    ms = _psd_messages(g, chi, dtype, seed=202)
    bpc = tq.BeliefPropagationCache(psi)
    bpc.setmessages(list(ms), list(ms.values()))
    c = _oracle(g, psi, ms)
    circ = [("Rzz", ["a", "b"], 0.7)]
    kw = dict(maxdim=chi, cutoff=1e-10, normalize_tensors=True, sqrt_cutoff=10 * EPS32)
    out, errs = tq.apply_gates(circ, bpc, apply_kwargs=kw, update_cache=False)
    st = out.stats()
    gm, gv = circuit_for_oracle(g, circ)
    c, oerrs, _ = orc.apply_gates(c, gm, gv, [], kw, update_cache=False)
    assert list(out.bond_dims()) == c.bond_dims()
    d_err = abs(errs[0] - oerrs[0]) / max(oerrs[0], 1e-3)
    d_sig = _rel(np.diag(out.message(("a", "b"))), np.diag(c.msg[(0, 1)]))
    got = np.array(tq.expect(out, [("Z", ["a"]), ("Z", ["b"]), ("X", ["a"]), ("X", ["b"])]))
    want = np.array([orc.expect_local(c, 0, Z), orc.expect_local(c, 1, Z), orc.expect_local(c, 0, X), orc.expect_local(c, 1, X)])
    d_exp = float(np.max(np.abs(got - want)))
    d_zz = 0.0
    if chi <= 32:  # the oracle's dense two-site contraction takes ~30 s at χ=64
        zz = tq.expect(out, ("ZZ", ["a", "b"]))
        d_zz = abs(zz - orc.expect_two_site(c, 0, 1, Z, Z))
    # messages leaving the hubs towards the leaves: one BP update of those edges
    seq = [("a", f"a{i}") for i in range(3)] + [("b", f"b{i}") for i in range(3)]
    out2 = tq.update(out, maxiter=1, tolerance=None, edge_sequence=seq)
    c2, _ = orc.bp_update(c, seq_idx(g, seq), maxiter=1, tolerance=None)
    d_msg = max(_rel(out2.message(e), c2.msg[(g.index[e[0]], g.index[e[1]])]) for e in seq)
    _report(test="hub_gate", chi=chi, dtype=np.dtype(dtype).name, truncerr_rel=float(d_err), sigma_rel=d_sig, expect_abs=d_exp,
            zz_abs=float(d_zz), bp_message_rel=d_msg, truncerr=float(oerrs[0]), tc_launches=int(st["tc_launches"]),
            tma_launches=int(st["tma_launches"]))
    assert d_err <= tol and d_sig <= tol and d_exp <= tol and d_zz <= tol and d_msg <= 5 * tol
