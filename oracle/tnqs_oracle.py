"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module, and only as the checker / reported CPU baseline.

A NumPy (OpenBLAS/LAPACK: heevd, geqrf, gesdd — the library family Julia's LinearAlgebra uses)
restatement of the reference's belief-propagation simple-update path.  Each function cites the
reference file:line it follows (paths under /root/reference).

PARITY UNPINNED: the reference is pure Julia on un-vendored ITensors.jl / NamedGraphs.jl and
neither `julia` nor those sources exist in this image, and the reference's tests hold no numeric
golden vectors for this path.  The oracle is therefore pinned only to (i) the
reference's own test *invariants* (tests/test_oracle.py), and (ii) an independent dense
state-vector simulator (oracle/statevector.py).  Un-vendored behaviour restated from published
ITensors/NDTensors semantics: thin QR, SVD + `truncate!` on σ² (maxdim first, then relative cutoff,
mindim=1), `factorize_svd(...; ortho="none")` = (U√S, √S V†), Hermitian `eigen` full spectrum.

Data model (identical to the device layout so that tests compare arrays directly):
  * graph: `nv`, `edges[e] = (u, v)` 0-based; `incident[v] = [(e, w), ...]` in increasing edge id
  * site tensor `T[v]` has shape `(d, χ_leg0, χ_leg1, ...)`, legs in `incident[v]` order
  * message `msg[(w, v)]` (directed w→v) has shape `(χ, χ)`, first index ket bond, second bra bond
    (`default_message` = δ, src/TensorNetworks/tensornetworkstate.jl:72-75); absent ⇒ identity
    (src/MessagePassing/abstractbeliefpropagationcache.jl:99-102).
"""
from __future__ import annotations

import copy as _copy
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def _real_dtype(dtype):
    return np.float32 if np.dtype(dtype) in (np.dtype(np.complex64), np.dtype(np.float32)) else np.float64


def _wide(dtype):
    return np.complex128 if np.dtype(dtype).kind == "c" else np.float64


class OracleCache:
    """State + BP messages: the oracle's `BeliefPropagationCache`
    (src/MessagePassing/beliefpropagationcache.jl:9-15,27-31)."""

    def __init__(self, nv: int, edges: Sequence[Tuple[int, int]], tensors: List[np.ndarray],
                 dtype=np.complex128):
        self.nv = nv
        self.edges = [tuple(e) for e in edges]
        self.dtype = np.dtype(dtype)
        self.incident: List[List[Tuple[int, int]]] = [[] for _ in range(nv)]
        for e, (u, v) in enumerate(self.edges):
            self.incident[u].append((e, v))
            self.incident[v].append((e, u))
        self.T = [np.ascontiguousarray(t, dtype=self.dtype) for t in tensors]
        for v in range(nv):
            assert self.T[v].ndim == 1 + len(self.incident[v]), (v, self.T[v].shape)
        self.msg: Dict[Tuple[int, int], np.ndarray] = {}

    def copy(self) -> "OracleCache":
        """`Base.copy` (beliefpropagationcache.jl:35-37); deep here, which is observationally the
        same because nothing mutates tensors in place."""
        c = _copy.copy(self)
        c.T = [t.copy() for t in self.T]
        c.msg = {k: m.copy() for k, m in self.msg.items()}
        return c

    # -- helpers ------------------------------------------------------------------------------
    def leg(self, v: int, w: int) -> int:
        """axis (≥1) of T[v] that is the bond to neighbour w."""
        for pos, (_, x) in enumerate(self.incident[v]):
            if x == w:
                return pos + 1
        raise KeyError((v, w))

    def has_edge(self, u: int, v: int) -> bool:
        return any(x == v for _, x in self.incident[u])

    def bond_dims(self) -> List[int]:
        return [self.T[u].shape[self.leg(u, v)] for (u, v) in self.edges]

    def maxvirtualdim(self) -> int:
        """src/TensorNetworks/abstracttensornetwork.jl:27-29."""
        return max(self.bond_dims(), default=1)

    def message(self, w: int, v: int) -> np.ndarray:
        """`message` with identity default (abstractbeliefpropagationcache.jl:99-102)."""
        m = self.msg.get((w, v))
        if m is None:
            chi = self.T[v].shape[self.leg(v, w)]
            m = np.eye(chi, dtype=self.dtype)
        return m

    def is_tree(self) -> bool:
        return len(self.edges) == self.nv - 1

    def directed_edges(self) -> List[Tuple[int, int]]:
        out = []
        for (u, v) in self.edges:
            out += [(u, v), (v, u)]
        return out


# ---------------------------------------------------------------------------------------------
# constructors (src/TensorNetworks/tensornetworkstate.jl:93-103,141-161)
# ---------------------------------------------------------------------------------------------

def product_state(nv, edges, local_states, dtype=np.complex128) -> OracleCache:
    """`tensornetworkstate(eltype, v -> state, g)`: T_v = state ⊗ onehot on χ=1 bonds."""
    deg = [0] * nv
    for u, v in edges:
        deg[u] += 1
        deg[v] += 1
    tensors = [np.asarray(local_states[v], dtype=dtype).reshape((-1,) + (1,) * deg[v]) for v in range(nv)]
    return OracleCache(nv, edges, tensors, dtype)


def random_state(nv, edges, d, chi, dtype=np.complex128, seed=1234) -> OracleCache:
    """`random_tensornetworkstate`: iid normal entries, all bonds = chi.  The RNG stream is ours
    (Julia's `Random.seed!` stream is not reproducible here)."""
    rng = np.random.default_rng(seed)
    deg = [0] * nv
    for u, v in edges:
        deg[u] += 1
        deg[v] += 1
    tensors = []
    for v in range(nv):
        shp = (d,) + (chi,) * deg[v]
        if np.dtype(dtype).kind == "c":
            t = rng.standard_normal(shp) + 1j * rng.standard_normal(shp)
        else:
            t = rng.standard_normal(shp)
        tensors.append(t.astype(dtype))
    return OracleCache(nv, edges, tensors, dtype)


# ---------------------------------------------------------------------------------------------
# belief propagation (src/MessagePassing/abstractbeliefpropagationcache.jl:150-259,
#                     src/MessagePassing/beliefpropagationcache.jl:17-21,52-80,103-119)
# ---------------------------------------------------------------------------------------------

def _absorb(t: np.ndarray, m: np.ndarray, axis: int) -> np.ndarray:
    """Y[.., a', ..] = Σ_a t[.., a, ..] m[a, a'] on `axis` (a BLAS GEMM via tensordot)."""
    y = np.tensordot(t, m, axes=([axis], [0]))
    return np.moveaxis(y, -1, axis)


def updated_message(c: OracleCache, u: int, v: int, normalize: bool = True) -> np.ndarray:
    """m_{u→v}[l,l'] = Σ T_u · ∏_{w≠v} m_{w→u} · conj(T_u), then divided by the sum of all
    entries unless that sum is exactly zero (abstractbeliefpropagationcache.jl:162-190; factors
    from tensornetworkstate.jl:50-70)."""
    t = c.T[u]
    y = t
    for _, w in c.incident[u]:
        if w == v:
            continue
        y = _absorb(y, c.message(w, u), c.leg(u, w))
    ax = c.leg(u, v)
    ym = np.moveaxis(y, ax, 0).reshape(y.shape[ax], -1)
    tm = np.moveaxis(t, ax, 0).reshape(t.shape[ax], -1)
    m = ym @ tm.conj().T
    if normalize:
        s = m.sum()
        if s != 0:
            m = m / s
    return m.astype(c.dtype)


def message_diff(a: np.ndarray, b: np.ndarray) -> float:
    """1 − |⟨a,b⟩|²/(‖a‖²‖b‖²) (beliefpropagationcache.jl:17-21)."""
    na, nb = np.linalg.norm(a), np.linalg.norm(b)
    f = abs(np.vdot(a, b) / (na * nb)) ** 2
    return float(1.0 - f)


def default_bp_update_kwargs(c: OracleCache) -> dict:
    """beliefpropagationcache.jl:103-119: tree → one sweep, no tolerance; loopy → 25 sweeps and
    1e-5 (32-bit) / 1e-8 (64-bit)."""
    if c.is_tree():
        return dict(maxiter=1, tolerance=None)
    tol = 1e-5 if _real_dtype(c.dtype) == np.float32 else 1e-8
    return dict(maxiter=25, tolerance=tol)


def bp_update(c: OracleCache, edge_sequence: Sequence[Tuple[int, int]], maxiter: int,
              tolerance: Optional[float]) -> Tuple[OracleCache, dict]:
    """`update(alg"bp")` (abstractbeliefpropagationcache.jl:223-255): up to `maxiter` sequential
    (Gauss–Seidel) sweeps over `edge_sequence`; converged when the mean per-edge `message_diff`
    of a sweep is ≤ tolerance.  Returns a copy plus a report."""
    c = c.copy()
    niter, avg = maxiter, None
    converged = False
    for it in range(1, maxiter + 1):
        diff = 0.0
        for (u, v) in edge_sequence:  # update_iteration! :204-218
            prev = c.message(u, v)
            new = updated_message(c, u, v)
            c.msg[(u, v)] = new
            if tolerance is not None:
                diff += message_diff(new, prev)
        if tolerance is not None:
            avg = diff / len(edge_sequence)
            if avg <= tolerance:
                converged, niter = True, it
                break
    return c, dict(niter=niter, diff=avg, converged=converged or tolerance is None)


# ---------------------------------------------------------------------------------------------
# simple update (src/Apply/simple_update.jl:21-77, src/utils.jl:18-35,94-108)
# ---------------------------------------------------------------------------------------------

def pseudo_sqrt_inv_sqrt(m: np.ndarray, cutoff: float) -> Tuple[np.ndarray, np.ndarray]:
    """Hermitian eigendecomposition in Float64/ComplexF64 (`safe_eigen`, utils.jl:94-108), then
    f(λ)=0 if λ==0 or |λ|<cutoff else √λ, g(λ)=0 or 1/√λ (utils.jl:18-26).  √ of a negative
    eigenvalue ≥ cutoff in magnitude is a DomainError in Julia; here a ValueError."""
    mw = np.asarray(m, dtype=_wide(m.dtype))
    lam, q = np.linalg.eigh(mw)  # uses the lower triangle, as LAPACK heev does
    keep = ~((lam == 0) | (np.abs(lam) < cutoff))
    if np.any(lam[keep] < 0):
        raise ValueError("DomainError: sqrt of a negative message eigenvalue")
    f = np.zeros_like(lam)
    g = np.zeros_like(lam)
    f[keep] = np.sqrt(lam[keep])
    g[keep] = 1.0 / f[keep]
    a = (q * f) @ q.conj().T
    b = (q * g) @ q.conj().T
    return a.astype(m.dtype), b.astype(m.dtype)


def truncate_spectrum(p: np.ndarray, maxdim: Optional[int], cutoff: Optional[float],
                      mindim: int = 1, use_absolute_cutoff: bool = False,
                      use_relative_cutoff: bool = True) -> Tuple[int, float]:
    """NDTensors `truncate!` on P = σ² sorted descending (called through `factorize_svd`,
    simple_update.jl:53-59; NDTensors is not vendored — behaviour as summarised in SURVEY.md §8a): drop
    from the tail while n > maxdim (whatever mindim says); then, while n > mindim, either the absolute
    test P[n] ≤ cutoff (truncerr left unscaled) or the summed test (discarded + P[n]) ≤ cutoff·scale with
    scale = ΣP when `use_relative_cutoff` (else 1) and truncerr /= scale.  A single candidate is never
    truncated.  Returns (kept n, truncerr)."""
    n = len(p)
    if n <= 1:
        return n, 0.0
    total = float(np.sum(p))
    disc = 0.0
    mindim = max(1, int(mindim))
    if maxdim is not None:
        while n > max(int(maxdim), 1):
            disc += float(p[n - 1])
            n -= 1
    if use_absolute_cutoff:
        if cutoff is not None:
            while n > mindim and float(p[n - 1]) <= cutoff:
                disc += float(p[n - 1])
                n -= 1
        return n, disc
    scale = (total if total > 0 else 1.0) if use_relative_cutoff else 1.0
    if cutoff is not None:
        while n > mindim and (disc + float(p[n - 1])) <= cutoff * scale:
            disc += float(p[n - 1])
            n -= 1
    return n, disc / scale


def simple_update_two_site(gate: np.ndarray, t1: np.ndarray, t2: np.ndarray, ax1: int, ax2: int,
                           envs1: Dict[int, np.ndarray], envs2: Dict[int, np.ndarray],
                           maxdim=None, cutoff=None, normalize_tensors=True, sqrt_cutoff=None,
                           mindim=1, use_absolute_cutoff=False, use_relative_cutoff=True, alg="divide_and_conquer"):
    """Two-site branch of `simple_update` (simple_update.jl:29-68).  `t1`,`t2` are the site
    tensors (d, legs…); `ax1`/`ax2` the axis of the shared bond in each; `envs_i` maps the axis of
    every *other* bond leg of site i to its incoming message.  `gate` is the d²×d² matrix in
    kron(site1, site2) order.  Returns (T1', T2', σ, truncerr)."""
    dtype = t1.dtype
    if sqrt_cutoff is None:  # :32-33
        sqrt_cutoff = 10 * np.finfo(_real_dtype(dtype)).eps
    out, rs, qs, inv_list = [], [], [], []
    for (t, ax, envs) in ((t1, ax1, envs1), (t2, ax2, envs2)):
        d = t.shape[0]
        sq = {a: pseudo_sqrt_inv_sqrt(m, sqrt_cutoff) for a, m in envs.items()}  # :38-41
        tt = t
        for a, (A, _) in sq.items():  # :43-44  ψ̃ = T × ∏ √M
            tt = _absorb(tt, A, a)
        # QR with rows = external bonds, columns = (s, shared bond)  :47-48
        ext_axes = [a for a in range(1, t.ndim) if a != ax]
        perm = ext_axes + [0, ax]
        mat = np.transpose(tt, perm)
        ext_shape = mat.shape[:-2]
        chi_b = t.shape[ax]
        mat = mat.reshape(int(np.prod(ext_shape, dtype=np.int64)), d * chi_b)
        q, r = np.linalg.qr(mat, mode="reduced")
        qs.append((q.reshape(ext_shape + (q.shape[1],)), ext_axes, ax))
        rs.append(r.reshape(r.shape[0], d, chi_b))
        inv_list.append({a: B for a, (_, B) in sq.items()})
    r1, r2 = rs
    d1, d2 = r1.shape[1], r2.shape[1]
    theta = np.einsum("asb,ctb->asct", r1, r2)  # :51  R1*R2 over the shared bond
    g4 = np.asarray(gate, dtype=dtype).reshape(d1, d2, d1, d2)  # [s1',s2',s1,s2]
    theta = np.einsum("xyst,asct->axcy", g4, theta)
    n1, n2 = r1.shape[0], r2.shape[0]
    mat = theta.reshape(n1 * d1, n2 * d2)
    if alg == "qr_iteration":  # `alg` of factorize_svd: LAPACK gesvd
        import scipy.linalg
        u, s, vh = scipy.linalg.svd(mat, full_matrices=False, lapack_driver="gesvd")
    elif alg in ("divide_and_conquer", "recursive"):
        try:
            u, s, vh = np.linalg.svd(mat, full_matrices=False)  # gesdd  :53-59
        except np.linalg.LinAlgError:  # NDTensors falls back to qr_iteration
            import scipy.linalg
            u, s, vh = scipy.linalg.svd(mat, full_matrices=False, lapack_driver="gesvd")
    else:
        raise ValueError(f"unknown SVD algorithm {alg!r}")
    keep, err = truncate_spectrum((s.astype(np.float64)) ** 2, maxdim, cutoff, mindim, use_absolute_cutoff, use_relative_cutoff)
    u, s, vh = u[:, :keep], s[:keep], vh[:keep, :]
    rs_ = np.sqrt(s).astype(s.dtype)
    f1 = (u * rs_).reshape(n1, d1, keep)                      # U√S      [r1,s1,c]
    f2 = (vh * rs_[:, None]).T.reshape(n2, d2, keep)          # (√S V†)ᵀ [r2,s2,c]
    for (q, ext_axes, ax), inv, f in zip(qs, inv_list, (f1, f2)):
        # un-gauge :62-63  Q[.., a, ..] = Σ_{a'} Q[.., a', ..] conj(B)[a, a']
        for pos, a in enumerate(ext_axes):
            q = _absorb(q, inv[a].conj().T, pos)
        tnew = np.tensordot(q, f, axes=([q.ndim - 1], [0]))    # :64  [ext…, s, c]
        # back to (s, legs in incident order) with the new bond at `ax`
        nd = tnew.ndim
        order = [nd - 2]
        ext_iter = iter(range(nd - 2))
        for a in range(1, nd):
            order.append(nd - 1 if a == ax else next(ext_iter))
        tnew = np.transpose(tnew, order)
        out.append(np.ascontiguousarray(tnew))
    sv = s.astype(np.float64)
    if normalize_tensors:  # :65-74
        sv = sv / np.linalg.norm(sv)
        out = [t / np.linalg.norm(t) for t in out]
    out = [np.ascontiguousarray(t, dtype=dtype) for t in out]
    return out[0], out[1], sv, float(err)


def apply_gate(c: OracleCache, gate: np.ndarray, verts: Sequence[int], maxdim=None, cutoff=None,
               normalize_tensors=True, sqrt_cutoff=None, mindim=1, use_absolute_cutoff=False,
               use_relative_cutoff=True, alg="divide_and_conquer") -> float:
    """`apply_gate!` (src/Apply/apply_gates.jl:101-143), in place on `c`; returns truncerr."""
    nvs = len(verts)
    if not 1 <= nvs <= 2:
        raise RuntimeError("apply_gate!: only one- and two-site gates are supported; "
                           f"received a gate acting on {nvs} vertices: {list(verts)}.")
    if nvs == 1:
        v = verts[0]
        t = np.tensordot(np.asarray(gate, dtype=c.dtype), c.T[v], axes=([1], [0]))  # simple_update.jl:26-28
        if normalize_tensors:  # the normalisation at :70-74 also covers the one-site branch
            t = t / np.linalg.norm(t)
        c.T[v] = np.ascontiguousarray(t, dtype=c.dtype)
        return 0.0
    v1, v2 = verts
    if not c.has_edge(v1, v2):
        raise RuntimeError("apply_gate!: cannot apply a two-site gate on the non-adjacent vertices "
                           f"{v1} and {v2}.")
    ax1, ax2 = c.leg(v1, v2), c.leg(v2, v1)
    envs1 = {c.leg(v1, w): c.message(w, v1) for _, w in c.incident[v1] if w != v2}  # :122
    envs2 = {c.leg(v2, w): c.message(w, v2) for _, w in c.incident[v2] if w != v1}
    t1, t2, s, err = simple_update_two_site(gate, c.T[v1], c.T[v2], ax1, ax2, envs1, envs2,
                                            maxdim, cutoff, normalize_tensors, sqrt_cutoff, mindim,
                                            use_absolute_cutoff, use_relative_cutoff, alg)
    m = np.diag(s).astype(c.dtype)  # :126-136, σ ≥ 0 so the sign fix is the identity
    c.msg[(v1, v2)] = m.copy()
    c.msg[(v2, v1)] = m.copy()
    c.T[v1], c.T[v2] = t1, t2  # :138-140
    return err


def segment_gates(gate_verts: Sequence[Sequence[int]]) -> List[int]:
    """Indices `i` at which a BP refresh fires *before* gate i (apply_gates.jl:60-90): a two-site
    gate touching any vertex touched since the last refresh."""
    affected = set()
    fires = []
    for i, vs in enumerate(gate_verts):
        if len(vs) >= 2 and any(v in affected for v in vs):
            fires.append(i)
            affected.clear()
        affected.update(vs)
    return fires


def apply_gates(c: OracleCache, gates: Sequence[np.ndarray], gate_verts: Sequence[Sequence[int]],
                edge_sequence, apply_kwargs: Optional[dict] = None,
                bp_update_kwargs: Optional[dict] = None, update_cache: bool = True):
    """`apply_gates(circuit::Vector{<:ITensor}, bpc; …)` (apply_gates.jl:46-98).  Returns
    (new cache, truncation errors, list of BP reports)."""
    apply_kwargs = dict(apply_kwargs or {})
    if bp_update_kwargs is None:
        bp_update_kwargs = default_bp_update_kwargs(c)
    c = c.copy()
    errs = np.zeros(len(gates))
    affected = set()
    reports = []
    for i, (g, vs) in enumerate(zip(gates, gate_verts)):
        need = len(vs) >= 2 and any(v in affected for v in vs)
        if update_cache and need:
            c, rep = bp_update(c, edge_sequence, **bp_update_kwargs)
            reports.append(rep)
            affected.clear()
        errs[i] = apply_gate(c, g, list(vs), **apply_kwargs)
        affected.update(vs)
    if update_cache:
        c, rep = bp_update(c, edge_sequence, **bp_update_kwargs)
        reports.append(rep)
    return c, errs, reports


def truncate(c: OracleCache, edge_groups, edge_sequence, maxdim: int, cutoff=None,
             normalize_tensors: bool = True, bp_update_kwargs: Optional[dict] = None) -> OracleCache:
    """`truncate(bpc; maxdim, cutoff, edge_color=true)` (src/truncate.jl:12-30): per edge-colour group an
    identity two-site gate on every truncatable edge (`truncatable_edge`, :5-10: bond dimension > 1)
    through `apply_gate!`, then one BP `update`.  `edge_groups` = lists of (v1, v2) index pairs."""
    if bp_update_kwargs is None:
        bp_update_kwargs = default_bp_update_kwargs(c)
    c = c.copy()
    for grp in edge_groups:
        for (a, b) in grp:
            if c.T[a].shape[c.leg(a, b)] <= 1:
                continue
            d = c.T[a].shape[0] * c.T[b].shape[0]
            apply_gate(c, np.eye(d, dtype=c.dtype), [a, b], maxdim=maxdim, cutoff=cutoff,
                       normalize_tensors=normalize_tensors)
        c, _ = bp_update(c, edge_sequence, **bp_update_kwargs)
    return c


# ---------------------------------------------------------------------------------------------
# local expectation values (src/expect.jl:59-82, tensornetworkstate.jl:50-67)
# ---------------------------------------------------------------------------------------------

def rdm_local(c: OracleCache, v: int) -> np.ndarray:
    """Un-normalised one-site BP density matrix ρ[s,s'] = Σ T[s,a..] ∏ m[a,a'] conj(T[s',a'..])."""
    t = c.T[v]
    y = t
    for _, w in c.incident[v]:
        y = _absorb(y, c.message(w, v), c.leg(v, w))
    d = t.shape[0]
    return y.reshape(d, -1) @ t.reshape(d, -1).conj().T


def expect_local(c: OracleCache, v: int, op: np.ndarray, coeff=1.0):
    """⟨O_v⟩ = coeff·N(O)/N(1), N(X) = Σ X[s',s] ρ[s,s'] (expect.jl:59-82)."""
    if coeff == 0:
        return 0.0
    rho = rdm_local(c, v).astype(_wide(c.dtype))
    return coeff * np.sum(np.asarray(op) * rho.T) / np.trace(rho)


# ---------------------------------------------------------------------------------------------
# scalars of the BP fixed point, rescaling, norm
# ---------------------------------------------------------------------------------------------
def vertex_scalar(c: OracleCache, v: int):
    """`vertex_scalar` (src/MessagePassing/abstractbeliefpropagationcache.jl:22-28)."""
    return np.trace(rdm_local(c, v).astype(_wide(c.dtype)))


def edge_scalar(c: OracleCache, u: int, v: int):
    """`edge_scalar` (src/MessagePassing/beliefpropagationcache.jl:47-49): message(e)·message(reverse(e))."""
    return np.sum(c.message(u, v).astype(_wide(c.dtype)) * c.message(v, u).astype(_wide(c.dtype)))


def freenergy(c: OracleCache):
    """`freenergy` (abstractbeliefpropagationcache.jl:289-300): Σ log(vertex scalars) − Σ log(edge scalars)."""
    num = np.array([vertex_scalar(c, v) for v in range(c.nv)], dtype=complex)
    den = np.array([edge_scalar(c, u, v) for (u, v) in c.edges], dtype=complex)
    if np.any(den == 0):
        return -np.inf
    return np.sum(np.log(num)) - np.sum(np.log(den))


def partitionfunction(c: OracleCache):
    """`partitionfunction` (:302-304); for a TensorNetworkState cache this is `norm_sqr(alg="bp")` (norm_sqr.jl:72-78)."""
    return np.exp(freenergy(c))


def rescale(c: OracleCache) -> OracleCache:
    """`rescale` = `rescale_messages!` (beliefpropagationcache.jl:127-140) then `rescale_vertices!` (:82-101)."""
    c = c.copy()
    for (u, v) in c.edges:
        me, mer = c.message(u, v), c.message(v, u)
        me = me / np.linalg.norm(me)
        mer = mer / np.linalg.norm(mer)
        n = np.sum(me.astype(_wide(c.dtype)) * mer.astype(_wide(c.dtype)))
        if np.imag(n) == 0:
            sg = np.sign(np.real(n))
            me = me * sg
            n = n * sg
        c.msg[(u, v)] = (me / np.sqrt(n)).astype(c.dtype)
        c.msg[(v, u)] = (mer / np.sqrt(n)).astype(c.dtype)
    for v in range(c.nv):
        vn = vertex_scalar(c, v)
        sg = np.sign(np.real(vn)) if np.imag(vn) == 0 else 1.0
        c.T[v] = (c.T[v] * (sg / np.sqrt(vn))).astype(c.dtype)
    return c


def symmetric_gauge_factors(mx: np.ndarray, my: np.ndarray, regularization: float):
    """The χ×χ algebra of one edge of `symmetric_gauge!` (src/symmetric_gauge.jl:12-40): from
    mx = message(src→dst), my = message(dst→src) return (X_src, X_dst, S) with
    ψ_src ← ψ_src ×_e X_src, ψ_dst ← ψ_dst ×_e X_dst, both new messages = diag(S).
    X_src = X^{-1/2}·U·√S, X_dst = Y^{-1/2}·conj(V)·√S where X^{1/2}·(Y^{1/2})ᵀ = U·S·V†."""
    # ITensors' `eigen` reads a tensor on (l, l') as the map l → l', i.e. as the matrix m[bra, ket] = mᵀ of our
    # m[ket, bra] storage; with m itself the transformed messages would only be symmetric for real messages
    # (derivation in tests/test_oracle.py::test_symmetric_gauge_…: the fixed point must survive the regauging).
    xd, xu = np.linalg.eigh(np.asarray(mx, dtype=np.complex128).T)
    yd, yu = np.linalg.eigh(np.asarray(my, dtype=np.complex128).T)
    xd = xd + regularization
    yd = yd + regularization
    if np.any(xd < 0) or np.any(yd < 0):
        raise ValueError("DomainError: sqrt of a negative message eigenvalue")
    root_x = (xu * np.sqrt(xd)) @ xu.conj().T
    root_y = (yu * np.sqrt(yd)) @ yu.conj().T
    inv_root_x = (xu / np.sqrt(xd)) @ xu.conj().T
    inv_root_y = (yu / np.sqrt(yd)) @ yu.conj().T
    ce = root_x @ root_y.T
    u, sv, vh = np.linalg.svd(ce)
    v = vh.conj().T
    x_src = inv_root_x @ u * np.sqrt(sv)
    x_dst = inv_root_y @ v.conj() * np.sqrt(sv)
    return x_src, x_dst, sv


def symmetric_gauge(c: OracleCache, regularization: Optional[float] = None) -> OracleCache:
    """`symmetric_gauge(bp_cache)` (src/symmetric_gauge.jl:1-56), no SVD truncation keywords."""
    c = c.copy()
    if regularization is None:
        regularization = 10 * np.finfo(_real_dtype(c.dtype)).eps
    for (a, b) in c.edges:
        x_src, x_dst, sv = symmetric_gauge_factors(c.message(a, b), c.message(b, a), regularization)
        c.T[a] = np.ascontiguousarray(_absorb(c.T[a].astype(np.complex128), x_src, c.leg(a, b)), dtype=c.dtype)
        c.T[b] = np.ascontiguousarray(_absorb(c.T[b].astype(np.complex128), x_dst, c.leg(b, a)), dtype=c.dtype)
        c.msg[(a, b)] = np.diag(sv).astype(c.dtype)
        c.msg[(b, a)] = np.diag(sv).astype(c.dtype)
    return c


def renyi_entropy_matrix(rho: np.ndarray, alpha: float, normalize: bool = True) -> float:
    """`renyi_entropy(ρ::AbstractMatrix, α)` (src/entanglement.jl:21-29)."""
    rho = np.asarray(rho)
    if normalize:
        rho = rho / np.trace(rho)
    lam = np.linalg.eigvalsh(rho)
    eps = np.finfo(lam.dtype).eps
    lam = lam[np.abs(lam) > 10 * eps]
    if alpha == 1:
        return float(-np.sum(lam * np.log(lam)))
    return float(np.log(np.sum(lam ** alpha)) / (1 - alpha))


def renyi_entropy(c: OracleCache, u: int, v: int, alpha: float = 1.0) -> float:
    """`renyi_entropy(bp_cache, e; α)` (src/entanglement.jl:73-86): ρ = √m2ᵀ·m1·√m2ᵀ from the two messages
    of the bond (m1 = message(e), m2 = message(reverse(e)), √ via `pseudo_sqrt_inv_sqrt` with the default cutoff)."""
    m1, m2 = c.message(u, v), c.message(v, u)
    eps = np.finfo(_real_dtype(c.dtype)).eps
    r = pseudo_sqrt_inv_sqrt(m2, 10 * eps)[0].astype(_wide(c.dtype))
    rho = r.T @ m1.astype(_wide(c.dtype)) @ r.T
    return renyi_entropy_matrix(rho, alpha)


def expect_two_site(c: OracleCache, v1: int, v2: int, op1: np.ndarray, op2: np.ndarray, coeff=1.0):
    """Adjacent two-site observable: region {v1,v2} is its own Steiner tree (expect.jl:67)."""
    if coeff == 0:
        return 0.0
    assert c.has_edge(v1, v2)
    halves = []
    for (v, o) in ((v1, v2), (v2, v1)):
        t = c.T[v]
        y = t
        for _, w in c.incident[v]:
            if w != o:
                y = _absorb(y, c.message(w, v), c.leg(v, w))
        ax = c.leg(v, o)
        d, chi = t.shape[0], t.shape[ax]
        ym = np.moveaxis(y, (0, ax), (0, 1)).reshape(d, chi, -1)
        tm = np.moveaxis(t, (0, ax), (0, 1)).reshape(d, chi, -1)
        halves.append(np.einsum("sbx,tcx->sbtc", ym, tm.conj()).astype(_wide(c.dtype)))  # [s,b,s',b']
    e1, e2 = halves
    rho = np.einsum("sbtc,ubvc->sutv", e1, e2)  # [s1,s2,s1',s2']
    num = np.einsum("ts,vu,sutv->", np.asarray(op1), np.asarray(op2), rho)
    den = np.einsum("susu->", rho)
    return coeff * num / den


# ---------------------------------------------------------------------------------------------
# dense contraction for small systems (test helper: `norm_sqr(ψ; alg="exact")` etc.)
# ---------------------------------------------------------------------------------------------

def steiner_path(c: OracleCache, v1: int, v2: int) -> List[int]:
    """Vertices of a shortest path v1 … v2 (BFS, neighbours in incidence order).  For two vertices the Steiner
    tree of expect.jl:67 is a shortest path; it is unique on trees (on loopy graphs the reference's choice among
    equal-length paths is an implementation detail of Graphs.steiner_tree)."""
    prev = {v1: -1}
    queue = [v1]
    while queue:
        x = queue.pop(0)
        if x == v2:
            break
        for _, w in c.incident[x]:
            if w not in prev:
                prev[w] = x
                queue.append(w)
    path = [v2]
    while path[-1] != v1:
        path.append(prev[path[-1]])
    return path[::-1]


def expect_region(c: OracleCache, region: Sequence[int], ops: Dict[int, np.ndarray], coeff=1.0):
    """`expect(Algorithm"bp", cache, obs)` for a multi-site observable (src/expect.jl:59-82): the tensors of the
    Steiner-tree region, their conjugates, the operators and the messages entering the region, contracted
    exactly; numerator / denominator.  `region` must be connected; dense contraction (small regions only)."""
    if coeff == 0:
        return 0.0
    region = list(region)
    inside = set(region)
    wide = _wide(c.dtype)
    # region state: axes = [phys of every region vertex..., external legs...]; internal bonds contracted
    letters = iter("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ")
    bond_letter: Dict[int, str] = {}
    phys_letter: Dict[int, str] = {}
    ext: List[Tuple[str, int, int]] = []  # (letter, inside vertex, outside neighbour)
    subs = []
    for v in region:
        phys_letter[v] = next(letters)
        sub = phys_letter[v]
        for e, w in c.incident[v]:
            if w in inside:
                if e not in bond_letter:
                    bond_letter[e] = next(letters)
                sub += bond_letter[e]
            else:
                l = next(letters)
                ext.append((l, v, w))
                sub += l
        subs.append(sub)
    out_sub = "".join(phys_letter[v] for v in region) + "".join(l for l, _, _ in ext)
    psi = np.einsum(",".join(subs) + "->" + out_sub, *[c.T[v].astype(wide) for v in region], optimize=True)
    n = len(region)
    # absorb the incoming messages on the ket side of every external leg: m[ket, bra]
    ket = psi
    for k, (_, v, w) in enumerate(ext):
        ket = np.moveaxis(np.tensordot(ket, c.message(w, v).astype(wide), axes=([n + k], [0])), -1, n + k)
    dims = psi.shape[:n]
    kmat = ket.reshape(int(np.prod(dims)), -1)
    bmat = psi.reshape(int(np.prod(dims)), -1).conj()
    rho = kmat @ bmat.T  # ρ[s, s'] = Σ_ext ket[s, ext'] conj(ψ)[s', ext']
    op = np.ones((1, 1), dtype=wide)
    for v in region:
        op = np.kron(op, np.asarray(ops.get(v, np.eye(c.T[v].shape[0])), dtype=wide))
    return coeff * np.sum(op * rho.T) / np.trace(rho)


def rdm_region(c: OracleCache, vs: Sequence[int], normalize: bool = True) -> np.ndarray:
    """`reduced_density_matrix(Algorithm"bp", cache, vs)` (src/rdm.jl:52-73): tensors of the Steiner region of `vs`
    (one vertex, or the shortest path between two), their conjugates and the incoming messages contracted exactly with
    the physical legs of `vs` left open; ρ[(s_1 s_2), (s_1' s_2')], trace-normalised (`normalize_rdm`)."""
    vs = list(vs)
    region = vs if len(vs) == 1 else steiner_path(c, vs[0], vs[1])
    inside = set(region)
    wide = _wide(c.dtype)
    letters = iter("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ")
    bond_letter: Dict[int, str] = {}
    phys_letter: Dict[int, str] = {}
    ext: List[Tuple[str, int, int]] = []
    subs = []
    for v in region:
        phys_letter[v] = next(letters)
        sub = phys_letter[v]
        for e, w in c.incident[v]:
            if w in inside:
                if e not in bond_letter:
                    bond_letter[e] = next(letters)
                sub += bond_letter[e]
            else:
                l = next(letters)
                ext.append((l, v, w))
                sub += l
        subs.append(sub)
    # open physical legs of vs first (in the order given), then the traced ones, then the external bonds
    closed = [v for v in region if v not in vs]
    out_sub = "".join(phys_letter[v] for v in vs) + "".join(phys_letter[v] for v in closed) + "".join(l for l, _, _ in ext)
    psi = np.einsum(",".join(subs) + "->" + out_sub, *[c.T[v].astype(wide) for v in region], optimize=True)
    n = len(region)
    ket = psi
    for k, (_, v, w) in enumerate(ext):
        ket = np.moveaxis(np.tensordot(ket, c.message(w, v).astype(wide), axes=([n + k], [0])), -1, n + k)
    dopen = int(np.prod(psi.shape[:len(vs)]))
    rho = ket.reshape(dopen, -1) @ psi.reshape(dopen, -1).conj().T
    return rho / np.trace(rho) if normalize else rho


def to_statevector(c: OracleCache) -> np.ndarray:
    """Contract the whole TNS into a dense vector ψ[s_0, …, s_{nv-1}] (small systems only)."""
    letters = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    assert c.nv + len(c.edges) <= len(letters)
    site = {v: letters[v] for v in range(c.nv)}
    bond = {e: letters[c.nv + e] for e in range(len(c.edges))}
    ops, subs = [], []
    for v in range(c.nv):
        subs.append(site[v] + "".join(bond[e] for e, _ in c.incident[v]))
        ops.append(c.T[v].astype(_wide(c.dtype)))
    expr = ",".join(subs) + "->" + "".join(site[v] for v in range(c.nv))
    return np.einsum(expr, *ops, optimize="greedy")
