"""Host mirror of the reference's Julia surface for the BP simple-update path, over the C-ABI.

Same names, argument meaning and error behaviour as the reference functions that
`/root/reference/examples/2dIsing_dynamics.jl` calls:

    tensornetworkstate / random_tensornetworkstate   src/TensorNetworks/tensornetworkstate.jl:93-161
    BeliefPropagationCache(ψ)                         src/MessagePassing/beliefpropagationcache.jl:27-31
    apply_gates / apply_circuit                       src/Apply/apply_gates.jl:17-98,145
    update                                            src/MessagePassing/abstractbeliefpropagationcache.jl:223-259
    expect                                            src/expect.jl:54-82,114-135
    network / maxvirtualdim / messages / message      beliefpropagationcache.jl:23-25, abstracttensornetwork.jl:27-29

All numerics run on the GPU behind `libtnqs_b200.so`; this module only marshals.
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .gates import STATES, ArgumentError, gate_matrix, observable_matrix
from .graphs import NamedGraph, _as_vertex_list, forest_cover_edge_sequence

_DT = {np.dtype(np.complex64): _lib.TNQS_C64, np.dtype(np.complex128): _lib.TNQS_C128}


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class TensorNetworkState:
    """Host-resident TNS: `graph` + one numpy tensor `(d, χ_leg0, χ_leg1, …)` per vertex, bond legs
    in increasing edge id (`tensornetworkstate.jl:12-15`)."""

    def __init__(self, graph: NamedGraph, tensors: Dict[Hashable, np.ndarray], dtype=None):
        self.graph = graph
        dtype = np.dtype(dtype if dtype is not None else next(iter(tensors.values())).dtype)
        if dtype not in _DT:
            raise ArgumentError("tnqs_b200 supports ComplexF32 (complex64) and ComplexF64 (complex128) states")
        self.dtype = dtype
        self.tensors = {v: np.ascontiguousarray(tensors[v], dtype=dtype) for v in graph.vertices()}
        for v in graph.vertices():
            if self.tensors[v].ndim != 1 + graph.degree(v):
                raise ArgumentError(f"tensor on vertex {v} must have 1 + degree indices")

    def __getitem__(self, v):
        return self.tensors[v]

    def vertices(self):
        return self.graph.vertices()

    def scalartype(self):
        return self.dtype

    def maxvirtualdim(self) -> int:
        return max((max(t.shape[1:], default=1) for t in self.tensors.values()), default=1)


def tensornetworkstate(eltype, f, g: NamedGraph, sitetype: str = "S=1/2") -> TensorNetworkState:
    """Product state: `f(v)` is a state name ("↑", "↓", "0", "1", "+", …) or a vector
    (`tensornetworkstate.jl:141-161`)."""
    if sitetype not in ("S=1/2", "Qubit", "S=½"):
        raise ArgumentError("only spin-1/2 / qubit sites are supported on the device path")
    tensors = {}
    for v in g.vertices():
        s = f(v)
        if isinstance(s, str):
            if s not in STATES:
                raise ArgumentError(f'Unrecognized local state "{s}"')
            vec = np.array(STATES[s], dtype=eltype)
        elif isinstance(s, (list, tuple, np.ndarray)):
            vec = np.asarray(s, dtype=eltype)
        else:
            raise ArgumentError("Unrecognized local state constructor. Currently supported: Strings and Vectors.")
        tensors[v] = vec.reshape((-1,) + (1,) * g.degree(v))
    return TensorNetworkState(g, tensors, eltype)


def zerostate(eltype, g: NamedGraph) -> TensorNetworkState:
    return tensornetworkstate(eltype, lambda v: "↑", g)


def random_tensornetworkstate(eltype, g: NamedGraph, bond_dimension: int = 1, d: int = 2, seed=None) -> TensorNetworkState:
    """iid normal entries (`tensornetworkstate.jl:93-103`); the RNG stream is NumPy's."""
    rng = np.random.default_rng(seed)
    tensors = {}
    for v in g.vertices():
        shp = (d,) + (bond_dimension,) * g.degree(v)
        t = rng.standard_normal(shp) + 1j * rng.standard_normal(shp)
        tensors[v] = t.astype(eltype)
    return TensorNetworkState(g, tensors, eltype)


def random_bpc_on_device(eltype, g: NamedGraph, bond_dimension: int = 1, d: int = 2, seed: int = 1234, device: int = 0,
                         normalize: bool = True, shard_fn=None) -> "BeliefPropagationCache":
    """`BeliefPropagationCache(random_tensornetworkstate(eltype, g; bond_dimension))` with the tensors generated on
    the device (`tnqs_randomize_sites`): iid normal entries keyed by (seed, vertex), each tensor scaled to unit
    Frobenius norm when `normalize`.  For synthetic states too large to build on the host (16×16 at χ=64 is 53 GB).
    `shard_fn(bpc)` (e.g. `tnqs_b200.shard`) is applied before the tensors are filled, so each rank generates only
    the tensors it owns."""
    lib = _lib.load()
    uv = np.array(g.edge_uv(), dtype=np.int32).reshape(-1, 2)
    phys = np.full(g.nv, d, dtype=np.int32)
    bond = np.full(g.ne, bond_dimension, dtype=np.int32)
    h = C.c_void_p()
    _, uv_p = _i32(uv)
    _, ph_p = _i32(phys)
    _, bo_p = _i32(bond)
    if shard_fn is None:
        _lib.check(lib.tnqs_create(_DT[np.dtype(eltype)], g.nv, g.ne, uv_p, ph_p, bo_p, device, C.byref(h)))
        bpc = BeliefPropagationCache(None, device, _handle=h, _graph=g, _dtype=np.dtype(eltype), _seq=None)
    else:
        # create at bond dimension 1 (tiny), shard, then declare the real shapes: only the owner allocates a tensor
        one = np.ones(g.ne, dtype=np.int32)
        _, one_p = _i32(one)
        _lib.check(lib.tnqs_create(_DT[np.dtype(eltype)], g.nv, g.ne, uv_p, ph_p, one_p, device, C.byref(h)))
        bpc = BeliefPropagationCache(None, device, _handle=h, _graph=g, _dtype=np.dtype(eltype), _seq=None)
        shard_fn(bpc)
        for i, v in enumerate(g.vertices()):
            shp = np.array((d,) + (bond_dimension,) * len(g.incident[i]), dtype=np.int64)
            owned = bpc.owner[i] == bpc.rank
            buf = np.zeros(int(np.prod(shp)) if owned else 1, dtype=eltype)
            _lib.check(lib.tnqs_set_site(h, i, buf.ctypes.data_as(C.c_void_p), len(shp), shp.ctypes.data_as(C.POINTER(C.c_int64))))
    bpc.set_edge_sequence(forest_cover_edge_sequence(g))
    _lib.check(lib.tnqs_randomize_sites(bpc._h, C.c_uint64(int(seed)), int(bool(normalize))))
    return bpc


class BeliefPropagationCache:
    """Device-resident `BeliefPropagationCache` (`beliefpropagationcache.jl:9-15`): site tensors and
    messages live in HBM behind a `tnqs_handle`; constructing it runs no BP and leaves every message
    at its identity default."""

    def __init__(self, psi: TensorNetworkState, device: int = 0, _handle=None, _graph=None, _dtype=None,
                 _seq=None):
        self._lib = _lib.load()
        if _handle is not None:
            self._h, self.graph, self.dtype, self._seq = _handle, _graph, _dtype, _seq
            self.device = device
            return
        g = psi.graph
        self.graph, self.dtype, self.device = g, psi.dtype, device
        uv = np.array(g.edge_uv(), dtype=np.int32).reshape(-1, 2)
        phys = np.array([psi.tensors[v].shape[0] for v in g.vertices()], dtype=np.int32)
        bond = np.ones(g.ne, dtype=np.int32)
        for i, v in enumerate(g.vertices()):
            for k, (e, _) in enumerate(g.incident[i]):
                bond[e] = psi.tensors[v].shape[1 + k]
        h = C.c_void_p()
        uv_a, uv_p = _i32(uv)
        ph_a, ph_p = _i32(phys)
        bo_a, bo_p = _i32(bond)
        _lib.check(self._lib.tnqs_create(_DT[psi.dtype], g.nv, g.ne, uv_p, ph_p, bo_p, device, C.byref(h)))
        self._h = h
        for i, v in enumerate(g.vertices()):
            t = psi.tensors[v]
            shp = np.array(t.shape, dtype=np.int64)
            _lib.check(self._lib.tnqs_set_site(self._h, i, t.ctypes.data_as(C.c_void_p), t.ndim,
                                               shp.ctypes.data_as(C.POINTER(C.c_int64))))
        self._seq = None
        self.set_edge_sequence(forest_cover_edge_sequence(g))

    # -- lifetime ---------------------------------------------------------------------------------
    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._lib.tnqs_destroy(h)
            except Exception:
                pass
            self._h = None

    def copy(self) -> "BeliefPropagationCache":
        """`Base.copy` (`beliefpropagationcache.jl:35-37`) — a device-to-device clone."""
        h = C.c_void_p()
        _lib.check(self._lib.tnqs_clone(self._h, C.byref(h)))
        out = BeliefPropagationCache(None, self.device, _handle=h, _graph=self.graph, _dtype=self.dtype,
                                     _seq=self._seq)
        for k in ("owner", "rank", "world"):
            if hasattr(self, k):
                setattr(out, k, getattr(self, k))
        return out

    # -- accessors --------------------------------------------------------------------------------
    def set_edge_sequence(self, seq: Sequence[Tuple[Hashable, Hashable]]):
        idx = np.array([[self.graph.index[a], self.graph.index[b]] for a, b in seq], dtype=np.int32).reshape(-1, 2)
        a, p = _i32(idx)
        _lib.check(self._lib.tnqs_set_edge_sequence(self._h, p, len(idx)))
        self._seq = list(seq)

    def edge_sequence(self):
        return list(self._seq)

    def scalartype(self):
        return self.dtype

    def vertices(self):
        return self.graph.vertices()

    def bond_dims(self) -> np.ndarray:
        out = np.zeros(self.graph.ne, dtype=np.int32)
        _lib.check(self._lib.tnqs_get_bond_dims(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def maxvirtualdim(self) -> int:
        """`maxvirtualdim` (`abstracttensornetwork.jl:27-29`)."""
        b = self.bond_dims()
        return int(b.max()) if len(b) else 1

    def site_shape(self, v) -> Tuple[int, ...]:
        """Shape (d, χ_0, …) of the site tensor of `v` (no data transfer)."""
        nd = C.c_int(64)
        shp = (C.c_int64 * 64)()
        _lib.check(self._lib.tnqs_site_shape(self._h, self.graph.index[v], C.byref(nd), shp))
        return tuple(int(shp[k]) for k in range(nd.value))

    def site(self, v) -> np.ndarray:
        """`network(ψ_bpc)[v]`: download one site tensor."""
        i = self.graph.index[v]
        nd = C.c_int(64)
        shp = (C.c_int64 * 64)()
        _lib.check(self._lib.tnqs_site_shape(self._h, i, C.byref(nd), shp))
        shape = tuple(int(shp[k]) for k in range(nd.value))
        out = np.empty(shape, dtype=self.dtype)
        _lib.check(self._lib.tnqs_get_site(self._h, i, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def network(self) -> TensorNetworkState:
        """`network(ψ_bpc)` (`beliefpropagationcache.jl:24`): materialise the TNS on the host.  On a
        sharded cache every rank must call this (site tensors are broadcast from their owners)."""
        owner = getattr(self, "owner", None)
        if owner is None or getattr(self, "world", 1) == 1:
            return TensorNetworkState(self.graph, {v: self.site(v) for v in self.graph.vertices()}, self.dtype)
        import torch.distributed as dist
        tensors = {}
        for i, v in enumerate(self.graph.vertices()):
            box = [self.site(v) if owner[i] == self.rank else None]
            dist.broadcast_object_list(box, src=owner[i])
            tensors[v] = box[0]
        return TensorNetworkState(self.graph, tensors, self.dtype)

    def message(self, edge) -> np.ndarray:
        """`message(bpc, src => dst)` with the identity default (`abstractbeliefpropagationcache.jl:99-102`)."""
        m, _ = self._message(edge)
        return m

    def _message(self, edge):
        a, b = edge
        chi = int(self.bond_dims()[self.graph.edge_id(a, b)])
        out = np.empty((chi, chi), dtype=self.dtype)
        c, s = C.c_int(), C.c_int()
        _lib.check(self._lib.tnqs_get_message(self._h, self.graph.index[a], self.graph.index[b],
                                              out.ctypes.data_as(C.c_void_p), out.size, C.byref(c), C.byref(s)))
        return out, bool(s.value)

    def messages(self) -> Dict[Tuple[Hashable, Hashable], np.ndarray]:
        """`messages(bpc)`: only messages that have been set (empty right after construction)."""
        out = {}
        for (a, b) in self.graph.edges:
            for e in ((a, b), (b, a)):
                m, is_set = self._message(e)
                if is_set:
                    out[e] = m
        return out

    def setmessage(self, edge, m: np.ndarray):
        a, b = edge
        m = np.ascontiguousarray(m, dtype=self.dtype)
        _lib.check(self._lib.tnqs_set_message(self._h, self.graph.index[a], self.graph.index[b],
                                              m.ctypes.data_as(C.c_void_p), m.shape[0]))
        return self

    def setmessages(self, edges, ms):
        for e, m in zip(edges, ms):
            self.setmessage(e, m)
        return self

    def stats(self, reset: bool = False) -> dict:
        s = _lib.Stats()
        _lib.check(self._lib.tnqs_get_stats(self._h, C.byref(s), int(reset)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def set_profiling(self, on: bool):
        _lib.check(self._lib.tnqs_set_profiling(self._h, int(on)))


# ---------------------------------------------------------------------------------------------
# kwargs → C structs
# ---------------------------------------------------------------------------------------------

_SVD_ALGS = {"divide_and_conquer": 0, "qr_iteration": 1, "recursive": 2}


def _apply_opts(kw: Optional[dict]) -> _lib.ApplyOpts:
    """`apply_kwargs` of `apply_gates` → C struct: `maxdim`, `cutoff`, `normalize_tensors`, `sqrt_cutoff`
    (`simple_update.jl:21-33`) and what `factorize_svd` takes (`simple_update.jl:53-59`): `mindim`,
    `use_absolute_cutoff`, `use_relative_cutoff`, `alg`."""
    kw = dict(kw or {})
    o = _lib.ApplyOpts(0, 1, -1.0, 1, -1.0, 0, 1, 0, 0)
    for k, v in kw.items():
        if k == "maxdim":
            o.maxdim = int(v) if v is not None else 0
        elif k == "mindim":
            o.mindim = int(v)
        elif k == "cutoff":
            o.cutoff = float(v) if v is not None else -1.0
        elif k == "normalize_tensors":
            o.normalize_tensors = int(bool(v))
        elif k == "sqrt_cutoff":
            o.sqrt_cutoff = float(v) if v is not None else -1.0
        elif k == "use_absolute_cutoff":
            o.use_absolute_cutoff = int(bool(v))
        elif k == "use_relative_cutoff":
            o.use_relative_cutoff = int(bool(v))
        elif k == "alg":
            if v not in _SVD_ALGS:
                raise ArgumentError(f"unknown SVD algorithm {v!r} (supported: {sorted(_SVD_ALGS)})")
            o.svd_alg = _SVD_ALGS[v]
        else:
            raise ArgumentError(f"unsupported apply keyword {k!r} (supported: maxdim, mindim, cutoff, normalize_tensors, "
                                "sqrt_cutoff, use_absolute_cutoff, use_relative_cutoff, alg)")
    return o


def default_bp_update_kwargs(bpc) -> dict:
    """`default_bp_update_kwargs` (`beliefpropagationcache.jl:110-119`)."""
    if bpc.graph.is_tree():
        return dict(maxiter=1, tolerance=None, verbose=False)
    return dict(maxiter=25, tolerance=1e-5 if np.dtype(bpc.dtype) == np.complex64 else 1e-8, verbose=False)


def _bp_opts(bpc: BeliefPropagationCache, kw: Optional[dict]):
    base = default_bp_update_kwargs(bpc) if kw is None else dict(kw)
    keep = []
    o = _lib.BpOpts(0, -1.0, 0, None, 0)
    maxiter = base.get("maxiter", None)
    if maxiter is None:
        maxiter = 1 if bpc.graph.is_tree() else 25  # default_bp_maxiter, beliefpropagationcache.jl:39
    o.maxiter = int(maxiter)
    tol = base.get("tolerance", None)  # default_tolerance(::Algorithm"bp") = nothing (:54)
    if tol is not None:
        o.use_tolerance, o.tolerance = 1, float(tol)
    seq = base.get("edge_sequence", None)
    if seq is not None:
        idx = np.array([[bpc.graph.index[a], bpc.graph.index[b]] for a, b in seq], dtype=np.int32).reshape(-1, 2)
        a, p = _i32(idx)
        keep.append(a)
        o.edge_sequence, o.n_seq = p, len(idx)
    unknown = set(base) - {"maxiter", "tolerance", "verbose", "edge_sequence"}
    if unknown:
        raise ArgumentError(f"unsupported BP keyword(s) {sorted(unknown)}")
    return o, keep, bool(base.get("verbose", False)), tol


def _report_bp(rep, tol, verbose):
    """Same text as `abstractbeliefpropagationcache.jl:245-252`."""
    if tol is None:
        return
    if rep.converged:
        if verbose:
            print(f"BP converged to desired precision after {rep.niter} iterations.")
    else:
        msg = (f"BP did not converge to tolerance {tol} after {rep.niter} iterations "
               f"(final average message change: {rep.diff}).")
        print(msg) if verbose else warnings.warn(msg)


# ---------------------------------------------------------------------------------------------
# the public functions
# ---------------------------------------------------------------------------------------------

def update(bpc: BeliefPropagationCache, inplace: bool = False, **kwargs) -> BeliefPropagationCache:
    """`update(bpc; maxiter, tolerance, verbose, edge_sequence)` → updated copy
    (`abstractbeliefpropagationcache.jl:257-259`).  Keywords left out take the `Algorithm"bp"`
    defaults of `beliefpropagationcache.jl:52-72`: maxiter 25 (1 on trees), tolerance `nothing`."""
    out = bpc if inplace else bpc.copy()
    o, keep, verbose, tol = _bp_opts(out, kwargs)
    rep = _lib.BpReport()
    _lib.check(out._lib.tnqs_bp_update(out._h, C.byref(o), C.byref(rep)))
    _report_bp(rep, tol, verbose)
    out.last_bp_report = dict(niter=rep.niter, converged=bool(rep.converged), diff=rep.diff)
    return out


def circuit_arrays(circuit: Sequence, g: NamedGraph):
    """`toitensor(circuit, g, siteinds)` (`gate_definitions.jl:110-153`) down to flat arrays."""
    nverts, verts, mats = [], [], []
    memo = {}  # a Trotter layer repeats a handful of (name, parameter) pairs hundreds of times: build each matrix once
    for gate in circuit:
        if isinstance(gate, tuple) and len(gate) >= 2:
            vs = _as_vertex_list(gate[1])
            for v in vs:
                if v not in g.index:
                    raise ArgumentError(f"gate vertex {v!r} is not a vertex of the graph")
            par = gate[2] if len(gate) > 2 else None
            key = None
            if isinstance(gate[0], str) and (par is None or isinstance(par, (int, float, complex))
                                             or (isinstance(par, tuple) and all(isinstance(x, (int, float, complex)) for x in par))):
                key = (gate[0], len(vs), par)
            m = memo.get(key) if key is not None else None
            if m is None:
                m = np.asarray(gate_matrix(gate[0], len(vs), par), dtype=np.complex128).reshape(-1)
                if key is not None:
                    memo[key] = m
        else:
            raise ArgumentError("circuit entries must be (name_or_matrix, vertices[, params]) tuples")
        nverts.append(len(vs))
        idx = [g.index[v] for v in vs]
        verts.append((idx + [-1, -1])[:2] if len(idx) <= 2 else idx[:2])
        mats.append(m)
    mats = np.concatenate(mats) if mats else np.zeros(0, dtype=np.complex128)
    return (np.array(nverts, dtype=np.int32), np.array(verts, dtype=np.int32).reshape(-1, 2),
            np.ascontiguousarray(mats).view(np.float64))


def apply_gates(circuit: Sequence, psi, apply_kwargs: Optional[dict] = None,
                bp_update_kwargs: Optional[dict] = None, update_cache: bool = True, verbose: bool = False,
                inplace: bool = False, device: int = 0):
    """`apply_gates(circuit, ψ | ψ_bpc; apply_kwargs, bp_update_kwargs, update_cache, verbose)`
    → `(ψ′, errs)` (`apply_gates.jl:17-98`).  The input cache is not mutated unless `inplace=True`
    (an extension: the examples rebind the result, so in-place is observationally the same)."""
    if isinstance(psi, TensorNetworkState):  # apply_gates.jl:17-27
        bpc = BeliefPropagationCache(psi, device=device)
        kw = bp_update_kwargs if bp_update_kwargs is not None else default_bp_update_kwargs(bpc)
        bpc = update(bpc, inplace=True, **kw)
        bpc, errs = apply_gates(circuit, bpc, apply_kwargs, kw, update_cache, verbose, inplace=True)
        return bpc.network(), errs
    bpc: BeliefPropagationCache = psi
    nverts, verts, mats = circuit_arrays(circuit, bpc.graph)
    out = bpc if inplace else bpc.copy()
    ao = _apply_opts(apply_kwargs)
    bo, keep, bverbose, tol = _bp_opts(out, bp_update_kwargs)
    ng = len(nverts)
    errs = np.zeros(ng, dtype=np.float64)
    max_rep = ng + 1
    reps = (_lib.BpReport * max_rep)()
    nrep = C.c_int(0)
    nv_a, nv_p = _i32(nverts)
    vs_a, vs_p = _i32(verts)
    m_a, m_p = _f64(mats)
    _lib.check(out._lib.tnqs_apply_gates(out._h, ng, nv_p, vs_p, m_p, C.byref(ao), C.byref(bo), int(update_cache),
                                         errs.ctypes.data_as(C.POINTER(C.c_double)), reps, max_rep, C.byref(nrep)))
    out.last_bp_reports = [dict(niter=reps[i].niter, converged=bool(reps[i].converged), diff=reps[i].diff)
                           for i in range(min(nrep.value, max_rep))]
    for i in range(min(nrep.value, max_rep)):
        if verbose:
            print("Updating BP cache")
        _report_bp(reps[i], tol, bverbose or verbose)
    return out, errs


apply_circuit = apply_gates


def truncate(bpc: BeliefPropagationCache, maxdim: int, cutoff=None, edge_color: bool = True,
             normalize_tensors: bool = True, bp_update_kwargs: Optional[dict] = None,
             edge_groups: Optional[Sequence] = None, inplace: bool = False) -> BeliefPropagationCache:
    """`truncate(bpc; bp_update_kwargs, maxdim, cutoff, edge_color, normalize_tensors)` (`src/truncate.jl:12-38`):
    an identity two-site gate through the simple update on every truncatable edge (bond dimension > 1), one
    edge-colour group at a time with a BP `update` after each group (`edge_color=False`: after each edge).
    No new device code: a colour group is one batched `tnqs_apply_gates` call with `update_cache = 0`.
    `edge_groups` overrides the colouring (the reference's comes from an integer program and is not unique)."""
    from .graphs import edge_color as _edge_color
    out = bpc if inplace else bpc.copy()
    g = out.graph
    kw = bp_update_kwargs if bp_update_kwargs is not None else default_bp_update_kwargs(out)
    akw = dict(maxdim=maxdim, normalize_tensors=normalize_tensors)
    if cutoff is not None:
        akw["cutoff"] = cutoff
    if edge_groups is None:
        if edge_color:
            z = max(g.degree(v) for v in g.vertices())
            edge_groups = _edge_color(g, z)
        else:
            edge_groups = [[e] for e in g.edges]
    phys = {v: out.site_shape(v)[0] for v in g.vertices()}
    for grp in edge_groups:
        dims = out.bond_dims()
        circ = []
        for (a, b) in grp:
            if dims[g.edge_id(a, b)] <= 1:  # truncatable_edge (truncate.jl:5-10)
                continue
            d = phys[a] * phys[b]
            circ.append((np.eye(d, dtype=complex), [a, b]))
        if circ:
            out, _ = apply_gates(circ, out, apply_kwargs=akw, update_cache=False, inplace=True)
        out = update(out, inplace=True, **kw)
    return out


# ---------------------------------------------------------------------------------------------
# scalars of the BP fixed point, rescaling, norm
# ---------------------------------------------------------------------------------------------
def vertex_scalars(bpc: BeliefPropagationCache, vertices: Optional[Sequence] = None) -> np.ndarray:
    """`vertex_scalars(bpc, vertices)` (`abstractbeliefpropagationcache.jl:22-28,134-138`): ⟨T_v|msgs|T_v⟩."""
    vs = list(bpc.graph.vertices()) if vertices is None else list(vertices)
    idx, idx_p = _i32(np.array([bpc.graph.index[v] for v in vs], dtype=np.int32))
    out = np.zeros(2 * len(vs), dtype=np.float64)
    _lib.check(bpc._lib.tnqs_vertex_scalars(bpc._h, len(vs), idx_p, out.ctypes.data_as(C.POINTER(C.c_double))))
    return out.view(np.complex128).copy()


def vertex_scalar(bpc: BeliefPropagationCache, v):
    return vertex_scalars(bpc, [v])[0]


def edge_scalar(bpc: BeliefPropagationCache, edge):
    """`edge_scalar(bpc, e)` (`beliefpropagationcache.jl:47-49`): scalar(message(e) * message(reverse(e)))."""
    a, b = edge
    return np.sum(bpc.message((a, b)).astype(np.complex128) * bpc.message((b, a)).astype(np.complex128))


def edge_scalars(bpc: BeliefPropagationCache, edges: Optional[Sequence] = None) -> np.ndarray:
    es = list(bpc.graph.edges) if edges is None else list(edges)
    return np.array([edge_scalar(bpc, e) for e in es], dtype=np.complex128)


def scalar_factors_quotient(bpc: BeliefPropagationCache):
    """`scalar_factors_quotient` (`abstractbeliefpropagationcache.jl:146-148`)."""
    return vertex_scalars(bpc), edge_scalars(bpc)


def freenergy(bpc: BeliefPropagationCache):
    """`freenergy` (`abstractbeliefpropagationcache.jl:289-300`)."""
    num, den = scalar_factors_quotient(bpc)
    if np.any(den == 0):
        return -np.inf
    return np.sum(np.log(num)) - np.sum(np.log(den))


def partitionfunction(bpc: BeliefPropagationCache):
    """`partitionfunction` (`abstractbeliefpropagationcache.jl:302-304`)."""
    return np.exp(freenergy(bpc))


def rescale_messages(bpc: BeliefPropagationCache, edges: Optional[Sequence] = None, inplace: bool = False):
    """`rescale_messages!` (`beliefpropagationcache.jl:127-140`): both messages of an edge normalised, then divided
    by the square root of their contraction so that `edge_scalar == 1`."""
    out = bpc if inplace else bpc.copy()
    for (a, b) in (list(out.graph.edges) if edges is None else list(edges)):
        me, mer = out.message((a, b)).astype(np.complex128), out.message((b, a)).astype(np.complex128)
        me /= np.linalg.norm(me)
        mer /= np.linalg.norm(mer)
        n = np.sum(me * mer)
        if n.imag == 0:
            sg = np.sign(n.real)
            me *= sg
            n *= sg
        out.setmessage((a, b), me / np.sqrt(n))
        out.setmessage((b, a), mer / np.sqrt(n))
    return out


def rescale_vertices(bpc: BeliefPropagationCache, vertices: Optional[Sequence] = None, inplace: bool = False):
    """`rescale_vertices!` (`beliefpropagationcache.jl:82-101`) for a TensorNetworkState: tn[v] *= s/√vertex_scalar."""
    out = bpc if inplace else bpc.copy()
    vs = list(out.graph.vertices()) if vertices is None else list(vertices)
    vn = vertex_scalars(out, vs)
    sg = np.where(vn.imag == 0, np.sign(vn.real), 1.0)
    f = np.ascontiguousarray(sg / np.sqrt(vn), dtype=np.complex128)
    idx, idx_p = _i32(np.array([out.graph.index[v] for v in vs], dtype=np.int32))
    _lib.check(out._lib.tnqs_scale_sites(out._h, len(vs), idx_p, f.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))))
    return out


def rescale(bpc: BeliefPropagationCache, inplace: bool = False) -> BeliefPropagationCache:
    """`rescale` / `rescale!` (`abstractbeliefpropagationcache.jl:318-328`)."""
    out = bpc if inplace else bpc.copy()
    rescale_messages(out, inplace=True)
    rescale_vertices(out, inplace=True)
    return out


def norm_sqr(psi, alg: str = "bp", cache_update_kwargs: Optional[dict] = None, device: int = 0):
    """`norm_sqr(ψ | ψ_bpc; alg="bp")` (`src/norm_sqr.jl:72-84`): the BP partition function of ⟨ψ|ψ⟩."""
    if alg != "bp":
        raise ArgumentError('norm_sqr: only alg="bp" is implemented on the device path')
    if isinstance(psi, TensorNetworkState):
        bpc = BeliefPropagationCache(psi, device=device)
        kw = cache_update_kwargs if cache_update_kwargs is not None else default_bp_update_kwargs(bpc)
        psi = update(bpc, inplace=True, **kw)
    return partitionfunction(psi)


def normalize(tns: TensorNetworkState, alg: str = "bp", cache_update_kwargs: Optional[dict] = None,
              device: int = 0) -> TensorNetworkState:
    """`normalize(tns; alg="bp")` (`src/normalize.jl:1-6`): update, rescale!, network."""
    if alg != "bp":
        raise ArgumentError('normalize: only alg="bp" is implemented')
    bpc = BeliefPropagationCache(tns, device=device)
    kw = cache_update_kwargs if cache_update_kwargs is not None else default_bp_update_kwargs(bpc)
    bpc = update(bpc, inplace=True, **kw)
    rescale(bpc, inplace=True)
    return bpc.network()


def _symmetric_gauge_factors(mx: np.ndarray, my: np.ndarray, regularization: float):
    """χ×χ algebra of one edge of `symmetric_gauge!` (`src/symmetric_gauge.jl:12-40`): returns (X_src, X_dst, S) with
    ψ_src ← ψ_src ×_e X_src, ψ_dst ← ψ_dst ×_e X_dst and both new messages diag(S).  ITensors' `eigen` reads a message
    on (l, l') as the map l → l', i.e. as the transpose of the m[ket, bra] storage used here.

    Written without the inverse square roots of the reference (the CPU test restatement follows the reference line by line,
    so the two are independent): with
    C = X^{1/2}·(Y^{1/2})ᵀ = U·S·V†, X^{-1/2}·U·√S = (Y^{1/2})ᵀ·V·S^{-1/2} and Y^{-1/2}·conj(V)·√S = (X^{1/2})ᵀ·conj(U)·S^{-1/2};
    the Hermitian square roots come from a Schur decomposition (scipy.linalg.sqrtm), the SVD from LAPACK gesvd."""
    import scipy.linalg
    n = mx.shape[0]
    xr = np.asarray(mx, dtype=np.complex128).T + regularization * np.eye(n)
    yr = np.asarray(my, dtype=np.complex128).T + regularization * np.eye(n)
    for h in (xr, yr):
        if np.min(np.linalg.eigvalsh(0.5 * (h + h.conj().T))) < 0:
            raise ValueError("DomainError: sqrt of a negative message eigenvalue")
    root_x = scipy.linalg.sqrtm(xr)
    root_y = scipy.linalg.sqrtm(yr)
    u, sv, vh = scipy.linalg.svd(root_x @ root_y.T, lapack_driver="gesvd")
    v = vh.conj().T
    isq = 1.0 / np.sqrt(sv)
    return (root_y.T @ v) * isq, (root_x.T @ u.conj()) * isq, sv


def symmetric_gauge(x, regularization: Optional[float] = None, cache_update_kwargs: Optional[dict] = None,
                    inplace: bool = False, device: int = 0):
    """`symmetric_gauge(bp_cache | tns; regularization)` (`src/symmetric_gauge.jl:1-68`), without SVD truncation
    keywords: every bond is regauged so that its two messages become the same diagonal matrix.  The χ×χ
    eigen/SVD algebra runs on the host (NumPy/LAPACK, as the reference runs it on the CPU); the site tensors are
    updated on the device in one batched `tnqs_apply_leg_matrices` call (edges touch disjoint legs)."""
    if isinstance(x, TensorNetworkState):
        bpc = BeliefPropagationCache(x, device=device)
        kw = cache_update_kwargs if cache_update_kwargs is not None else dict(maxiter=40)
        bpc = update(bpc, inplace=True, **kw)
        return symmetric_gauge(bpc, regularization=regularization, inplace=True).network()
    out = x if inplace else x.copy()
    g = out.graph
    if regularization is None:
        regularization = 10 * np.finfo(np.float32 if out.dtype == np.complex64 else np.float64).eps
    verts, nbrs, mats, new_msgs = [], [], [], []
    for (a, b) in g.edges:
        xs, xd_, sv = _symmetric_gauge_factors(out.message((a, b)), out.message((b, a)), regularization)
        verts += [g.index[a], g.index[b]]
        nbrs += [g.index[b], g.index[a]]
        mats += [np.ascontiguousarray(xs, dtype=np.complex128).reshape(-1), np.ascontiguousarray(xd_, dtype=np.complex128).reshape(-1)]
        new_msgs.append(((a, b), np.diag(sv)))
    if verts:
        v_a, v_p = _i32(np.array(verts, dtype=np.int32))
        n_a, n_p = _i32(np.array(nbrs, dtype=np.int32))
        m = np.ascontiguousarray(np.concatenate(mats)).view(np.float64)
        _lib.check(out._lib.tnqs_apply_leg_matrices(out._h, len(verts), v_p, n_p, m.ctypes.data_as(C.POINTER(C.c_double))))
    for (a, b), s_ in new_msgs:
        out.setmessage((a, b), s_)
        out.setmessage((b, a), s_)
    return out


def symmetrize_and_normalize(bpc: BeliefPropagationCache, **kwargs) -> BeliefPropagationCache:
    """`symmetrize_and_normalize` (`src/symmetric_gauge.jl:70-74`): rescale, then symmetric gauge."""
    return symmetric_gauge(rescale(bpc), inplace=True, **kwargs)


def renyi_entropy(bpc: BeliefPropagationCache, edge, alpha: float = 1.0) -> float:
    """`renyi_entropy(bp_cache, e; α)` (`src/entanglement.jl:73-86`): Rényi entropy across a bond from the two
    converged messages on it.  The reference forms ρ = √m2ᵀ·m1·√m2ᵀ (√ via `pseudo_sqrt_inv_sqrt`, eigenvalues below
    10·eps dropped) and takes the spectrum of ρ / tr ρ (`:21-29`).  Here the same spectrum is read off the similar matrix
    (m2⁺)ᵀ·m1, m2⁺ = m2 with its sub-cutoff eigenvalues zeroed — no matrix square root (the CPU test restatement follows
    the reference's route; the two are independent).  χ×χ host algebra on `tnqs_get_message`; exact on trees."""
    a, b = edge
    m1 = bpc.message((a, b)).astype(np.complex128)
    m2 = bpc.message((b, a)).astype(np.complex128)
    eps = np.finfo(np.float32 if bpc.dtype == np.complex64 else np.float64).eps
    lam, q = np.linalg.eigh(m2)  # lower triangle, as LAPACK heev (safe_eigen, utils.jl:94-108)
    keep = ~((lam == 0) | (np.abs(lam) < 10 * eps))
    if np.any(lam[keep] < 0):
        raise ValueError("DomainError: sqrt of a negative message eigenvalue")
    m2p = (q * np.where(keep, lam, 0.0)) @ q.conj().T
    ev = np.real(np.linalg.eigvals(m2p.T @ m1))  # spectrum of √m2ᵀ·m1·√m2ᵀ (similar matrices), real and ≥ 0 at a BP fixed point
    ev = ev / np.sum(ev)
    ev = ev[np.abs(ev) > 10 * eps]  # eps of the state's real type (entanglement.jl:26)
    if alpha == 1:
        return float(-np.sum(ev * np.log(ev)))
    return float(np.log(np.sum(ev ** alpha)) / (1 - alpha))


def von_neumann_entanglement_entropy(psi, edge, alg: str = "bp", cache_update_kwargs: Optional[dict] = None,
                                     device: int = 0) -> float:
    """`von_neumann_entanglement_entropy(ψ | bpc, e; alg="bp")` (`src/entanglement.jl`): α = 1."""
    if alg != "bp":
        raise ArgumentError('von_neumann_entanglement_entropy: only alg="bp" is implemented')
    if isinstance(psi, TensorNetworkState):
        bpc = BeliefPropagationCache(psi, device=device)
        kw = cache_update_kwargs if cache_update_kwargs is not None else default_bp_update_kwargs(bpc)
        psi = update(bpc, inplace=True, **kw)
    return renyi_entropy(psi, edge, 1.0)


def _collect_observable(obs, g: NamedGraph):
    """`collectobservable` (`expect.jl:159-175`)."""
    coeff = 1 if len(obs) == 2 else obs[-1]
    verts = _as_vertex_list(obs[1])
    op = obs[0]
    if isinstance(op, str):
        ops = [c for c in op]
    elif isinstance(op, (list, tuple)) and all(isinstance(o, str) for o in op):
        ops = list(op)
    else:
        raise RuntimeError("Invalid observable, did not recognize operator specification. Either a single "
                           "string (one pauli character per vertex) or a vector of strings (one string per "
                           "vertex) expected.")
    if len(ops) != len(verts):
        raise RuntimeError("Invalid observable: need as many operators as vertices passed.")
    return ops, verts, coeff


def steiner_path(g: NamedGraph, terminals: Sequence) -> Optional[List]:
    """The Steiner tree of `terminals` (`expect.jl:67`, `rdm.jl:58`: `steiner_tree(network(cache), vs)`) when it is a
    simple path: shortest paths (BFS, neighbours in incidence order) between consecutive terminals, ordered along the
    path.  For two terminals this is the reference's region; returns None when the union is not a simple path."""
    idx = [g.index[v] for v in terminals]

    def bfs(a, b):
        prev = {a: -1}
        queue = [a]
        while queue:
            x = queue.pop(0)
            if x == b:
                break
            for _, w in g.incident[x]:
                if w not in prev:
                    prev[w] = x
                    queue.append(w)
        path = [b]
        while path[-1] != a:
            path.append(prev[path[-1]])
        return path[::-1]

    # order the terminals along a path: start from one end of the two farthest-apart terminals
    if len(idx) == 2:
        order = idx
    else:
        dist = {(a, b): len(bfs(a, b)) for a in idx for b in idx if a != b}
        a0, _ = max(dist, key=dist.get)
        order = sorted(idx, key=lambda t: 0 if t == a0 else dist[(a0, t)])
    path = [order[0]]
    for a, b in zip(order[:-1], order[1:]):
        path += bfs(a, b)[1:]
    if len(set(path)) != len(path):
        return None
    return path


def _site_contract(bpc: "BeliefPropagationCache", v: int, custom: Dict[int, np.ndarray], open_nbr: int, open_phys: bool,
                   op: Optional[np.ndarray]) -> np.ndarray:
    """`tnqs_site_contract`: one site of a path-region contraction; returns the n×n complex128 matrix out[ket][bra]."""
    lib = bpc._lib
    nb = np.array(list(custom.keys()), dtype=np.int32)
    mats = (np.concatenate([np.ascontiguousarray(m, dtype=np.complex128).reshape(-1) for m in custom.values()]).view(np.float64)
            if custom else np.zeros(0))
    d = 2
    chi_max = int(max(bpc.bond_dims())) if bpc.graph.ne else 1
    cap = (4 * chi_max) ** 2
    out = np.zeros(cap, dtype=np.complex128)
    n = C.c_int(0)
    nb_a, nb_p = _i32(nb)
    m_a, m_p = _f64(mats)
    op_p = None
    if op is not None:
        op_a = np.ascontiguousarray(op, dtype=np.complex128)
        op_p = op_a.ctypes.data_as(C.POINTER(C.c_double))
    _lib.check(lib.tnqs_site_contract(bpc._h, int(v), len(nb), nb_p if len(nb) else None, m_p if len(nb) else None,
                                      int(open_nbr), int(bool(open_phys)), op_p, out.ctypes.data_as(C.POINTER(C.c_double)),
                                      cap, C.byref(n)))
    return out[:n.value * n.value].reshape(n.value, n.value).copy()


def _path_numerator(bpc: "BeliefPropagationCache", path: List[int], ops: Dict[int, np.ndarray]):
    """Σ over the path region of ⟨T| ops ⊗ incoming messages |T⟩ (`contract_region`, expect.jl:71-78), walking the path:
    every site passes a χ×χ partial contraction to the next one as if it were the message on that bond."""
    L = None
    for k, v in enumerate(path):
        custom = {} if L is None else {path[k - 1]: L}
        if k + 1 < len(path):
            L = _site_contract(bpc, v, custom, path[k + 1], False, ops.get(v))
        else:
            rho = _site_contract(bpc, v, custom, -1, True, ops.get(v))
            return np.trace(rho)


def expect(psi, observable, alg: Optional[str] = "bp", cache_update_kwargs: Optional[dict] = None, device: int = 0):
    """`expect(ψ | ψ_bpc, obs; alg="bp")` (`expect.jl:54-82,114-135`).  `obs = (ops, vertices[, coeff])` or a list of
    them.  Device path: single-site, adjacent two-site and — for regions whose Steiner tree is a path — multi-site
    observables (`expect.jl:67-81`)."""
    if alg != "bp":
        raise RuntimeError("Expected alg = \"bp\": exact and boundary-MPS contraction are outside the "
                           "accelerated path (export with network(ψ_bpc) and use the reference for those)")
    if isinstance(psi, TensorNetworkState):  # expect.jl:123-135
        bpc = BeliefPropagationCache(psi, device=device)
        kw = cache_update_kwargs or default_bp_update_kwargs(bpc)
        bpc = update(bpc, inplace=True, **kw)
        return expect(bpc, observable, alg)
    bpc: BeliefPropagationCache = psi
    single = isinstance(observable, tuple)
    obs_list = [observable] if single else list(observable)
    out: List = [None] * len(obs_list)
    one, two, many = [], [], []
    for i, obs in enumerate(obs_list):
        ops, verts, coeff = _collect_observable(obs, bpc.graph)
        if coeff == 0:
            out[i] = 0 * coeff
        elif len(verts) == 1:
            one.append((i, ops, verts, coeff))
        elif len(verts) == 2 and bpc.graph.has_edge(verts[0], verts[1]):
            two.append((i, ops, verts, coeff))
        else:
            many.append((i, ops, verts, coeff))
    lib = bpc._lib
    if one:
        vs_a, vs_p = _i32([bpc.graph.index[o[2][0]] for o in one])
        m_a, m_p = _f64(np.concatenate([observable_matrix(o[1][0]).reshape(-1) for o in one]).view(np.float64))
        res = np.zeros(2 * len(one))
        _lib.check(lib.tnqs_expect_local(bpc._h, len(one), vs_p, m_p, res.ctypes.data_as(C.POINTER(C.c_double))))
        for k, o in enumerate(one):
            out[o[0]] = o[3] * complex(res[2 * k], res[2 * k + 1])
    if two:
        vs_a, vs_p = _i32([[bpc.graph.index[v] for v in o[2]] for o in two])
        m_a, m_p = _f64(np.concatenate([observable_matrix(c).reshape(-1) for o in two for c in o[1]]).view(np.float64))
        res = np.zeros(2 * len(two))
        _lib.check(lib.tnqs_expect_two_site(bpc._h, len(two), vs_p, m_p, res.ctypes.data_as(C.POINTER(C.c_double))))
        for k, o in enumerate(two):
            out[o[0]] = o[3] * complex(res[2 * k], res[2 * k + 1])
    for (i, ops, verts, coeff) in many:  # expect.jl:67-81 over a Steiner path
        if len(set(verts)) != len(verts):
            raise RuntimeError("Invalid observable: a vertex appears twice.")
        path = steiner_path(bpc.graph, verts)
        if path is None:
            raise NotImplementedError("device expect: the Steiner tree of these vertices is not a path")
        opm = {bpc.graph.index[v]: observable_matrix(o) for v, o in zip(verts, ops)}
        numer = _path_numerator(bpc, path, opm)
        denom = _path_numerator(bpc, path, {})
        out[i] = coeff * numer / denom
    return out[0] if single else out


def reduced_density_matrix(psi, vs, alg: Optional[str] = "bp", normalize: bool = True,
                           cache_update_kwargs: Optional[dict] = None, device: int = 0) -> np.ndarray:
    """`reduced_density_matrix(ψ | ψ_bpc, vs; alg="bp", normalize)` (`src/rdm.jl:52-73`): BP reduced density matrix of one
    vertex, or of two vertices joined through their Steiner path; ρ[(s_1, s_2), (s_1', s_2')] with `vs[0]` the slower
    index, trace 1 when `normalize` (`normalize_rdm`)."""
    if alg != "bp":
        raise RuntimeError("Expected alg = \"bp\" on the device path")
    if isinstance(psi, TensorNetworkState):
        bpc = BeliefPropagationCache(psi, device=device)
        kw = cache_update_kwargs or default_bp_update_kwargs(bpc)
        bpc = update(bpc, inplace=True, **kw)
        return reduced_density_matrix(bpc, vs, alg, normalize)
    bpc: BeliefPropagationCache = psi
    vs = _as_vertex_list(vs)
    g = bpc.graph
    if len(vs) == 1:
        rho = _site_contract(bpc, g.index[vs[0]], {}, -1, True, None)
    elif len(vs) == 2:
        path = steiner_path(g, vs)
        d = 2
        # open the first site: E[(s,b),(s',b')]; every (s,s') block is a χ×χ matrix that travels down the path
        E = _site_contract(bpc, path[0], {}, path[1], True, None)
        chi = E.shape[0] // d
        E = E.reshape(d, chi, d, chi)
        rho = np.zeros((d, d, d, d), dtype=np.complex128)  # [s1, s2, s1', s2']
        for s in range(d):
            for sp in range(d):
                L = np.ascontiguousarray(E[s, :, sp, :])
                for k in range(1, len(path) - 1):
                    L = _site_contract(bpc, path[k], {path[k - 1]: L}, path[k + 1], False, None)
                r2 = _site_contract(bpc, path[-1], {path[-2]: L}, -1, True, None)
                rho[s, :, sp, :] = r2
        rho = rho.reshape(d * d, d * d)
    else:
        raise NotImplementedError("device reduced_density_matrix supports one or two vertices")
    if normalize:
        rho = rho / np.trace(rho)
    return rho


def network(bpc: BeliefPropagationCache) -> TensorNetworkState:
    return bpc.network()


def maxvirtualdim(x) -> int:
    return x.maxvirtualdim()


def messages(bpc: BeliefPropagationCache):
    return bpc.messages()


def message(bpc: BeliefPropagationCache, edge):
    return bpc.message(edge)
