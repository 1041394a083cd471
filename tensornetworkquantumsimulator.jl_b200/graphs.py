"""Host-side graph construction for the BP / simple-update hot path.

The reference keeps graph construction, edge colouring and the BP edge schedule in the host
language (NamedGraphs.jl / SimpleGraphAlgorithms.jl, un-vendored; call sites
`/root/reference/src/MessagePassing/beliefpropagationcache.jl:27-29`,
`/root/reference/examples/2dIsing_dynamics.jl:9,25`).  This module is the Python host mirror:
it produces the integer (vertex, edge) tables that cross the C-ABI; no numerics live here.

Vertex names are arbitrary hashables (tuples for grids, ints for Eagle), exactly like
`named_grid` / `NamedGraph` in the reference; the device only ever sees 0-based integers.
"""
from __future__ import annotations

from collections import deque
from typing import Dict, Hashable, Iterable, List, Sequence, Tuple

Vertex = Hashable


class NamedGraph:
    """Undirected simple graph with named vertices and a fixed edge numbering.

    `edges[e] = (u, v)` (names, in insertion orientation).  `incident[i]` lists, for vertex index
    `i`, the `(edge_id, neighbour_index)` pairs in increasing edge id: this order IS the bond-leg
    order of the site tensor `T_v[s, leg_0, leg_1, ...]` on host and on the device.
    """

    def __init__(self, vertices: Iterable[Vertex], edges: Iterable[Tuple[Vertex, Vertex]] = ()):
        self.vertex_names: List[Vertex] = list(vertices)
        self.index: Dict[Vertex, int] = {v: i for i, v in enumerate(self.vertex_names)}
        if len(self.index) != len(self.vertex_names):
            raise ValueError("duplicate vertex names")
        self.edges: List[Tuple[Vertex, Vertex]] = []
        self._edge_id: Dict[Tuple[int, int], int] = {}
        self.incident: List[List[Tuple[int, int]]] = [[] for _ in self.vertex_names]
        for u, v in edges:
            self.add_edge(u, v)

    # -- construction -------------------------------------------------------------------------
    def add_edge(self, u: Vertex, v: Vertex) -> int:
        iu, iv = self.index[u], self.index[v]
        if iu == iv:
            raise ValueError("self loops are not supported")
        key = (min(iu, iv), max(iu, iv))
        if key in self._edge_id:
            return self._edge_id[key]
        e = len(self.edges)
        self.edges.append((u, v))
        self._edge_id[key] = e
        self.incident[iu].append((e, iv))
        self.incident[iv].append((e, iu))
        return e

    # -- queries ------------------------------------------------------------------------------
    @property
    def nv(self) -> int:
        return len(self.vertex_names)

    @property
    def ne(self) -> int:
        return len(self.edges)

    def vertices(self) -> List[Vertex]:
        return list(self.vertex_names)

    def has_edge(self, u: Vertex, v: Vertex) -> bool:
        if u not in self.index or v not in self.index:
            return False
        iu, iv = self.index[u], self.index[v]
        return (min(iu, iv), max(iu, iv)) in self._edge_id

    def edge_id(self, u: Vertex, v: Vertex) -> int:
        iu, iv = self.index[u], self.index[v]
        return self._edge_id[(min(iu, iv), max(iu, iv))]

    def neighbors(self, v: Vertex) -> List[Vertex]:
        return [self.vertex_names[j] for _, j in self.incident[self.index[v]]]

    def degree(self, v: Vertex) -> int:
        return len(self.incident[self.index[v]])

    def edge_uv(self) -> List[Tuple[int, int]]:
        """0-based integer endpoints per edge id (what `tnqs_create` receives)."""
        return [(self.index[u], self.index[v]) for u, v in self.edges]

    def leg_of(self, iv: int, e: int) -> int:
        """Position of edge `e` among the bond legs of vertex index `iv`."""
        for pos, (ee, _) in enumerate(self.incident[iv]):
            if ee == e:
                return pos
        raise KeyError((iv, e))

    def is_connected(self) -> bool:
        if self.nv == 0:
            return True
        seen = {0}
        dq = deque([0])
        while dq:
            i = dq.popleft()
            for _, j in self.incident[i]:
                if j not in seen:
                    seen.add(j)
                    dq.append(j)
        return len(seen) == self.nv

    def is_tree(self) -> bool:
        return self.is_connected() and self.ne == self.nv - 1

    def bipartition(self):
        """Return a 0/1 colour per vertex index, or None if the graph is not bipartite."""
        col = [-1] * self.nv
        for s in range(self.nv):
            if col[s] >= 0:
                continue
            col[s] = 0
            dq = deque([s])
            while dq:
                i = dq.popleft()
                for _, j in self.incident[i]:
                    if col[j] < 0:
                        col[j] = 1 - col[i]
                        dq.append(j)
                    elif col[j] == col[i]:
                        return None
        return col

    def center(self) -> List[Vertex]:
        """Vertices of minimum eccentricity (Graphs.jl `center`)."""
        ecc = []
        for s in range(self.nv):
            dist = {s: 0}
            dq = deque([s])
            while dq:
                i = dq.popleft()
                for _, j in self.incident[i]:
                    if j not in dist:
                        dist[j] = dist[i] + 1
                        dq.append(j)
            ecc.append(max(dist.values()))
        m = min(ecc)
        return [self.vertex_names[i] for i, x in enumerate(ecc) if x == m]


# ---------------------------------------------------------------------------------------------
# lattice builders
# ---------------------------------------------------------------------------------------------

def named_grid(dims, periodic: bool = False) -> NamedGraph:
    """`named_grid((nx, ny[, nz]); periodic)`: hyper-cubic lattice, 1-based tuple names, first
    coordinate fastest (Julia CartesianIndices order), as used by
    `/root/reference/examples/2dIsing_dynamics.jl:9` and `3dIsing_dynamics.jl`."""
    if isinstance(dims, int):
        dims = (dims,)
    dims = tuple(int(d) for d in dims)
    nd = len(dims)

    def names():
        idx = [1] * nd
        total = 1
        for d in dims:
            total *= d
        for _ in range(total):
            yield tuple(idx) if nd > 1 else idx[0]
            for k in range(nd):
                idx[k] += 1
                if idx[k] <= dims[k]:
                    break
                idx[k] = 1

    g = NamedGraph(names())
    for v in g.vertices():
        vt = v if nd > 1 else (v,)
        for k in range(nd):
            if dims[k] == 1:
                continue
            w = list(vt)
            if vt[k] < dims[k]:
                w[k] += 1
            elif periodic and dims[k] > 2:
                w[k] = 1
            else:
                continue
            wn = tuple(w) if nd > 1 else w[0]
            g.add_edge(v, wn)
    return g


def named_path_graph(n: int) -> NamedGraph:
    return NamedGraph(range(1, n + 1), [(i, i + 1) for i in range(1, n)])


def named_comb_tree(dims) -> NamedGraph:
    """`named_comb_tree((nx, ny))`: backbone (i,1), i=1..nx, with a tooth (i,1)-(i,2)-…-(i,ny) on
    every backbone vertex (used by `/root/reference/test/test_beliefpropagation.jl:11`)."""
    nx, ny = dims
    g = NamedGraph((i, j) for j in range(1, ny + 1) for i in range(1, nx + 1))
    for i in range(1, nx):
        g.add_edge((i, 1), (i + 1, 1))
    for i in range(1, nx + 1):
        for j in range(1, ny):
            g.add_edge((i, j), (i, j + 1))
    return g


def eagle_heavy_hex() -> NamedGraph:
    """IBM Eagle 127-qubit heavy-hex coupling map: 7 qubit rows chained
    linearly plus 24 bridge qubits.  127 vertices, 144 edges, max degree 3, bipartite."""
    rows = [range(0, 14), range(18, 33), range(37, 52), range(56, 71), range(75, 90),
            range(94, 109), range(113, 127)]
    g = NamedGraph(range(127))
    for r in rows:
        r = list(r)
        for a, b in zip(r[:-1], r[1:]):
            g.add_edge(a, b)
    bridges = [(0, 14, 18), (4, 15, 22), (8, 16, 26), (12, 17, 30),
               (20, 33, 39), (24, 34, 43), (28, 35, 47), (32, 36, 51),
               (37, 52, 56), (41, 53, 60), (45, 54, 64), (49, 55, 68),
               (58, 71, 77), (62, 72, 81), (66, 73, 85), (70, 74, 89),
               (75, 90, 94), (79, 91, 98), (83, 92, 102), (87, 93, 106),
               (96, 109, 114), (100, 110, 118), (104, 111, 122), (108, 112, 126)]
    for a, b, c in bridges:
        g.add_edge(a, b)
        g.add_edge(b, c)
    return g


def build_graph_from_gates(circuit: Sequence) -> NamedGraph:
    """Graph induced by a circuit of `(name, vertices[, param])` tuples
    (`/root/reference/src/graph_ops.jl:50-64`); errors if it is disconnected."""
    names: List[Vertex] = []
    seen = set()
    for gate in circuit:
        for v in _as_vertex_list(gate[1]):
            if v not in seen:
                seen.add(v)
                names.append(v)
    g = NamedGraph(names)
    for gate in circuit:
        q = _as_vertex_list(gate[1])
        if len(q) == 2:
            g.add_edge(q[0], q[1])
    if not g.is_connected():
        raise RuntimeError(
            "The circuit graph is not connected, meaning the resulting tensor network will be "
            "disconnected which we do not support.")
    return g


build_graph_from_circuit = build_graph_from_gates


def _as_vertex_list(vs) -> list:
    """`collect_vertices` (`/root/reference/src/utils.jl:110-160`): a list of vertices stays a
    list; a single vertex name (e.g. a tuple `(1,1)` or an int) becomes a one-element list."""
    if isinstance(vs, list):
        return list(vs)
    return [vs]


# ---------------------------------------------------------------------------------------------
# edge colouring (host side; the reference calls SimpleGraphAlgorithms.edge_color, an IP solver)
# ---------------------------------------------------------------------------------------------

def edge_color(g: NamedGraph, k: int) -> List[List[Tuple[Vertex, Vertex]]]:
    """Proper edge colouring with at most `k` colours: list of colour groups, each a list of
    `(u, v)` vertex-name pairs forming a matching.  Bipartite graphs are coloured with exactly
    max-degree colours (König, alternating-path recolouring); other graphs greedily.  Raises if
    `k` colours do not suffice for the method used."""
    nv = g.nv
    uv = g.edge_uv()
    delta = max((len(x) for x in g.incident), default=0)
    colour_of = [-1] * g.ne
    # at[v][c] = edge id using colour c at vertex v
    ncol = max(delta, 1)
    if g.bipartition() is not None:
        at = [[-1] * ncol for _ in range(nv)]
        for e, (u, v) in enumerate(uv):
            cu = next(c for c in range(ncol) if at[u][c] < 0)
            cv = next(c for c in range(ncol) if at[v][c] < 0)
            if cu != cv:
                # flip the cu/cv alternating path starting at v so that cu becomes free at v
                path = []
                x, c = v, cu
                while at[x][c] >= 0:
                    ee = at[x][c]
                    path.append(ee)
                    a, b = uv[ee]
                    x = b if a == x else a
                    c = cv if c == cu else cu
                for ee in path:
                    a, b = uv[ee]
                    old = colour_of[ee]
                    at[a][old] = -1
                    at[b][old] = -1
                for ee in path:
                    a, b = uv[ee]
                    new = cv if colour_of[ee] == cu else cu
                    colour_of[ee] = new
                    at[a][new] = ee
                    at[b][new] = ee
            colour_of[e] = cu
            at[u][cu] = e
            at[v][cu] = e
        used = ncol
    else:
        used = 0
        at_set = [set() for _ in range(nv)]
        for e, (u, v) in enumerate(uv):
            c = 0
            while c in at_set[u] or c in at_set[v]:
                c += 1
            colour_of[e] = c
            at_set[u].add(c)
            at_set[v].add(c)
            used = max(used, c + 1)
    if used > k:
        raise ValueError(f"edge_color: needs {used} colours, only {k} allowed")
    groups: List[List[Tuple[Vertex, Vertex]]] = [[] for _ in range(used)]
    for e, c in enumerate(colour_of):
        groups[c].append(g.edges[e])
    return [grp for grp in groups if grp]


# ---------------------------------------------------------------------------------------------
# BP edge schedules
# ---------------------------------------------------------------------------------------------

def forest_cover_edge_sequence(g: NamedGraph) -> List[Tuple[Vertex, Vertex]]:
    """Default BP schedule of the reference (`beliefpropagationcache.jl:27-29` →
    NamedGraphs.GraphsExtensions.forest_cover_edge_sequence, un-vendored): peel spanning forests
    off the graph until every edge is covered; for each tree emit the post-order DFS edges
    directed leaf→root followed by the same edges reversed root→leaf.  Every directed edge occurs
    exactly once.  Spanning trees are BFS trees rooted at the lowest-index vertex of each component
    (restated from upstream knowledge; the exact tree choice of NamedGraphs cannot be checked here
    and only changes the order, which callers may override through `edge_sequence=` anyway)."""
    nv = g.nv
    remaining = set(range(g.ne))
    seq: List[Tuple[Vertex, Vertex]] = []
    names = g.vertex_names
    while remaining:
        visited = [False] * nv
        for root in range(nv):
            if visited[root]:
                continue
            # BFS spanning tree of the component of `root` in the remaining-edge graph
            visited[root] = True
            children: Dict[int, List[int]] = {root: []}
            used_edges: List[int] = []
            dq = deque([root])
            while dq:
                i = dq.popleft()
                for e, j in g.incident[i]:
                    if e in remaining and not visited[j]:
                        visited[j] = True
                        children[i].append(j)
                        children[j] = []
                        used_edges.append(e)
                        dq.append(j)
            if not used_edges:
                continue
            # post-order DFS edges child→parent
            up: List[Tuple[int, int]] = []
            stack = [(root, iter(children[root]))]
            while stack:
                node, it = stack[-1]
                nxt = next(it, None)
                if nxt is None:
                    stack.pop()
                    if stack:
                        up.append((node, stack[-1][0]))
                else:
                    stack.append((nxt, iter(children[nxt])))
            remaining.difference_update(used_edges)
            seq.extend((names[a], names[b]) for a, b in up)
            seq.extend((names[b], names[a]) for a, b in reversed(up))
    return seq


def bipartite_edge_sequence(g: NamedGraph) -> List[Tuple[Vertex, Vertex]]:
    """A legal `edge_sequence=` for `update` (`beliefpropagationcache.jl:66`) whose sequential
    (Gauss–Seidel) semantics need only as many dependency levels as there are vertex colours: all
    messages leaving colour-0 vertices, then all leaving colour-1 vertices, ...  On a bipartite
    lattice that is two fully parallel levels.  Non-bipartite graphs get a greedy vertex colouring."""
    col = g.bipartition()
    if col is None:
        col = [-1] * g.nv
        for i in range(g.nv):
            used = {col[j] for _, j in g.incident[i] if col[j] >= 0}
            c = 0
            while c in used:
                c += 1
            col[i] = c
    ncol = max(col) + 1 if col else 0
    seq = []
    for c in range(ncol):
        for i in range(g.nv):
            if col[i] == c:
                for _, j in g.incident[i]:
                    seq.append((g.vertex_names[i], g.vertex_names[j]))
    return seq
