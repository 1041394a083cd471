"""Gate registry of the host mirror: circuit tuples `(name, vertices[, params])` → dense matrices.

Mirrors `/root/reference/src/Apply/gate_definitions.jl:21-64` (registry, qiskit θ→θ/2 rescale of
Rxx/Ryy/Rzz at `:49-51`), `:110-153` (`toitensor`: Pauli-string sugar, alias lookup, parameter
arity check, `ArgumentError` with suggestions) and `:189-239` (`register_gate!`,
`register_alias!`, `unregister_gate!` with locked built-ins).  The matrices are the ITensors.jl
"S=1/2"/"Qubit" `op` definitions (un-vendored dependency, restated from its published
conventions); two-site matrices use the `kron(first, second)` basis order, i.e. row/column index
`2*s1 + s2`.  Only matrices cross the C-ABI (host side only).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Sequence, Tuple

import numpy as np

_I2 = np.eye(2, dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)
_H = np.array([[1, 1], [1, -1]], dtype=complex) / math.sqrt(2)
_PAULI = {"X": _X, "Y": _Y, "Z": _Z, "I": _I2}


def _expm_herm(h: np.ndarray, t: float) -> np.ndarray:
    """exp(-i t h) for Hermitian h."""
    w, v = np.linalg.eigh(h)
    return (v * np.exp(-1j * t * w)) @ v.conj().T


def _rx(theta):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)


def _ry(theta):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -s], [s, c]], dtype=complex)


def _rz(theta):
    return np.diag([np.exp(-0.5j * theta), np.exp(0.5j * theta)]).astype(complex)


def _phase(phi):
    return np.diag([1.0, np.exp(1j * phi)]).astype(complex)


def _controlled(u):
    m = np.eye(4, dtype=complex)
    m[2:, 2:] = u
    return m


def _rpp(p):
    pp = np.kron(p, p)

    def f(phi):  # ITensors convention: exp(-i ϕ P⊗P)
        return math.cos(phi) * np.eye(4, dtype=complex) - 1j * math.sin(phi) * pp

    return f


_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)
_ISWAP = np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=complex)
_SQSWAP = np.array([[1, 0, 0, 0], [0, (1 + 1j) / 2, (1 - 1j) / 2, 0],
                    [0, (1 - 1j) / 2, (1 + 1j) / 2, 0], [0, 0, 0, 1]], dtype=complex)
_SQISWAP = np.array([[1, 0, 0, 0], [0, 1 / math.sqrt(2), 1j / math.sqrt(2), 0],
                     [0, 1j / math.sqrt(2), 1 / math.sqrt(2), 0], [0, 0, 0, 1]], dtype=complex)


def _xx_plus_yy(theta, beta):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[1, 0, 0, 0],
                     [0, c, -1j * s * np.exp(-1j * beta), 0],
                     [0, -1j * s * np.exp(1j * beta), c, 0],
                     [0, 0, 0, 1]], dtype=complex)


_HXXYY = 0.5 * (np.kron(_X, _X) + np.kron(_Y, _Y))
_HXXYYZZ = _HXXYY + 0.5 * np.kron(_Z, _Z)


class GateSpec:
    """`GateSpec` of `gate_definitions.jl:13-18`: matrix builder, its arity, the number of
    parameters, and the rescale applied to user parameters first."""

    def __init__(self, build: Callable[..., np.ndarray], nsites: int, nparams: int = 0,
                 rescale: Callable = lambda x: x):
        self.build, self.nsites, self.nparams, self.rescale = build, nsites, nparams, rescale


def _half(theta):
    return theta / 2


GATES: Dict[str, GateSpec] = {
    "X": GateSpec(lambda: _X.copy(), 1), "Y": GateSpec(lambda: _Y.copy(), 1),
    "Z": GateSpec(lambda: _Z.copy(), 1), "H": GateSpec(lambda: _H.copy(), 1),
    "Rx": GateSpec(_rx, 1, 1), "Ry": GateSpec(_ry, 1, 1), "Rz": GateSpec(_rz, 1, 1),
    "P": GateSpec(_phase, 1, 1),
    "CNOT": GateSpec(lambda: _controlled(_X), 2), "CX": GateSpec(lambda: _controlled(_X), 2),
    "CY": GateSpec(lambda: _controlled(_Y), 2), "CZ": GateSpec(lambda: _controlled(_Z), 2),
    "SWAP": GateSpec(lambda: _SWAP.copy(), 2), "iSWAP": GateSpec(lambda: _ISWAP.copy(), 2),
    "√SWAP": GateSpec(lambda: _SQSWAP.copy(), 2), "√iSWAP": GateSpec(lambda: _SQISWAP.copy(), 2),
    "Rxx": GateSpec(_rpp(_X), 2, 1, _half), "Ryy": GateSpec(_rpp(_Y), 2, 1, _half),
    "Rzz": GateSpec(_rpp(_Z), 2, 1, _half),
    "CRx": GateSpec(lambda t: _controlled(_rx(t)), 2, 1),
    "CRy": GateSpec(lambda t: _controlled(_ry(t)), 2, 1),
    "CRz": GateSpec(lambda t: _controlled(_rz(t)), 2, 1),
    "CPHASE": GateSpec(lambda p: _controlled(_phase(p)), 2, 1),
    "Rxxyy": GateSpec(lambda t: _expm_herm(_HXXYY, t), 2, 1),
    "Rxxyyzz": GateSpec(lambda t: _expm_herm(_HXXYYZZ, t), 2, 1),
    "xx_plus_yy": GateSpec(_xx_plus_yy, 2, 2),
}
# "Rz+" / "Rz+z+" are listed by the reference registry (`gate_definitions.jl:33,58`) but their
# matrices live in un-vendored ITensors code this build cannot see; they are left unregistered and
# raise the same ArgumentError an unknown name does.

BUILTIN_GATES = frozenset(GATES)
ALIASES: Dict[str, str] = {k.lower(): k for k in GATES if k.lower() != k}
ALIASES["cp"] = "CPHASE"


class ArgumentError(ValueError):
    """Stands in for Julia's `ArgumentError` (thrown at `gate_definitions.jl:139,148,195,213,231`)."""


def _levenshtein(a: str, b: str) -> int:
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def _resolve(name: str):
    spec = GATES.get(name)
    if spec is not None:
        return spec
    canon = ALIASES.get(name)
    return GATES.get(canon) if canon is not None else None


def register_gate(name: str, build: Callable[..., np.ndarray], nsites: int, nparams: int = 0,
                  rescale: Callable = lambda x: x) -> str:
    """`register_gate!` (`gate_definitions.jl:189-202`); built-ins are locked."""
    if name in BUILTIN_GATES:
        raise ArgumentError(f'"{name}" is a built-in gate and cannot be overwritten.')
    GATES[name] = GateSpec(build, nsites, nparams, rescale)
    return name


def register_alias(alias: str, canonical: str) -> str:
    """`register_alias!` (`gate_definitions.jl:212-220`)."""
    if canonical not in GATES:
        raise ArgumentError(
            f'Cannot register alias "{alias}" → "{canonical}": canonical gate is not registered.')
    ALIASES[alias] = canonical
    return alias


def unregister_gate(name: str) -> str:
    """`unregister_gate!` (`gate_definitions.jl:230-239`)."""
    if name in BUILTIN_GATES:
        raise ArgumentError(f'"{name}" is a built-in gate and cannot be unregistered.')
    GATES.pop(name, None)
    for a, c in list(ALIASES.items()):
        if c == name:
            del ALIASES[a]
    return name


def gate_matrix(name, nsites: int, params=None) -> np.ndarray:
    """Dense complex128 matrix (d^n × d^n, kron(first, second) order) of a circuit-tuple gate.
    An `np.ndarray` passed as `name` is returned as is (the reference's ITensor pass-through,
    `gate_definitions.jl:116`)."""
    if isinstance(name, np.ndarray):
        return np.asarray(name, dtype=complex)
    if len(name) > 1 and all(c in "XYZxyz" for c in name):
        if len(name) != nsites:
            raise ArgumentError(f'Pauli string "{name}" acts on {len(name)} sites, got {nsites}.')
        m = np.ones((1, 1), dtype=complex)
        for c in name:
            m = np.kron(m, _PAULI[c.upper()])
        return m
    spec = _resolve(name)
    if spec is None:
        lname = name.lower()
        scored = sorted(((_levenshtein(lname, g.lower()), g) for g in GATES))
        sugg = [g for d, g in scored if d <= 2][:3]
        msg = f'Unknown gate "{name}".'
        if sugg:
            msg += " Did you mean: " + ", ".join(f'"{s}"' for s in sugg) + "?"
        else:
            msg += f" Registered gates: {sorted(GATES)}."
        raise ArgumentError(msg)
    if spec.nparams == 0:
        m = spec.build()
    else:
        raw = spec.rescale(params)
        pvals = tuple(raw) if isinstance(raw, (tuple, list, np.ndarray)) else (raw,)
        if len(pvals) != spec.nparams:
            raise ArgumentError(
                f'Gate "{name}" expects {spec.nparams} parameter(s), got {len(pvals)}.')
        m = spec.build(*pvals)
    m = np.asarray(m, dtype=complex)
    if m.shape != (2 ** nsites, 2 ** nsites):
        raise ArgumentError(f'Gate "{name}" is a {int(math.log2(m.shape[0]))}-site gate, '
                            f'got {nsites} vertices.')
    return m


def observable_matrix(op: str) -> np.ndarray:
    """Single-site operator named by one character/string of an observable tuple
    (`/root/reference/src/expect.jl:159-181`): Paulis and the identity."""
    key = op.upper() if len(op) == 1 else op
    if key in _PAULI:
        return _PAULI[key].copy()
    spec = _resolve(op)
    if spec is not None and spec.nsites == 1 and spec.nparams == 0:
        return np.asarray(spec.build(), dtype=complex)
    raise ArgumentError(f'Unknown observable operator "{op}".')


STATES = {"↑": (1.0, 0.0), "↓": (0.0, 1.0), "Up": (1.0, 0.0), "Dn": (0.0, 1.0),
          "0": (1.0, 0.0), "1": (0.0, 1.0), "Z+": (1.0, 0.0), "Z-": (0.0, 1.0),
          "+": (2 ** -0.5, 2 ** -0.5), "-": (2 ** -0.5, -(2 ** -0.5)),
          "X+": (2 ** -0.5, 2 ** -0.5), "X-": (2 ** -0.5, -(2 ** -0.5))}
