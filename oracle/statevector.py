"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Dense state-vector simulator, independent of the
tensor-network oracle, used to pin it: simple update without truncation is exact
(/root/reference/src/Apply/simple_update.jl:4) and BP is exact on trees
(/root/reference/test/test_beliefpropagation.jl:44-54)."""
from __future__ import annotations

import numpy as np


class StateVector:
    def __init__(self, local_states):
        psi = np.ones((), dtype=np.complex128)
        for s in local_states:
            psi = np.tensordot(psi, np.asarray(s, dtype=np.complex128), axes=0)
        self.psi = psi  # shape (d,)*n

    @property
    def n(self):
        return self.psi.ndim

    def apply(self, gate, verts):
        gate = np.asarray(gate, dtype=np.complex128)
        k = len(verts)
        d = self.psi.shape[verts[0]]
        g = gate.reshape((d,) * (2 * k))  # [out..., in...]
        psi = np.tensordot(g, self.psi, axes=(list(range(k, 2 * k)), list(verts)))
        self.psi = np.moveaxis(psi, list(range(k)), list(verts))

    def expect(self, op, verts):
        op = np.asarray(op, dtype=np.complex128)
        k = len(verts)
        d = self.psi.shape[verts[0]]
        g = op.reshape((d,) * (2 * k))
        phi = np.tensordot(g, self.psi, axes=(list(range(k, 2 * k)), list(verts)))
        phi = np.moveaxis(phi, list(range(k)), list(verts))
        return np.vdot(self.psi, phi) / np.vdot(self.psi, self.psi)

    def rdm(self, v):
        m = np.moveaxis(self.psi, v, 0).reshape(self.psi.shape[v], -1)
        return m @ m.conj().T
