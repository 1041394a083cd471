"""NumPy prototype of the DEVICE formulation of simple update (Gram-based reduced factors), using the
same index conventions as csrc/kernels_small.cuh, checked against the oracle.  Dev tool only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tnqs_b200 as tq
from oracle import tnqs_oracle as orc
from helpers import ragged_state, random_psd_messages, oracle_from_tns, circuit_for_oracle, state_overlap


def absorb(t, m, axis):
    return np.moveaxis(np.tensordot(t, m, axes=([axis], [0])), -1, axis)


def msg_eig(M, cutoff):
    H = 0.5 * (M.astype(complex) + M.astype(complex).conj().T)
    lam, V = np.linalg.eigh(H)
    A = H @ V
    lam2 = np.real(np.sum(V.conj() * A, axis=0))
    keep = ~((lam2 == 0) | (np.abs(lam2) < cutoff))
    f = np.where(keep, np.sqrt(np.maximum(lam2, 0)), 0)
    S = (V * f) @ V.conj().T
    P = (V * keep) @ V.conj().T
    return S, P, keep.all()


def device_su(gate, T, pos, envs, maxdim, cutoff, normalize, sqrt_cutoff):
    """T = [T0, T1] (d, legs…), pos = shared-bond leg position (0-based among bond legs) per site,
    envs[s] = {leg position: message}"""
    d = [T[0].shape[0], T[1].shape[0]]
    chi = T[0].shape[1 + pos[0]]
    Tt, TP, GV, sq, isq = [], [], [], [], []
    for s in range(2):
        t, tp = T[s], T[s]
        for p, M in envs[s].items():
            S, P, allk = msg_eig(M, sqrt_cutoff)
            t = absorb(t, S, 1 + p)
            if not allk:
                tp = absorb(tp, P, 1 + p)
        Tt.append(t); TP.append(tp)
        # Gram with planes: A[j=(p,l)][col]; out[i][j] = Σ conj(X[i]) Y[j]
        a = np.moveaxis(t, 1 + pos[s], 1).reshape(d[s] * chi, -1)
        G = a.conj() @ a.T
        H = 0.5 * (G + G.conj().T)
        lam, V = np.linalg.eigh(H)
        lam = np.real(np.sum(V.conj() * (H @ V), axis=0))
        kept = (lam > 64 * 2.2e-16 * lam.max()) & (lam > 0)
        sq.append(np.where(kept, np.sqrt(np.abs(lam)), 0)); isq.append(np.where(kept, 1 / np.sqrt(np.abs(lam) + (~kept)), 0))
        GV.append(V)
    n0, n1 = d[0] * chi, d[1] * chi
    # R_s[r,(s,b)] = sq[r] conj(V[(s,b), r])
    R0 = (sq[0][:, None] * GV[0].conj().T).reshape(n0, d[0], chi)
    R1 = (sq[1][:, None] * GV[1].conj().T).reshape(n1, d[1], chi)
    W = np.einsum("asb,ctb->asct", R0, R1)
    g4 = gate.reshape(d[0], d[1], d[0], d[1])
    th = np.einsum("xyst,asct->axcy", g4, W).reshape(n0 * d[0], n1 * d[1])   # rows (r0,s0'), cols (r1,s1')
    U, S, Vh = np.linalg.svd(th, full_matrices=False)
    ext = [T[s].size // (d[s] * chi) for s in range(2)]
    full = min(min(ext[0], n0) * d[0], min(ext[1], n1) * d[1])   # what the reference's thin QR leaves
    keep, err = orc.truncate_spectrum(S[:full] ** 2, maxdim, cutoff, 1)
    US = U * S   # columns σ u  (what the Jacobi leaves)
    sig = S[:keep]
    L = US[:, :keep] / np.sqrt(sig)                      # (r0,s0') × c
    Rp = (th.T @ US[:, :keep].conj()) / sig ** 1.5       # Rp[κ,c] = Σ_ρ θ[ρ,κ] conj(Uσ[ρ,c]) / σ^1.5
    X0 = np.einsum("ir,rpc->ipc", GV[0] * isq[0], L.reshape(n0, d[0], keep)).reshape(n0, d[0] * keep)
    X1 = np.einsum("ir,rpc->ipc", GV[1] * isq[1], Rp.reshape(n1, d[1], keep)).reshape(n1, d[1] * keep)
    out = []
    for s, Xm in ((0, X0), (1, X1)):
        tp = TP[s]
        # Out[p',o,c,n] = Σ_{p,b} In[p,o,b,n] Mat[(p,b),(p',c)]
        a = np.moveaxis(tp, 1 + pos[s], 1)                # [p, b, rest…]
        rest = a.shape[2:]
        res = np.tensordot(Xm.reshape(d[s], chi, d[s], keep), a, axes=([0, 1], [0, 1]))  # [p', c, rest]
        out.append(np.moveaxis(res, 1, 1 + pos[s]))
    sv = sig.copy()
    if normalize:
        sv = sv / np.linalg.norm(sv)
        out = [t / np.linalg.norm(t) for t in out]
    return out, sv, err


if __name__ == "__main__":
    g = tq.named_grid((3, 2))
    dims = [2, 3, 2, 3, 2, 3, 2][:g.ne]
    worst = 0
    for dtype in (np.complex128, np.complex64):
        psi = ragged_state(g, dims, dtype, seed=11)
        ms = random_psd_messages(g, dims, dtype, seed=12)
        for gate_name in ("Rzz", "CNOT", "Rxxyy"):
            for e_id in range(g.ne):
                a, b = g.edges[e_id]
                for maxdim, cutoff in ((None, None), (3, 1e-12), (2, None)):
                    c = oracle_from_tns(psi)
                    for (x, y), m in ms.items():
                        c.msg[(g.index[x], g.index[y])] = m
                    circ = [(gate_name, [a, b], 0.37)] if gate_name.startswith("R") else [(gate_name, [a, b])]
                    gm, gv = circuit_for_oracle(g, circ)
                    c0 = c.copy()
                    c, oerrs, _ = orc.apply_gates(c, gm, gv, [], dict(maxdim=maxdim, cutoff=cutoff, normalize_tensors=True), update_cache=False)
                    ia, ib = g.index[a], g.index[b]
                    pos = [g.leg_of(ia, e_id), g.leg_of(ib, e_id)]
                    envs = []
                    for v, o in ((ia, ib), (ib, ia)):
                        envs.append({p: c0.message(w, v) for p, (ee, w) in enumerate(g.incident[v]) if w != o})
                    eps = np.finfo(np.float32 if dtype == np.complex64 else np.float64).eps
                    out, sv, err = device_su(gm[0], [c0.T[ia].astype(complex), c0.T[ib].astype(complex)], pos, envs, maxdim, cutoff, True, 10 * eps)
                    c2 = c0.copy()
                    c2.T[ia], c2.T[ib] = out[0].astype(complex), out[1].astype(complex)
                    c2.dtype = np.dtype(complex)
                    ov, n1, n2 = state_overlap(c2, c)
                    dsv = np.max(np.abs(sv - np.diag(c.msg[(ia, ib)]).real))
                    worst = max(worst, abs(ov - 1), abs(n1 / n2 - 1), dsv if dtype == np.complex128 else 0, abs(err - oerrs[0]))
                    tol = 1e-9 if dtype == np.complex128 else 1e-4
                    assert abs(ov - 1) < tol and abs(n1 / n2 - 1) < tol and dsv < tol and abs(err - oerrs[0]) < tol, (dtype, gate_name, e_id, maxdim, ov, n1 / n2, dsv, err, oerrs[0])
    print("prototype of the device SU formulation matches the oracle; worst deviation", worst)
