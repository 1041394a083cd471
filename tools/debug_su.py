import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tnqs_b200 as tq
from oracle import tnqs_oracle as orc
from helpers import *

g = tq.named_grid((3, 2))
for dtype in (np.complex128,):
  for dimname, dims in (("uni2", [2]*7), ("uni3", [3]*7), ("uni4", [4]*7), ("ragged", [2,3,2,3,2,3,2])):
    psi = ragged_state(g, dims, dtype, seed=11)
    ms = random_psd_messages(g, dims, dtype, seed=12)
    for with_msgs in (False, True):
      for e_id in range(g.ne):
        a, b = g.edges[e_id]
        bpc = tq.BeliefPropagationCache(psi)
        c = oracle_from_tns(psi)
        if with_msgs:
            bpc.setmessages(list(ms), list(ms.values()))
            for (x, y), m in ms.items():
                c.msg[(g.index[x], g.index[y])] = m
        circ = [("Rzz", [a, b], 0.37)]
        kw = dict(normalize_tensors=True)
        out, errs = tq.apply_gates(circ, bpc, apply_kwargs=kw, update_cache=False)
        gm, gv = circuit_for_oracle(g, circ)
        c, oerrs, _ = orc.apply_gates(c, gm, gv, [], kw, update_cache=False)
        sd = np.diag(out.message((a, b))).real; so = np.diag(c.msg[(g.index[a], g.index[b])]).real
        n = min(len(sd), len(so))
        ov, n1, n2 = state_overlap(oracle_from_bpc(out), c)
        ia, ib = g.index[a], g.index[b]
        print(dimname, "msgs" if with_msgs else "nomsg", "edge", e_id, "pos", g.leg_of(ia, e_id), g.leg_of(ib, e_id), "deg", g.degree(a), g.degree(b),
              "dimsdev/orc", len(sd), len(so), "max|dsigma| %.2e" % np.max(np.abs(sd[:n] - so[:n])), "1-ov %.2e" % abs(1 - ov), "n1/n2-1 %.2e" % abs(n1/n2 - 1))
