"""Mint the golden vectors of tests/golden/ with the CPU oracle (oracle/tnqs_oracle.py).

The reference is Julia and cannot run in this image, and its own tests hold no numeric goldens for this path
(SURVEY.md §8c), so these vectors pin the ORACLE (regression) and give the device path a fixed target; they are
not outputs of the reference.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tnqs_b200 as tq  # noqa: E402  (host-side graph / gate utilities only: no device call)
from oracle import tnqs_oracle as orc  # noqa: E402
from helpers_cpu import tfim_layer_cpu  # noqa: E402

Z = np.diag([1.0, -1.0])


def config1(nlayers=4):
    """BASELINE config 1: examples/2dIsing_dynamics.jl (5×5, dt=.25, hx=1, hz=.8, J=.5), maxdim=4, ComplexF64,
    cutoff=1e-10, normalize_tensors=false, default BP kwargs and forest-cover edge sequence."""
    g = tq.named_grid((5, 5))
    layer, gm, gv = tfim_layer_cpu(g)
    seq = [(g.index[a], g.index[b]) for a, b in tq.forest_cover_edge_sequence(g)]
    kw = dict(maxdim=4, cutoff=1e-10, normalize_tensors=False)
    c = orc.product_state(g.nv, g.edge_uv(), [(1.0, 0.0)] * g.nv, np.complex128)
    out = {"config": "5x5 TFIM, dt=0.25 hx=1 hz=0.8 J=0.5, maxdim=4, cutoff=1e-10, normalize_tensors=false, complex128",
           "layers": []}
    for _ in range(nlayers):
        c, errs, _ = orc.apply_gates(c, gm, gv, seq, kw)
        out["layers"].append({
            "maxvirtualdim": int(c.maxvirtualdim()),
            "max_trunc_err": float(np.max(errs)),
            "sum_trunc_err": float(np.sum(errs)),
            "trunc_err": [float(x) for x in errs],
            "sz_center": float(np.real(orc.expect_local(c, g.index[(3, 3)], Z))),
            "sz_all": [float(np.real(orc.expect_local(c, i, Z))) for i in range(g.nv)],
        })
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1_5x5_tfim_maxdim4_c128.json")
    with open(path, "w") as f:
        json.dump(config1(), f, indent=1)
    print("wrote", path)
