"""torchrun --nproc-per-node 2 tests/mgpu_check.py : sharded run vs the CPU oracle (small lattice)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
import torch.distributed as dist
import tnqs_b200 as tq
from oracle import tnqs_oracle as orc
from helpers import circuit_for_oracle, seq_idx, tfim_layer, Z

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for dtype, tol in ((np.complex128, 1e-9), (np.complex64, 2e-4)):
    g = tq.named_grid((4, 4))
    layer = tfim_layer(g)
    seq = tq.bipartite_edge_sequence(g)
    kw = dict(maxdim=4, cutoff=1e-12, normalize_tensors=True)
    bp = dict(maxiter=200, tolerance=1e-13 if dtype == np.complex128 else 1e-9, edge_sequence=seq)
    psi = tq.BeliefPropagationCache(tq.zerostate(dtype, g), device=local)
    tq.shard(psi)
    c = orc.product_state(g.nv, g.edge_uv(), [(1.0, 0.0)] * g.nv, np.complex128)
    gm, gv = circuit_for_oracle(g, layer)
    for l in range(4):
        psi, errs = tq.apply_gates(layer, psi, apply_kwargs=kw, bp_update_kwargs=bp)
        c, oerrs, _ = orc.apply_gates(c, gm, gv, seq_idx(g, seq), kw, dict(maxiter=200, tolerance=1e-13))
        zs = np.array(tq.expect(psi, [("Z", [v]) for v in g.vertices()]))
        zo = np.array([orc.expect_local(c, i, Z) for i in range(g.nv)])
        dz, de = float(np.max(np.abs(zs - zo))), float(np.max(np.abs(errs - oerrs)))
        same_bonds = list(psi.bond_dims()) == c.bond_dims()
        good = dz < tol and de < tol * max(1.0, float(np.max(oerrs)) / 1e-3) and same_bonds
        ok = ok and good
        if rank == 0:
            print(f"{np.dtype(dtype).name} layer {l+1}: max|dZ| {dz:.2e} max|derr| {de:.2e} bonds_equal {same_bonds} -> {'ok' if good else 'FAIL'}", flush=True)
    tns = psi.network()  # gathered from the owners
    if rank == 0:
        print("network() gathered", len(tns.tensors), "site tensors; owner map", psi.owner, flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MGPU CHECK", "PASSED" if int(flag) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)
