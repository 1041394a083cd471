"""world_size-2 gloo tests (CPU) of the N>1 host logic: the vertex partition, the unique-id
hand-off, and the exchange protocol of SURVEY.md §8e — each rank updates only the messages that
leave vertices it owns and receives the others — emulated with the oracle as the compute and gloo as
the transport, and compared with the single-process oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import tnqs_b200 as tq
    from tnqs_b200.distributed import broadcast_unique_id
    from oracle import tnqs_oracle as orc
    from helpers import ragged_state, oracle_from_tns, seq_idx

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = broadcast_unique_id(lambda: bytes(range(128)), rank, world)
        assert uid == bytes(range(128))
        g = tq.named_grid((4, 3))
        owner = tq.partition_vertices(g, world)
        dims = [2 + (e % 2) for e in range(g.ne)]
        psi = ragged_state(g, dims, np.complex128, seed=4)
        seq = seq_idx(g, tq.bipartite_edge_sequence(g))
        # reference: one process does everything
        ref, _ = orc.bp_update(oracle_from_tns(psi), seq, maxiter=4, tolerance=None)
        # sharded: dependency levels of the bipartite schedule = the two colour classes
        c = oracle_from_tns(psi)
        col = g.bipartition()
        for _ in range(4):
            for colour in (0, 1):
                level = [(u, v) for (u, v) in seq if col[u] == colour]
                mine = {(u, v): orc.updated_message(c, u, v) for (u, v) in level if owner[u] == rank}
                gathered = [None] * world
                dist.all_gather_object(gathered, mine)
                for part in gathered:
                    c.msg.update(part)
        err = max(float(np.max(np.abs(c.msg[k] - ref.msg[k]))) for k in ref.msg)
        q.put((rank, err, len(tq.cut_edges(g, owner))))
    finally:
        dist.destroy_process_group()


def test_partition_is_balanced_and_contiguous():
    sys.path.insert(0, ROOT)
    import tnqs_b200 as tq
    g = tq.named_grid((16, 16))
    for n in (1, 2, 4, 8):
        owner = tq.partition_vertices(g, n)
        assert len(owner) == g.nv and sorted(set(owner)) == list(range(n))
        assert owner == sorted(owner)
        counts = np.bincount(owner)
        assert counts.max() - counts.min() <= 1
        # row strips: 16 cut edges per strip boundary
        assert len(tq.cut_edges(g, owner)) == 16 * (n - 1)


def test_two_rank_exchange_protocol_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    res = [q.get(timeout=10) for _ in range(2)]
    for rank, err, ncut in res:
        assert err < 1e-14
        assert ncut == 5  # 4x3 grid split 6|6 vertices: 4 vertical + 1 horizontal cut edge
