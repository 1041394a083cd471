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

        # gate batch (engine.cu, step 4): gate k of a vertex-disjoint batch is factorised on rank k mod R; the
        # results travel in rank-major records — slot(k) = (k mod R)·per + k div R — through ONE all-gather of
        # equal-sized byte buffers, and every rank applies all of them.  Emulated with the oracle's apply_gate.
        import torch
        batch = [tuple(p) for p in tq.edge_color(g, 4)[0]]
        gate = tq.gate_matrix("Rzz", 2, 0.3)
        kw = dict(maxdim=2, cutoff=1e-12, normalize_tensors=True)
        one = c.copy()
        ref_errs = [orc.apply_gate(one, gate, [g.index[a], g.index[b]], **kw) for (a, b) in batch]
        ng, R = len(batch), world
        per = (ng + R - 1) // R
        slot = lambda k: (k % R) * per + k // R
        maxel = max(one.T[v].size for v in range(g.nv))
        maxchi = max(one.bond_dims())
        rec = 2 + 2 * 2 * maxel + maxchi  # err, keep, two tensors (re,im) padded, sigma padded
        mine = np.zeros((per, rec))
        for k, (a, b) in enumerate(batch):
            if k % R != rank:
                continue
            w = c.copy()
            ia, ib = g.index[a], g.index[b]
            e = orc.apply_gate(w, gate, [ia, ib], **kw)
            sig = np.real(np.diag(w.msg[(ia, ib)]))
            row = mine[k // R]
            row[0], row[1] = e, len(sig)
            for j, t in enumerate((w.T[ia], w.T[ib])):
                flat = t.reshape(-1).view(np.float64)
                row[2 + j * 2 * maxel: 2 + j * 2 * maxel + flat.size] = flat
            row[2 + 4 * maxel: 2 + 4 * maxel + len(sig)] = sig
        buf = torch.from_numpy(mine.copy())
        parts = [torch.zeros_like(buf) for _ in range(R)]
        dist.all_gather(parts, buf)
        allrec = torch.cat(parts).numpy()
        got = c.copy()
        got_errs = []
        for k, (a, b) in enumerate(batch):
            row = allrec[slot(k)]
            ia, ib = g.index[a], g.index[b]
            keep = int(row[1])
            got_errs.append(row[0])
            for j, v in enumerate((ia, ib)):
                shp = list(got.T[v].shape)
                shp[got.leg(v, ib if v == ia else ia)] = keep
                n = int(np.prod(shp))
                got.T[v] = row[2 + j * 2 * maxel: 2 + j * 2 * maxel + 2 * n].copy().view(np.complex128).reshape(shp)
            sig = row[2 + 4 * maxel: 2 + 4 * maxel + keep]
            got.msg[(ia, ib)] = np.diag(sig).astype(np.complex128)
            got.msg[(ib, ia)] = np.diag(sig).astype(np.complex128)
        err2 = max(float(np.max(np.abs(got.T[v] - one.T[v]))) for v in range(g.nv))
        err2 = max(err2, float(np.max(np.abs(np.array(got_errs) - np.array(ref_errs)))))
        err2 = max(err2, max(float(np.max(np.abs(got.msg[k2] - one.msg[k2]))) for k2 in one.msg))
        # level exchange (engine.cu, Engine::exchange): every item (a message of the level, a Gram matrix of a batch) has a
        # root rank; the replicated item list gives every rank the same packing — the items of root r, in list order, in
        # slots r·per … of one staging buffer; each rank packs what it owns, ONE all-gather moves everything, each rank
        # unpacks what it does not own.  Items of different sizes share the slot size of the largest one.
        rng = np.random.default_rng(7)
        sizes = [int(x) for x in rng.integers(3, 40, size=23)]
        roots = [int(x) for x in rng.integers(0, R, size=23)]
        truth = [np.random.default_rng(100 + i).standard_normal(n) for i, n in enumerate(sizes)]
        have = [truth[i].copy() if roots[i] == rank else np.full(sizes[i], np.nan) for i in range(23)]
        slot_n = max(sizes)
        cnt = [roots.count(r) for r in range(R)]
        per_x = max(cnt)
        stage = np.zeros((R, per_x, slot_n))
        nxt = [0] * R
        where = []
        for i in range(23):
            where.append((roots[i], nxt[roots[i]]))
            nxt[roots[i]] += 1
            if roots[i] == rank:
                stage[where[i][0], where[i][1], :sizes[i]] = have[i]
        mine_t = torch.from_numpy(stage[rank].copy())
        parts = [torch.zeros_like(mine_t) for _ in range(R)]
        dist.all_gather(parts, mine_t)
        for i in range(23):
            if roots[i] != rank:
                have[i] = parts[where[i][0]][where[i][1], :sizes[i]].numpy().copy()
        err3 = max(float(np.max(np.abs(have[i] - truth[i]))) for i in range(23))
        q.put((rank, max(err, err2, err3), len(tq.cut_edges(g, owner))))
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
