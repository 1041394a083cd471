"""Multi-GPU sharding of one cache across the ranks of a `torch.distributed` job.

One process per GPU.  Every rank keeps the graph, the bond dimensions and all messages; the site
tensor of vertex v lives only on `owner[v]`.  The C library exchanges the O(χ²) data (the new
messages of a BP level, the reduced-factor Gram matrices of a gate batch) with NCCL over NVLink;
`torch.distributed` is only used here to hand every rank the ncclUniqueId.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .graphs import NamedGraph


def partition_vertices(g: NamedGraph, nranks: int) -> List[int]:
    """Owner rank per vertex index: contiguous, balanced blocks of the vertex order.  For
    `named_grid` (first coordinate fastest) these are strips of rows / slabs, so every rank talks to
    at most two neighbours; for other graphs it is an id-range split."""
    if nranks < 1:
        raise ValueError("nranks must be >= 1")
    nv = g.nv
    base, extra = divmod(nv, nranks)
    owner: List[int] = []
    for r in range(nranks):
        owner += [r] * (base + (1 if r < extra else 0))
    return owner


def cut_edges(g: NamedGraph, owner: Sequence[int]) -> List[int]:
    """Edge ids whose endpoints live on different ranks (the only ones whose messages / Gram matrices
    carry information across NVLink)."""
    return [e for e, (u, v) in enumerate(g.edge_uv()) if owner[u] != owner[v]]


def broadcast_unique_id(make_id, rank: int, world: int) -> bytes:
    """Rank 0 calls `make_id()` (128 bytes); everyone receives it through torch.distributed."""
    import torch.distributed as dist
    payload = [make_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(payload, src=0)
    uid = payload[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("bad NCCL unique id")
    return bytes(uid)


def _make_nccl_id() -> bytes:
    lib = _lib.load()
    buf = (C.c_char * 128)()
    _lib.check(lib.tnqs_comm_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf.raw)


def shard(bpc, owner: Optional[Sequence[int]] = None, rank: Optional[int] = None, world: Optional[int] = None):
    """Shard `bpc` (built identically on every rank) across the job: joins an NCCL communicator and
    drops the site tensors this rank does not own.  Returns `bpc`."""
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if owner is None:
        owner = partition_vertices(bpc.graph, world)
    owner = np.ascontiguousarray(owner, dtype=np.int32)
    if len(owner) != bpc.graph.nv:
        raise ValueError("owner must have one entry per vertex")
    uid = broadcast_unique_id(_make_nccl_id, rank, world) if world > 1 else bytes(128)
    _lib.check(bpc._lib.tnqs_comm_init(bpc._h, rank, world, C.c_char_p(uid),
                                       owner.ctypes.data_as(C.POINTER(C.c_int32))))
    bpc.owner, bpc.rank, bpc.world = [int(x) for x in owner], rank, world
    return bpc
