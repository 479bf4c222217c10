"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL/NVLink): the index is built once on
rank 0 and broadcast to the peers; queries are sharded, there is no steady-state collective.
torch is used only for the communicator and as a typed view of device memory owned by libcgx_b200.so."""
from __future__ import annotations

import numpy as np

from ._lib import IndexArrays


def shard_queries(qry_off, world: int, rank: int):
    """Contiguous, token-balanced query ranges (sentence-granular).  Returns (q_begin, q_end)."""
    qry_off = np.asarray(qry_off, dtype=np.int64)
    Q = len(qry_off) - 1
    T = int(qry_off[-1])
    if world <= 1 or Q == 0:
        return 0, Q
    targets = [T * r // world for r in range(world + 1)]
    cuts = [int(np.searchsorted(qry_off, t, side="left")) for t in targets]
    cuts[0], cuts[-1] = 0, Q
    for i in range(1, world + 1):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts[rank], cuts[rank + 1]


class _DevView:
    """__cuda_array_interface__ wrapper of a raw device pointer (uint8 view)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = dict(shape=(nbytes,), typestr="|u1", data=(ptr, False), version=2)


def device_tensor(ptr: int, nbytes: int, device):
    import torch
    return torch.as_tensor(_DevView(ptr, nbytes), device=device)


def broadcast_shape(shape_vec, src: int = 0):
    """shape_vec: int64 tensor [4 + 100] = (n, m, lex_count, max_token, freq_list...) on the communicator's device."""
    import torch.distributed as dist
    dist.broadcast(shape_vec, src=src)
    return shape_vec


def broadcast_index(ex, src: int = 0):
    """Broadcast rank `src`'s built index to every rank's GrammarExtractor `ex` (NCCL)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    dev = torch.device("cuda", ex.device)
    vec = torch.zeros(104, dtype=torch.int64, device=dev)
    if rank == src:
        a = ex.export_index()
        vec[:4] = torch.tensor([a.n, a.m, a.lex_count, a.max_token], dtype=torch.int64)
        vec[4:] = torch.tensor(list(a.freq_list), dtype=torch.int64)
    broadcast_shape(vec, src)
    if rank != src:
        shape = IndexArrays()
        v = vec.cpu().tolist()
        shape.n, shape.m, shape.lex_count, shape.max_token = int(v[0]), int(v[1]), int(v[2]), int(v[3])
        for i in range(100):
            shape.freq_list[i] = int(v[4 + i])
        a = ex.alloc_index(shape)
    nbytes = 0
    for name, _, _ in IndexArrays.ARRAYS:
        nb = a.nbytes(name)
        if nb == 0:
            continue
        t = device_tensor(getattr(a, name), nb, dev)
        dist.broadcast(t, src=src)
        nbytes += nb
    torch.cuda.synchronize(dev)
    if rank != src:
        ex.commit_index()
    return nbytes
