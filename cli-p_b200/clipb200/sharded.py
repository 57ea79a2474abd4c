"""Database-sharded exact search, one process per GPU.

Rows are split contiguously: rank r owns global ids [base_r, base_r + n_r).  A
query batch is replicated (2 KB/query), every rank scans its own shard and
produces a sorted local top-k with GLOBAL ids, then ONE collective moves the
per-rank results (each rank's D block and I block packed back to back in one
byte buffer) and the merge kernel picks the global top-k by (-score, id), which
makes the answer bit-identical to the single-GPU answer.

torch.distributed is plumbing here (NCCL over NVLink on the GPU box, gloo in the
CPU tests); `local_search` / `merge` are injectable so the host logic can be
exercised without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank` (rank r owns
    [r*ceil(N/R), min(N, (r+1)*ceil(N/R))))."""
    per = -(-n // world) if world > 0 else n
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def packed_bytes(nq: int, k: int) -> Tuple[int, int]:
    """(offset of the I block, total bytes) of one rank's packed result."""
    d_bytes = (nq * k * 4 + 7) // 8 * 8
    return d_bytes, d_bytes + nq * k * 8


class DistributedFlatIP:
    def __init__(self, index=None, group=None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 device: Optional[torch.device] = None):
        self.index = index
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._local_search = local_search or self._native_search
        self._merge = merge or self._native_merge
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.id_base = 0
        self.ntotal_global = 0
        self._buf = None
        self._gather = None
        # Opt-in (CLIPB200_SEARCH_GRAPHS=1): replay small-nq searches as one CUDA graph per (nq, k).
        # Single-GPU gain is small (1.533 -> 1.529 ms); with NCCL inside the capture a 2-rank run
        # hung in round 1, so it is restricted to world == 1 until that is understood.
        self.use_graphs = os.environ.get("CLIPB200_SEARCH_GRAPHS", "0") == "1" and self.world == 1
        self._graphs = {}

    # ---- ingest ------------------------------------------------------------------
    def finalize(self, n_local: Optional[int] = None) -> None:
        """Exchange shard sizes once so every rank knows its global id base."""
        if n_local is None:
            n_local = self.index.ntotal
        if self.world == 1:
            self.id_base, self.ntotal_global = 0, n_local
            return
        t = torch.tensor([n_local], dtype=torch.int64, device=self.device)
        sizes = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(sizes, t, group=self.group)
        sizes = [int(s.item()) for s in sizes]
        self.id_base = sum(sizes[:self.rank])
        self.ntotal_global = sum(sizes)

    # ---- search --------------------------------------------------------------------
    def _native_search(self, q, k, D, I, id_base):
        from . import _native as N
        sh = self.index._shards[0]
        stream = torch.cuda.current_stream(q.device).cuda_stream
        N.check(N.lib().cb_flatip_search_device(sh.handle, q.shape[0], C.c_void_p(q.data_ptr()), k,
                                                C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()),
                                                id_base, C.c_void_p(stream)))

    def _native_merge(self, gathered, R, nq, k, off_I, stride_bytes):
        from .faiss import merge_topk_device
        Dv = gathered.view(torch.float32)
        Iv = gathered[off_I:].view(torch.int64)
        return merge_topk_device(Dv, Iv, k, shard_stride_D=stride_bytes // 4,
                                 shard_stride_I=stride_bytes // 8, R=R, nq=nq)

    def _buffers(self, nq: int, k: int):
        off_I, total = packed_bytes(nq, k)
        if self._buf is None or self._buf.numel() != total:
            self._buf = torch.empty(total, dtype=torch.uint8, device=self.device)
            self._gather = torch.empty(total * self.world, dtype=torch.uint8, device=self.device)
        return off_I, total

    def search(self, q: torch.Tensor, k: int):
        """q: (nq, d) float32 on this rank's device, identical on all ranks.
        Returns (D, I) on every rank."""
        nq = q.shape[0]
        if (self.use_graphs and q.is_cuda and nq < 16 and self._local_search == self._native_search
                and self.ntotal_global > 0):
            out = self._search_graphed(q, k)
            if out is not None:
                return out
        return self._search_eager(q, k)

    def _search_graphed(self, q: torch.Tensor, k: int):
        key = (q.shape[0], k, q.shape[1])
        ent = self._graphs.get(key)
        if ent is None:
            try:
                q_static = torch.empty_like(q)
                q_static.copy_(q)
                for _ in range(2):                      # warm up: workspace allocation, func attributes, NCCL
                    self._search_eager(q_static, k)
                torch.cuda.synchronize(q.device)
                # the graph owns its packed result / gather buffers (the eager ones are re-sized freely)
                off_I, total = packed_bytes(q.shape[0], k)
                bufs = (torch.empty(total, dtype=torch.uint8, device=self.device),
                        torch.empty(total * self.world, dtype=torch.uint8, device=self.device))
                self._search_eager(q_static, k, bufs)
                torch.cuda.synchronize(q.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    D, I = self._search_eager(q_static, k, bufs)
                ent = (g, q_static, D, I, bufs)
            except Exception as e:                      # capture not possible here: stay eager, say so once
                import sys
                print(f"clipb200: CUDA-graph capture of the sharded search failed ({e}); running eagerly",
                      file=sys.stderr)
                self.use_graphs = False
                return None
            self._graphs[key] = ent
        g, q_static, D, I = ent[:4]
        q_static.copy_(q)
        g.replay()
        return D, I          # graph-owned outputs: valid until the next search with the same (nq, k)

    def _search_eager(self, q: torch.Tensor, k: int, bufs=None):
        nq = q.shape[0]
        if bufs is None:
            off_I, total = self._buffers(nq, k)
            buf, gather = self._buf, self._gather
        else:
            off_I, total = packed_bytes(nq, k)
            buf, gather = bufs
        D = buf[:nq * k * 4].view(torch.float32).view(nq, k)
        I = buf[off_I:].view(torch.int64).view(nq, k)
        self._local_search(q, k, D, I, self.id_base)
        if self.world == 1:
            return D.clone(), I.clone()
        if gather.is_cuda:
            dist.all_gather_into_tensor(gather, buf, group=self.group)
        else:  # gloo (CPU tests)
            parts = list(gather.view(self.world, total).unbind(0))
            dist.all_gather(parts, buf, group=self.group)
        return self._merge(gather, self.world, nq, k, off_I, total)
