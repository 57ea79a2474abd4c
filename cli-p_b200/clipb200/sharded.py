"""Database-sharded exact search, one process per GPU.

Rows are split contiguously: rank r owns global ids [base_r, base_r + n_r).  A
query batch is replicated (2 KB/query), every rank scans its own shard and
produces a sorted local top-k with GLOBAL ids, then the per-rank results travel ONCE
(by peer stores into rank 0's mailbox, or by one collective of the packed D | I blocks)
and the merge kernel picks the global top-k by (-score, id), which makes the answer
bit-identical to the single-GPU answer.

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
    """One rank's view of the sharded index.

    transport "p2p" (default on GPUs): the per-rank top-k lists are stored straight into a mailbox
    in rank 0's HBM over NVLink by the kernel that produces them, and rank 0's merge kernel waits on
    per-rank counters (include/clipb200.h, cb_flatip_*_p2p_*): one kernel chain per GPU, no
    collective, no Python between scan and answer.  The answer exists on rank 0 only.
    transport "nccl": ONE all_gather_into_tensor of the packed per-rank results, then the merge
    kernel on every rank (also the CPU/gloo test path, with injected local_search / merge).
    """

    def __init__(self, index=None, group=None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None,
                 device: Optional[torch.device] = None, transport: Optional[str] = None,
                 mailbox_elems: int = 128 * 1024):
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
        if transport is None:
            transport = os.environ.get("CLIPB200_SEARCH_TRANSPORT", "p2p")
        if local_search is not None or self.device.type != "cuda":
            transport = "nccl"            # injected CPU stand-ins: the collective path (gloo in the tests)
        assert transport in ("p2p", "nccl"), transport
        self.transport = transport
        self.mailbox_elems = int(mailbox_elems)
        self._p2p_ready = False

    # ---- ingest ------------------------------------------------------------------
    def finalize(self, n_local: Optional[int] = None) -> None:
        """Exchange shard sizes once so every rank knows its global id base; with the p2p
        transport also map rank 0's mailbox into every rank (cudaIpc handle, exchanged once)."""
        if n_local is None:
            n_local = self.index.ntotal
        if self.world == 1:
            self.id_base, self.ntotal_global = 0, n_local
        else:
            t = torch.tensor([n_local], dtype=torch.int64, device=self._plumbing_device())
            sizes = [torch.zeros_like(t) for _ in range(self.world)]
            dist.all_gather(sizes, t, group=self.group)
            sizes = [int(s.item()) for s in sizes]
            self.id_base = sum(sizes[:self.rank])
            self.ntotal_global = sum(sizes)
        if self.transport == "p2p" and self.world > 1 and not self._p2p_ready:
            self._p2p_attach()

    def _handle(self):
        return self.index._shards[0].handle

    def _plumbing_device(self):
        """Where the few bytes of set-up traffic live: NCCL moves CUDA tensors, gloo CPU tensors."""
        return self.device if dist.get_backend(self.group) == "nccl" else torch.device("cpu")

    def _p2p_attach(self) -> None:
        from . import _native as N
        raw = (C.c_ubyte * 64)()
        N.check(N.lib().cb_flatip_p2p_init(self._handle(), self.rank, self.world, self.mailbox_elems, raw))
        h = torch.tensor(list(raw), dtype=torch.uint8, device=self._plumbing_device())
        dist.broadcast(h, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        raw_root = (C.c_ubyte * 64)(*h.cpu().tolist())
        N.check(N.lib().cb_flatip_p2p_connect(self._handle(), raw_root))
        dist.barrier(group=self.group)          # nobody searches before every rank has mapped the mailbox
        self._p2p_ready = True

    def p2p_error(self) -> int:
        """Sticky mailbox error flag (a peer did not deliver in time).  Synchronises."""
        from . import _native as N
        e = C.c_int(0)
        N.check(N.lib().cb_flatip_p2p_status(self._handle(), C.byref(e)))
        return int(e.value)

    # ---- search --------------------------------------------------------------------
    def _native_search(self, q, k, D, I, id_base):
        from . import _native as N
        stream = torch.cuda.current_stream(q.device).cuda_stream
        N.check(N.lib().cb_flatip_search_device(self._handle(), q.shape[0], C.c_void_p(q.data_ptr()), k,
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
        Returns (D, I): on every rank with the nccl transport, on rank 0 only (None, None
        elsewhere) with the p2p transport."""
        nq = q.shape[0]
        if self.world == 1:
            D = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            self._local_search(q, k, D, I, self.id_base)
            return D, I
        if self.transport == "p2p":
            return self._search_p2p(q, k)
        return self._search_collective(q, k)

    def _search_p2p(self, q: torch.Tensor, k: int):
        from . import _native as N
        assert self._p2p_ready, "call finalize() first"
        nq = q.shape[0]
        D = I = None
        dp = ip = None
        if self.rank == 0:
            D = torch.empty((nq, k), dtype=torch.float32, device=self.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=self.device)
            dp, ip = C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr())
        stream = torch.cuda.current_stream(q.device).cuda_stream
        N.check(N.lib().cb_flatip_search_p2p_device(self._handle(), nq, C.c_void_p(q.data_ptr()), k, dp, ip,
                                                    self.id_base, C.c_void_p(stream)))
        return D, I

    def search_host(self, q, k: int):
        """numpy (nq, d) float32 in, numpy (D, I) out on rank 0 ((None, None) elsewhere): ONE C call per query
        batch (pinned staging, H2D, kernel chain, peer delivery, merge, D2H, one synchronisation)."""
        import numpy as np
        from . import _native as N
        q = np.ascontiguousarray(q, dtype=np.float32)
        nq = q.shape[0]
        if self.world == 1 or self.transport != "p2p":
            Dt, It = self.search(torch.from_numpy(q).to(self.device), k)
            return (Dt.cpu().numpy(), It.cpu().numpy()) if Dt is not None else (None, None)
        D = I = None
        dp = ip = None
        if self.rank == 0:
            D, I = np.empty((nq, k), np.float32), np.empty((nq, k), np.int64)
            dp, ip = C.c_void_p(D.ctypes.data), C.c_void_p(I.ctypes.data)
        N.check(N.lib().cb_flatip_search_p2p(self._handle(), nq, C.c_void_p(q.ctypes.data), k, dp, ip, self.id_base))
        return D, I

    # ---- pipelined query stream ---------------------------------------------------------------------
    def submit(self, q: torch.Tensor, k: int, out=None):
        """Queue one search and return at once (cb_flatip_submit_search_device): consecutive submissions
        alternate between two lanes, so the selection / exchange / merge tail of one query overlaps the pass
        over the shard of the next.  `out` = (D, I) tensors to fill (rank 0; allocated when omitted).  The
        results are valid after join(); q and out must stay untouched until then."""
        from . import _native as N
        assert self.transport == "p2p" or self.world == 1, "submit() needs the p2p transport"
        nq = q.shape[0]
        D = I = None
        dp = ip = None
        if self.rank == 0:
            D, I = out if out is not None else (torch.empty((nq, k), dtype=torch.float32, device=self.device),
                                                torch.empty((nq, k), dtype=torch.int64, device=self.device))
            dp, ip = C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr())
        stream = torch.cuda.current_stream(q.device).cuda_stream
        N.check(N.lib().cb_flatip_submit_search_device(self._handle(), nq, C.c_void_p(q.data_ptr()), k, dp, ip,
                                                       self.id_base, C.c_void_p(stream)))
        return D, I

    def join(self) -> None:
        """Order the current torch stream after every submitted search (no host synchronisation)."""
        from . import _native as N
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(N.lib().cb_flatip_join(self._handle(), C.c_void_p(stream)))

    def _search_collective(self, q: torch.Tensor, k: int):
        nq = q.shape[0]
        off_I, total = self._buffers(nq, k)
        buf, gather = self._buf, self._gather
        D = buf[:nq * k * 4].view(torch.float32).view(nq, k)
        I = buf[off_I:].view(torch.int64).view(nq, k)
        self._local_search(q, k, D, I, self.id_base)
        if gather.is_cuda:
            dist.all_gather_into_tensor(gather, buf, group=self.group)
        else:  # gloo (CPU tests)
            parts = list(gather.view(self.world, total).unbind(0))
            dist.all_gather(parts, buf, group=self.group)
        return self._merge(gather, self.world, nq, k, off_I, total)
