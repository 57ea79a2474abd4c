"""Drop-in for the subset of `import faiss` that CLI-P uses, backed by libclipb200.

Reference call sites (under /root/reference):
  faiss.IndexFlatIP(512)                                    build-index.py:80
  faiss.IndexIVFFlat(quantizer, 512, 100, METRIC_INNER_PRODUCT)   build-index.py:81
  index.train(x) / index.add(x)                             build-index.py:96,99,105,107
  faiss.write_index(index, "images.index")                  build-index.py:109
  faiss.read_index("images.index"); index.nprobe = 32       query-index.py:29-30,51
  D, I = index.search(features, k + offset + 1)             query-index.py:111

Semantics follow faiss's Python wrapper [UPSTREAM class_wrappers.py]: inputs are
C-contiguous float32 numpy arrays of shape (n, d); search returns (D float32
(nq,k), I int64 (nq,k)); shape/type problems raise AssertionError / TypeError;
`k > 0`.  IndexIVFFlat is served by the same exact flat scan (a recall
superset of IVF probing, as the north star demands): train() is accepted and
ignored, nprobe is stored and validated only by the caller.

Every search runs on the GPU through the C ABI.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import _native as N
from . import faiss_io as _io

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1



def _storage_code(storage) -> int:
    if storage is None:
        storage = os.environ.get("CLIPB200_STORAGE", "f32")
    if storage in (N.CB_F32, N.CB_F16) and not isinstance(storage, str):
        return int(storage)
    s = str(storage).lower()
    if s in ("f32", "fp32", "float32"):
        return N.CB_F32
    if s in ("f16", "fp16", "float16", "half"):
        return N.CB_F16
    raise ValueError(f"unknown storage dtype {storage!r}")


def _as_f32_2d(x, d: int, what: str) -> np.ndarray:
    x = np.ascontiguousarray(x)
    if x.dtype != np.float32:
        raise TypeError(f"{what}: expected float32, got {x.dtype}")
    assert x.ndim == 2, f"{what}: expected a 2-D array, got shape {x.shape}"
    assert x.shape[1] == d, f"{what}: dimension {x.shape[1]} != index dimension {d}"
    return x


class _Shard:
    """Borrowed view of one device-resident shard (a cb_index handle owned by the cb_sharded)."""

    def __init__(self, handle, device: int):
        self.handle = handle
        self.device = device

    @property
    def ntotal(self) -> int:
        return int(N.lib().cb_flatip_ntotal(self.handle))


class IndexFlatIP:
    """Exact inner-product index; rows live in HBM, ids are add() order.

    devices: CUDA ordinals holding the shards.  One device (default: the current
    torch device if torch is imported and CUDA is initialised, else 0) gives the
    plain single-GPU index.  Several devices shard every add() call contiguously;
    a search is ONE C call (cb_sharded_search): the query fans out, every device
    runs its kernel chain and stores its top-k into a mailbox in devices[0]'s
    memory over NVLink, and devices[0] merges (include/clipb200.h, cb_sharded_*).
    """

    def __init__(self, d: int = 512, storage=None, devices: Optional[Sequence[int]] = None):
        self.d = int(d)
        self.metric_type = METRIC_INNER_PRODUCT
        self.is_trained = True
        self.verbose = False
        self._storage = _storage_code(storage)
        if devices is None:
            env = os.environ.get("CLIPB200_DEVICES")
            devices = [int(t) for t in env.split(",")] if env else [_default_device()]
        self._devices = [int(t) for t in devices]
        assert len(self._devices) >= 1
        self._handle = C.c_void_p()
        arr = (C.c_int * len(self._devices))(*self._devices)
        N.check(N.lib().cb_sharded_create(self.d, self._storage, len(self._devices), arr, C.byref(self._handle)))
        self._shards: List[_Shard] = [
            _Shard(C.c_void_p(N.lib().cb_sharded_shard(self._handle, r)), dev) for r, dev in enumerate(self._devices)]

    def close(self):
        if getattr(self, "_handle", None):
            N.lib().cb_sharded_free(self._handle)
            self._handle = None
            self._shards = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- faiss attributes ------------------------------------------------------
    @property
    def ntotal(self) -> int:
        return int(N.lib().cb_sharded_ntotal(self._handle))

    @property
    def storage(self) -> str:
        return "f16" if self._storage == N.CB_F16 else "f32"

    def train(self, x) -> None:  # flat indexes need no training
        _as_f32_2d(x, self.d, "train")

    def reset(self) -> None:
        N.check(N.lib().cb_sharded_reset(self._handle))

    def reserve(self, n: int) -> None:
        N.check(N.lib().cb_sharded_reserve(self._handle, int(n)))

    # -- add -------------------------------------------------------------------
    def add(self, x) -> None:
        x = _as_f32_2d(x, self.d, "add")
        if x.shape[0]:
            N.check(N.lib().cb_sharded_add(self._handle, x.shape[0], x.ctypes.data_as(C.c_void_p)))

    def add_device(self, x_dev, shard: int = 0) -> None:
        """Append rows already on the GPU (torch CUDA tensor, fp16 or fp32, (n, d)) to one shard;
        they get the next n ids."""
        import torch
        assert x_dev.is_cuda and x_dev.dim() == 2 and x_dev.shape[1] == self.d
        assert x_dev.dtype in (torch.float16, torch.float32)
        x_dev = x_dev.contiguous()
        assert x_dev.device.index == self._devices[shard]
        src = N.CB_F16 if x_dev.dtype == torch.float16 else N.CB_F32
        stream = torch.cuda.current_stream(x_dev.device).cuda_stream
        N.check(N.lib().cb_sharded_add_device(self._handle, shard, x_dev.shape[0], C.c_void_p(x_dev.data_ptr()), src,
                                              C.c_void_p(stream)))

    # -- search ----------------------------------------------------------------
    def search(self, x, k: int):
        x = _as_f32_2d(x, self.d, "search")
        k = int(k)
        assert k > 0, "search: k must be > 0"
        nq = x.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        if nq:
            N.check(N.lib().cb_sharded_search(self._handle, nq, x.ctypes.data_as(C.c_void_p), k,
                                              D.ctypes.data_as(C.c_void_p), I.ctypes.data_as(C.c_void_p)))
        return D, I

    def search_device(self, q, k: int):
        """Device-resident search: q is a torch tensor (nq, d) float32 (host or any
        device); returns torch CUDA tensors (D, I) on devices[0].  No host sync."""
        import torch
        k = int(k)
        assert k > 0
        nq = q.shape[0]
        dev0 = torch.device("cuda", self._devices[0])
        with torch.cuda.device(dev0):
            qd = q.to(device=dev0, dtype=torch.float32, non_blocking=True).contiguous()
            D = torch.empty((nq, k), dtype=torch.float32, device=dev0)
            I = torch.empty((nq, k), dtype=torch.int64, device=dev0)
            stream = torch.cuda.current_stream(dev0).cuda_stream
            N.check(N.lib().cb_sharded_search_device(self._handle, nq, C.c_void_p(qd.data_ptr()), k,
                                                     C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()),
                                                     C.c_void_p(stream)))
            qd.record_stream(torch.cuda.current_stream(dev0))
        return D, I

    # -- reconstruct -------------------------------------------------------------
    def reconstruct_n(self, n0: int = 0, ni: int = -1) -> np.ndarray:
        if ni < 0:
            ni = self.ntotal - n0
        assert 0 <= n0 and n0 + ni <= self.ntotal
        out = np.empty((ni, self.d), dtype=np.float32)
        if ni:
            N.check(N.lib().cb_sharded_get_rows(self._handle, n0, ni, out.ctypes.data_as(C.c_void_p)))
        return out

    def reconstruct(self, i: int) -> np.ndarray:
        return self.reconstruct_n(int(i), 1)[0]


class IndexIVFFlat:
    """Accepted for surface compatibility (build-index.py:81); searches exactly.

    The north star replaces IVF probing by an exact flat scan, which returns a
    superset-quality result (identical to IVF with nprobe == nlist).  `train` is a
    no-op apart from argument validation; `nprobe` is kept as a plain attribute.
    """

    def __init__(self, quantizer: IndexFlatIP, d: int, nlist: int, metric: int = METRIC_INNER_PRODUCT):
        assert isinstance(quantizer, IndexFlatIP), "quantizer must be an IndexFlatIP"
        assert metric == METRIC_INNER_PRODUCT, "only METRIC_INNER_PRODUCT is supported"
        assert int(d) == quantizer.d
        self.quantizer = quantizer       # stays empty, as faiss's coarse quantizer would before train()
        self.d = int(d)
        self.nlist = int(nlist)
        self.nprobe = 1
        self.metric_type = metric
        self.is_trained = False
        # the rows live in a flat store of the IVF's own, on the quantizer's devices / storage dtype
        self._flat = IndexFlatIP(self.d, storage=quantizer._storage, devices=quantizer._devices)

    @property
    def ntotal(self) -> int:
        return self._flat.ntotal

    def train(self, x) -> None:
        _as_f32_2d(x, self.d, "train")
        self.is_trained = True

    def add(self, x) -> None:
        assert self.is_trained, "IndexIVFFlat.add before train (faiss raises here too)"
        self._flat.add(x)

    def search(self, x, k: int):
        return self._flat.search(x, k)

    def reset(self) -> None:
        self._flat.reset()


# ---- merge (the step after the cross-GPU gather) ---------------------------------

def merge_topk_device(Dall, Iall, k: int, shard_stride_D: int = 0, shard_stride_I: int = 0,
                      R: Optional[int] = None, nq: Optional[int] = None):
    """Merge per-shard (R, nq, k) CUDA tensors into (nq, k) on the same device."""
    import torch
    if R is None:
        R, nq = Dall.shape[0], Dall.shape[1]
    dev = Dall.device
    D = torch.empty((nq, k), dtype=torch.float32, device=dev)
    I = torch.empty((nq, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().cb_topk_merge_device(R, nq, k, C.c_void_p(Dall.data_ptr()), C.c_void_p(Iall.data_ptr()),
                                             shard_stride_D, shard_stride_I,
                                             C.c_void_p(D.data_ptr()), C.c_void_p(I.data_ptr()),
                                             C.c_void_p(stream)))
    return D, I


# ---- index file I/O ---------------------------------------------------------------
# `images.index` is written and read in faiss's own on-disk format (faiss_io.py; SURVEY 8f
# row 4), so files travel between faiss and clipb200 in both directions:
#   IndexFlatIP   -> "IxFI"
#   IndexIVFFlat  -> "IwFl" with nlist = 1: one inverted list holding every row in id order and a
#                    one-centroid quantizer.  clipb200 serves IVF by an exact scan, so it has no
#                    k-means partition to record; a one-list IVF searches exactly in faiss as well.
# read_index flattens any IwFl file (full or sparse list table, any nlist) back into add order.

def write_index(index, path: str) -> None:
    """faiss.write_index(index, "images.index")  (build-index.py:109)."""
    ivf = isinstance(index, IndexIVFFlat)
    flat = index._flat if ivf else index
    if not isinstance(flat, IndexFlatIP):
        raise TypeError(f"write_index: unsupported index type {type(index).__name__}")
    if ivf:
        _io.write_ivf_single_list(path, flat.d, flat.ntotal, flat.reconstruct_n, nprobe=getattr(index, "nprobe", 1))
    else:
        _io.write_flat(path, flat.d, flat.ntotal, flat.reconstruct_n)


def read_index(path: str, storage=None, devices: Optional[Sequence[int]] = None):
    """faiss.read_index("images.index")  (query-index.py:29).  The rows go to the GPU(s) in id
    order; `storage` / CLIPB200_STORAGE picks fp32 (default) or fp16 rows in HBM."""
    try:
        parsed = _io.parse(path)
    except _io.FaissFormatError as e:
        raise RuntimeError(str(e)) from None
    if parsed.metric != METRIC_INNER_PRODUCT:
        raise RuntimeError(f"{path}: metric type {parsed.metric} is not supported (CLI-P uses inner product, "
                           f"build-index.py:80-81)")
    flat = IndexFlatIP(parsed.d, storage=storage, devices=devices)
    flat.reserve(parsed.ntotal)
    for rows in parsed.iter_rows():
        flat.add(rows)
    kind, nlist, nprobe = (1 if parsed.kind == "ivf" else 0), parsed.nlist, parsed.nprobe
    if kind == 1:
        ivf = IndexIVFFlat.__new__(IndexIVFFlat)
        ivf.quantizer = IndexFlatIP(parsed.d, storage=storage, devices=devices)
        ivf.d, ivf.nlist, ivf.nprobe, ivf.metric_type = flat.d, nlist, nprobe, METRIC_INNER_PRODUCT
        ivf.is_trained = True
        ivf._flat = flat
        return ivf
    return flat


def _default_device() -> int:
    import sys
    t = sys.modules.get("torch")
    if t is not None and t.cuda.is_available() and t.cuda.is_initialized():
        return int(t.cuda.current_device())
    return int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("CLIPB200_USE_LOCAL_RANK") else 0
