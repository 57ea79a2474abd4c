"""Batched JPEG decode into HBM through the C ABI (csrc/jpeg.cu): the index-time replacement for
`Image.open(tfn)` + `transform(image)` at /root/reference/build-index.py:47-48.

One call decodes a list of files with N host threads (nvjpeg's decoupled host/device phases, own
state and CUDA stream per thread) into a uint8 [n,224,224,3] CUDA tensor; files that are not
224 x 224 go through the Pillow-exact resize kernel.  Per-file status: 0 ok, 1 unreadable, 2 not a
decodable JPEG, 3 CUDA error, 4 unsupported here (decode it on the CPU instead).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native as N

OK, UNREADABLE, BAD_JPEG, CUDA_ERROR, UNSUPPORTED = 0, 1, 2, 3, 4


class Decoder:
    def __init__(self, device: int = 0, threads: int = 0):
        self.handle = C.c_void_p()
        self.device = int(device)
        N.check(N.lib().cb_jpeg_create(self.device, int(threads), C.byref(self.handle)))
        self.threads = int(N.lib().cb_jpeg_threads(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None):
            N.lib().cb_jpeg_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check_out(self, n: int, out: Optional[torch.Tensor]) -> torch.Tensor:
        if out is None:
            out = torch.empty((n, 224, 224, 3), dtype=torch.uint8, device=torch.device("cuda", self.device))
        assert out.is_cuda and out.device.index == self.device and out.dtype == torch.uint8 and out.is_contiguous()
        assert out.dim() == 4 and out.shape[0] >= n and tuple(out.shape[1:]) == (224, 224, 3)
        return out

    def decode_files(self, paths: Sequence[str], out: Optional[torch.Tensor] = None):
        """-> (pixels uint8 [n,224,224,3] on the device, status int32 [n]).  Synchronous: the pixels of
        every status-0 file are complete on return (they were written on the decoder's own streams;
        work queued on other streams before this call is not waited for)."""
        n = len(paths)
        out = self._check_out(n, out)
        status = np.full(n, -1, dtype=np.int32)
        if n:
            arr = (C.c_char_p * n)(*[p.encode() if isinstance(p, str) else p for p in paths])
            N.check(N.lib().cb_jpeg_decode_files(self.handle, n, arr, C.c_void_p(out.data_ptr()),
                                                 C.c_void_p(status.ctypes.data)))
        return out, status

    def decode_bytes(self, blobs: Sequence[bytes], out: Optional[torch.Tensor] = None):
        n = len(blobs)
        out = self._check_out(n, out)
        status = np.full(n, -1, dtype=np.int32)
        if n:
            keep = [np.frombuffer(b, dtype=np.uint8) for b in blobs]
            ptrs = (C.c_void_p * n)(*[k.ctypes.data if k.size else None for k in keep])
            sizes = (C.c_int64 * n)(*[k.size for k in keep])
            N.check(N.lib().cb_jpeg_decode_memory(self.handle, n, ptrs, sizes, C.c_void_p(out.data_ptr()),
                                                  C.c_void_p(status.ctypes.data)))
        return out, status
