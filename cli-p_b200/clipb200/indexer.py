"""Index-time driver: the caller of hot path A re-shaped to batches.

The reference embeds one image per forward pass, with one host->device copy, one sync and
one committed LMDB transaction per file (build-index.py:30-58).  Here files are decoded by
a thread pool (PIL, exactly the reference's Resize/CenterCrop/RGB steps), packed into pinned
uint8 batches of up to 256 and streamed through the pipelined encode entry point
(cb_clip_submit_image_u8: the copy of batch i+1 overlaps the forward pass of batch i);
ToTensor/Normalize, the ViT-B/32 forward and the L2 normalisation of build-index.py:50 run on
the GPU.  Observable behaviour is kept: same folder listing rules (non-recursive, `base_path +
fn`, .jpg/.jpeg/.png), resume by key presence in fn_db, `.` per embedded image / `#` per
failed file, 2048-byte little-endian float32 values, ids assigned in fn_db key order.
"""
from __future__ import annotations

import os
import sys
from concurrent.futures import ThreadPoolExecutor
from typing import Iterable, List, Optional, Tuple

import numpy as np
import torch

EXTS = (".jpg", ".jpeg", ".png")


def list_images(base_path: str) -> List[str]:
    """build-index.py:30-34 -- os.listdir order, plain concatenation (caller passes a trailing '/')."""
    out = []
    for fn in os.listdir(base_path):
        if os.path.splitext(fn)[1].lower() in EXTS:
            out.append(base_path + fn)
    return out


def _decode(tfn: str) -> Optional[np.ndarray]:
    """PIL decode + the reference's Resize/CenterCrop/RGB on the CPU -> uint8 [224,224,3]."""
    from PIL import Image
    from .clip import image_to_u8
    try:
        with Image.open(tfn) as im:
            return image_to_u8(im)
    except KeyboardInterrupt:
        raise
    except Exception:
        return None


def _decode_full(tfn: str):
    """PIL decode only.  RGB images come back at full resolution (uint8 [h,w,3]) for the GPU
    resize; other modes (L, P, RGBA, ...) must be resized in their own mode before the RGB
    conversion (clip._transform order), so they take the CPU path and come back as [224,224,3]."""
    from PIL import Image
    from .clip import image_to_u8
    try:
        with Image.open(tfn) as im:
            if im.mode == "RGB":
                return np.array(im, dtype=np.uint8)
            return image_to_u8(im)
    except KeyboardInterrupt:
        raise
    except Exception:
        return None


def _read_bytes(tfn: str):
    try:
        with open(tfn, "rb") as fh:
            return fh.read()
    except Exception:
        return None


def embed_folders(folders: Iterable[str], env, model, batch: int = 256, workers: Optional[int] = None,
                  out=sys.stdout, resize: str = "cpu", decode: str = "pil") -> Tuple[int, int]:
    """Embed every new image under `folders` into fn_db.  Returns (embedded, failed).

    resize="cpu": Pillow resizes (exactly the reference's transform); resize="gpu": full-resolution
    RGB pixels are uploaded and resized by cb_resize224_u8_device (bit-identical to Pillow).
    decode="nvjpeg" (implies resize="gpu"): JPEG files are decoded on the GPU by nvjpeg through
    torchvision (library work; pixels may differ from libjpeg-turbo by +-1)."""
    if decode == "nvjpeg" or resize == "gpu":
        return _embed_folders_gpu_resize(folders, env, model, batch, workers, out, decode)
    fn_db = env.open_db(b"fn_db")
    skip_db = env.open_db(b"skip_db")
    batch = min(batch, model.max_image_batch)
    workers = workers or min(32, os.cpu_count() or 4)
    n_ok = n_bad = 0
    nbuf = 3
    bufs = [torch.empty((batch, 224, 224, 3), dtype=torch.uint8).pin_memory() for _ in range(nbuf)]
    outs = [torch.empty((batch, 512), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    from . import _native as N
    import ctypes as C
    L = N.lib()

    def commit(names: List[str], vecs: torch.Tensor):
        nonlocal n_ok
        with env.begin(db=fn_db, write=True) as txn:
            for name, v in zip(names, vecs.numpy()):
                txn.put(name.encode(), v.astype("<f4", copy=False).tobytes())
                print(".", end="", flush=True, file=out)
                n_ok += 1

    with ThreadPoolExecutor(max_workers=workers) as pool:
        for base_path in folders:
            print(f"CLIPing {base_path}...", file=out)
            todo = []
            with env.begin(db=skip_db) as st, env.begin(db=fn_db) as ft:
                for tfn in list_images(base_path):
                    key = tfn.encode()
                    if len(key) > 511:              # LMDB rejects the key; the reference skips silently (:40-41)
                        continue
                    if st.get(key) is not None or ft.get(key) is not None:
                        continue
                    todo.append(tfn)
            inflight: List[Tuple[int, List[str], int]] = []   # (slot, names, n)
            slot, names, fill = 0, [], 0

            def flush():
                nonlocal slot, names, fill
                if fill == 0:
                    return
                N.check(L.cb_clip_submit_image_u8(model.handle, fill, C.c_void_p(bufs[slot].data_ptr()),
                                                  C.c_void_p(outs[slot].data_ptr()), 1))
                inflight.append((slot, names, fill))
                # at most two batches are in flight on the device; drain the oldest before its slot is reused
                if len(inflight) >= nbuf - 1:
                    N.check(L.cb_clip_sync(model.handle))
                    for s_, nm_, n_ in inflight:
                        commit(nm_, outs[s_][:n_])
                    inflight.clear()
                slot = (slot + 1) % nbuf
                names, fill = [], 0

            for tfn, px in zip(todo, pool.map(_decode, todo)):
                if px is None:
                    print("#", end="", flush=True, file=out)
                    n_bad += 1
                    continue
                bufs[slot][fill] = torch.from_numpy(px)
                names.append(tfn)
                fill += 1
                if fill == batch:
                    flush()
            flush()
            N.check(L.cb_clip_sync(model.handle))
            for s_, nm_, n_ in inflight:
                commit(nm_, outs[s_][:n_])
            inflight.clear()
            print(flush=True, file=out)
    return n_ok, n_bad


def _nvjpeg_chunks(todo: List[str], chunk: int, pool: ThreadPoolExecutor, dev, threads: Optional[int] = None):
    """Yield (names, items) per chunk of files, in order, decoded up to 2 x `threads` chunks ahead of the
    consumer by `threads` worker threads (CLIPB200_NVJPEG_THREADS, default 2).  torchvision serialises its
    nvjpeg calls (~0.18 ms per 224 px image, measured), so more threads only overlap the file reads.  A JPEG item is
    the decoded CHW uint8 CUDA tensor of one batched nvjpeg call per chunk (torchvision: a list
    decodes far faster than one call per file); if the batched call fails - one corrupt file fails
    the whole list - the chunk is decoded file by file so only the bad ones are lost.  Other
    formats come back as the decoded array of the CPU path.  A failed file is None."""
    import torchvision.io as tvio
    from concurrent.futures import ThreadPoolExecutor as _TPE

    read_op = torch.ops.image.read_file             # the op itself: tvio.read_file adds ~0.1 ms of API logging

    def read_jpeg(tfn: str):
        try:
            return read_op(tfn)                     # uint8 tensor straight from the file, no Python copy
        except Exception:
            return None

    def prepare(names: List[str]):
        items: List = [None] * len(names)
        jidx, jten = [], []
        is_jpeg = [tfn.lower().endswith((".jpg", ".jpeg")) for tfn in names]
        raw = list(pool.map(lambda a: read_jpeg(a[0]) if a[1] else None, zip(names, is_jpeg)))
        for i, tfn in enumerate(names):
            if is_jpeg[i]:
                t = raw[i]
                if t is not None and t.numel() > 0:
                    jidx.append(i)
                    jten.append(t)
            else:
                items[i] = _decode_full(tfn)
        if jten:
            with torch.cuda.device(dev):
                try:
                    dec = tvio.decode_jpeg(jten, device=dev, mode=tvio.ImageReadMode.RGB)
                except Exception:
                    dec = []
                    for t in jten:
                        try:
                            dec.append(tvio.decode_jpeg(t, device=dev, mode=tvio.ImageReadMode.RGB))
                        except Exception:
                            dec.append(None)
            for i, d in zip(jidx, dec):
                items[i] = d
            # the common case - every file of the chunk decoded to 224 x 224 - travels as ONE [m,3,224,224]
            # tensor, so the consumer places it with one strided copy instead of one launch per image
            if len(jidx) == len(names) and all(d is not None and tuple(d.shape) == (3, 224, 224) for d in dec):
                with torch.cuda.device(dev):
                    return names, torch.stack(dec)
        return names, items

    threads = threads or int(os.environ.get("CLIPB200_NVJPEG_THREADS", "2"))
    chunk = max(16, min(chunk, 64))
    chunks = [todo[i:i + chunk] for i in range(0, len(todo), chunk)]
    with _TPE(max_workers=threads) as stage:
        pending = []
        for c in chunks:
            pending.append(stage.submit(prepare, c))
            if len(pending) > 2 * threads:
                yield pending.pop(0).result()
        for f in pending:
            yield f.result()


def _embed_folders_gpu_resize(folders, env, model, batch, workers, out, decode) -> Tuple[int, int]:
    import ctypes as C
    from . import _native as N
    L = N.lib()
    fn_db = env.open_db(b"fn_db")
    skip_db = env.open_db(b"skip_db")
    batch = min(batch, model.max_image_batch)
    workers = workers or min(32, os.cpu_count() or 4)
    dev = model.device
    n_ok = n_bad = 0
    nbuf = 3
    dbuf = [torch.empty((batch, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
    dout = [torch.empty((batch, 512), dtype=torch.float32, device=dev) for _ in range(nbuf)]
    use_nvjpeg = decode == "nvjpeg"

    def commit(names, vecs):
        nonlocal n_ok
        with env.begin(db=fn_db, write=True) as txn:
            for name, v in zip(names, vecs.numpy()):
                txn.put(name.encode(), v.astype("<f4", copy=False).tobytes())
                print(".", end="", flush=True, file=out)
                n_ok += 1

    def to_device_224(item, tfn, dst):
        """item: raw file bytes (nvjpeg) or a decoded array.  Writes [224,224,3] into dst."""
        if torch.is_tensor(item):                       # nvjpeg output: CHW uint8 on the device
            if tuple(item.shape) == (3, 224, 224):
                dst.copy_(item.permute(1, 2, 0))
                return True
            src = item.permute(1, 2, 0).contiguous()
        else:
            src = torch.from_numpy(item).to(dev, non_blocking=True)
        if tuple(src.shape) == (224, 224, 3):
            dst.copy_(src)
        else:
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            N.check(L.cb_resize224_u8_device(C.c_void_p(src.data_ptr()), src.shape[0], src.shape[1],
                                             C.c_void_p(dst.data_ptr()), stream))
            src.record_stream(torch.cuda.current_stream(dev))
        return True

    with torch.cuda.device(dev), ThreadPoolExecutor(max_workers=workers) as pool:
        for base_path in folders:
            print(f"CLIPing {base_path}...", file=out)
            todo = []
            with env.begin(db=skip_db) as st, env.begin(db=fn_db) as ft:
                for tfn in list_images(base_path):
                    key = tfn.encode()
                    if len(key) > 511 or st.get(key) is not None or ft.get(key) is not None:
                        continue
                    todo.append(tfn)
            inflight = []
            slot, names, fill = 0, [], 0

            def drain():
                N.check(L.cb_clip_join(model.handle, model._stream()))
                for s_, nm_, n_ in inflight:
                    commit(nm_, dout[s_][:n_].cpu())
                inflight.clear()

            def flush():
                nonlocal slot, names, fill
                if fill == 0:
                    return
                N.check(L.cb_clip_submit_image_u8_device(model.handle, fill, C.c_void_p(dbuf[slot].data_ptr()),
                                                         C.c_void_p(dout[slot].data_ptr()), 1, model._stream()))
                inflight.append((slot, names, fill))
                if len(inflight) >= nbuf - 1:
                    drain()
                slot = (slot + 1) % nbuf
                names, fill = [], 0

            def nvjpeg_items():
                nonlocal names, fill
                for names_, items_ in _nvjpeg_chunks(todo, batch, pool, dev):
                    if not torch.is_tensor(items_):
                        yield from zip(names_, items_)
                        continue
                    a = 0                                   # whole chunk as one NCHW tensor: bulk placement
                    while a < len(names_):
                        take = min(batch - fill, len(names_) - a)
                        dbuf[slot][fill:fill + take].copy_(items_[a:a + take].permute(0, 2, 3, 1))
                        names.extend(names_[a:a + take])
                        fill += take
                        a += take
                        if fill == batch:
                            flush()

            stream_items = nvjpeg_items() if use_nvjpeg else zip(todo, pool.map(_decode_full, todo))
            for tfn, item in stream_items:
                ok = False
                if item is not None:
                    try:
                        ok = to_device_224(item, tfn, dbuf[slot][fill])
                    except KeyboardInterrupt:
                        raise
                    except Exception:
                        ok = False
                if not ok:
                    print("#", end="", flush=True, file=out)
                    n_bad += 1
                    continue
                names.append(tfn)
                fill += 1
                if fill == batch:
                    flush()
            flush()
            drain()
            print(flush=True, file=out)
    return n_ok, n_bad


def build_index(env, faiss, storage=None, index_path: Optional[str] = "images.index", out=sys.stdout):
    """build-index.py:66-109 -- ids follow fn_db key order; idx_db maps str(id) -> path."""
    fn_db = env.open_db(b"fn_db")
    idx_db = env.open_db(b"idx_db")
    with env.begin(db=fn_db) as txn:
        n = txn.stat()["entries"]
        if n == 0:
            print("Done!", file=out)
            return None
        print(f"Preparing index for {n} entries...", file=out)
        images = np.empty((n, 512), dtype=np.float32)
        print(f"Generating {images.shape} matrix...", file=out)
        names = []
        for i, (tfn, vector) in enumerate(txn.cursor()):
            images[i] = np.frombuffer(vector, dtype=np.float32)
            names.append(tfn)
    with env.begin(db=idx_db, write=True) as itxn:
        for i, tfn in enumerate(names):
            itxn.put(f"{i}".encode(), tfn, dupdata=False, overwrite=True)
    quantizer = faiss.IndexFlatIP(512, storage=storage) if storage is not None else faiss.IndexFlatIP(512)
    index = faiss.IndexIVFFlat(quantizer, 512, 100, faiss.METRIC_INNER_PRODUCT)
    print(f"Training index {images.shape}...", file=out)
    index.train(images)
    print("Adding to index...", file=out)
    index.add(images)
    if index_path:
        print("Saving index...", file=out)
        faiss.write_index(index, index_path)
    print("Done!", file=out)
    return index


class Searcher:
    """Query-time calls of query-index.py:104-119 without the REPL/viewer (out of scope)."""

    def __init__(self, env, index, model=None):
        self.env, self.index, self.model = env, index, model
        self.idx_db = env.open_db(b"idx_db")
        self.fn_db = env.open_db(b"fn_db")

    def features_for_tokens(self, tokens: torch.Tensor) -> np.ndarray:
        f = self.model.encode_text(tokens).detach().cpu().numpy().astype("float32")
        norm = np.linalg.norm(f)                      # query-index.py:13-17
        return f if norm < 1e-9 else f / norm

    def features_for_text(self, text: str) -> np.ndarray:
        from . import clip
        return self.features_for_tokens(clip.tokenize([text]))

    def features_for_id(self, image_id: int) -> np.ndarray:   # `i ID`, query-index.py:86-99
        with self.env.begin(db=self.idx_db) as txn:
            key = txn.get(f"{image_id}".encode())
        with self.env.begin(db=self.fn_db) as txn:
            return np.frombuffer(txn.get(key), dtype=np.float32).reshape((1, 512))

    def path_for_id(self, image_id: int) -> str:
        with self.env.begin(db=self.idx_db) as txn:
            return txn.get(f"{image_id}".encode()).decode()

    def results(self, features: np.ndarray, k: int = 50, offset: int = 0) -> List[Tuple[float, int, str]]:
        """Rows the REPL prints: search k+offset+1, skip ranks j <= offset (query-index.py:111-119).
        Stops at the -1 padding when fewer rows exist than were asked for."""
        D, I = self.index.search(features, k + offset + 1)
        rows = []
        with self.env.begin(db=self.idx_db) as txn:
            for j, i in enumerate(I[0]):
                if j <= offset:
                    continue
                if i < 0:
                    break
                rows.append((float(D[0][j]), int(i), txn.get(f"{i}".encode()).decode()))
        return rows

    @staticmethod
    def format_row(row) -> str:
        return f"{row[0]:.4f} {row[1]} {row[2]}"
