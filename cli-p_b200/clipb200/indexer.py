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
    from .pil_transform import image_to_u8
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
    from .pil_transform import image_to_u8
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


class _PilProcessPool:
    """Pillow decode in worker PROCESSES (the thread pool tops out near 1.5 k images/s: about half of a
    small file's decode time is Python code under the GIL).  Workers write uint8 pixels straight into
    shared-memory blocks of [chunk,224,224,3]; only file names and ok-flags cross the pipes.  Each worker
    is a plain `python -m clipb200.pil_transform` child (numpy + Pillow, no torch, no CUDA, nothing of
    this process's __main__) driven by its own dispatcher thread over stdin/stdout."""

    def __init__(self, nproc: int, chunk: int, depth: int = 3):
        import subprocess
        import threading
        from multiprocessing import shared_memory
        self.nproc, self.chunk, self.depth = nproc, chunk, depth
        self.slots = [shared_memory.SharedMemory(create=True, size=chunk * 224 * 224 * 3) for _ in range(depth + 1)]
        self.views = [np.ndarray((chunk, 224, 224, 3), dtype=np.uint8, buffer=s.buf) for s in self.slots]
        self._local = threading.local()
        self._children: List = []
        self._lock = threading.Lock()
        pkg_parent = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ)
        env["PYTHONPATH"] = pkg_parent + os.pathsep + env.get("PYTHONPATH", "")

        def child():
            c = getattr(self._local, "child", None)
            if c is None:
                c = subprocess.Popen([sys.executable, "-m", "clipb200.pil_transform"], stdin=subprocess.PIPE,
                                     stdout=subprocess.PIPE, env=env, text=True, bufsize=1)
                self._local.child = c
                with self._lock:
                    self._children.append(c)
            return c

        self._child = child
        self.pool = ThreadPoolExecutor(max_workers=nproc)

    def _task(self, shm_name: str, row0: int, files: List[str]) -> List[bool]:
        import json
        c = self._child()
        c.stdin.write(json.dumps({"shm": shm_name, "rows": self.chunk, "row0": row0, "files": files}) + "\n")
        c.stdin.flush()
        line = c.stdout.readline()
        if not line:
            raise RuntimeError("Pillow decode worker exited unexpectedly")
        return json.loads(line)["ok"]

    def chunks(self, todo: List[str]):
        """Yield (names, pixels, ok) per chunk, in order; `pixels` is a view into a shared block that stays
        valid until the generator is resumed."""
        parts = [todo[i:i + self.chunk] for i in range(0, len(todo), self.chunk)]
        pending = []

        def submit(k: int, names: List[str]):
            b = k % len(self.slots)
            sub = max(1, -(-len(names) // (2 * self.nproc)))
            futs = [self.pool.submit(self._task, self.slots[b].name, r0, names[r0:r0 + sub])
                    for r0 in range(0, len(names), sub)]
            return names, b, futs

        k = 0
        while k < len(parts) or pending:
            while k < len(parts) and len(pending) < self.depth:
                pending.append(submit(k, parts[k]))
                k += 1
            names, b, futs = pending.pop(0)
            ok = np.array([flag for f in futs for flag in f.result()], dtype=bool)
            yield names, self.views[b], ok

    def close(self):
        self.pool.shutdown(wait=True, cancel_futures=True)
        for c in self._children:
            try:
                c.stdin.close()
                c.wait(timeout=10)
            except Exception:
                c.kill()
        self._children = []
        self.views = []
        for s in self.slots:
            try:
                s.close()
                s.unlink()
            except Exception:
                pass
        self.slots = []


META_DB = b"clipb200_meta"      # the reference opens vectors.lmdb with max_dbs=4 and uses three: this is the fourth


def check_weights_stamp(env, model) -> None:
    """vectors.lmdb resumes by key presence (build-index.py:40-42), so rows written with other weights are
    never recomputed.  The store therefore carries the identity of the checkpoint that filled it
    (clipb200_meta / b"weights"): a fresh store is stamped, a store stamped by different weights is refused
    (CLIPB200_ALLOW_MIXED_WEIGHTS=1 overrides), a store without a stamp that already has rows (made by the
    reference, or by an older clipb200) is left alone."""
    m = model[0] if isinstance(model, (list, tuple)) else model
    wid = getattr(m, "weights_id", None)
    if not wid:
        return
    have = None
    try:
        meta = env.open_db(META_DB, create=False)
        with env.begin(db=meta) as txn:
            have = txn.get(b"weights")
    except Exception:
        meta = None                             # no stamp database yet
    if have is not None:
        if bytes(have).decode() != wid and os.environ.get("CLIPB200_ALLOW_MIXED_WEIGHTS") != "1":
            raise RuntimeError(
                f"vectors.lmdb was filled with CLIP weights {bytes(have).decode()}, this model is {wid}: rows of both would "
                "mix in one index and resume-by-key would never recompute the old ones. Use a fresh vectors.lmdb "
                "(or CLIPB200_ALLOW_MIXED_WEIGHTS=1 if you know they are the same model).")
        return
    fn_db = env.open_db(b"fn_db")
    with env.begin(db=fn_db) as txn:
        empty = txn.stat(fn_db)["entries"] == 0
    if empty:
        try:
            meta = env.open_db(META_DB)
        except Exception:
            return                              # max_dbs exhausted or a read-only environment: nothing to record
        with env.begin(db=meta, write=True) as txn:
            txn.put(b"weights", wid.encode())


def embed_folders(folders: Iterable[str], env, model, batch: int = 256, workers: Optional[int] = None,
                  out=sys.stdout, resize: str = "cpu", decode: str = "pil") -> Tuple[int, int]:
    """Embed every new image under `folders` into fn_db.  Returns (embedded, failed).
    `model` may be a list of CLIPB200 handles, one per GPU: the files are then split over the devices
    (GPU decode/resize path), each device runs independently and this thread commits to LMDB.

    resize="cpu": Pillow resizes (exactly the reference's transform); resize="gpu": full-resolution
    RGB pixels are uploaded and resized by cb_resize224_u8_device (bit-identical to Pillow).
    decode="nvjpeg" (implies resize="gpu"): JPEG files are read and decoded a batch per call by
    clipb200.jpeg.Decoder (host threads + nvjpeg behind the C ABI; library work for the decode itself,
    pixels may differ from libjpeg-turbo by +-1); other formats still go through Pillow."""
    check_weights_stamp(env, model)
    if decode == "nvjpeg" or resize == "gpu" or isinstance(model, (list, tuple)):
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

    # decode workers: processes for folders worth their start-up (CLIPB200_PIL_PROCESSES=N pins the count,
    # 0 keeps the thread pool), threads otherwise
    env_np = os.environ.get("CLIPB200_PIL_PROCESSES")
    nproc = int(env_np) if env_np is not None else (min(32, os.cpu_count() or 1) if (os.cpu_count() or 1) >= 4 else 0)
    procs: Optional[_PilProcessPool] = None
    with ThreadPoolExecutor(max_workers=workers) as pool:
      try:
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
            use_procs = nproc > 0 and (env_np is not None or len(todo) >= 2048)
            if use_procs and procs is None:
                try:
                    # the shared batch blocks live in /dev/shm: a worker writing past a too-small tmpfs dies of SIGBUS
                    fs = os.statvfs("/dev/shm")
                    if fs.f_bavail * fs.f_frsize < 6 * batch * 224 * 224 * 3:
                        raise OSError("/dev/shm is too small for the shared batch blocks")
                    procs = _PilProcessPool(nproc, batch)
                except Exception as e:
                    print(f"clipb200: decode worker processes unavailable ({e}); using the thread pool", file=sys.stderr)
                    nproc, use_procs = 0, False

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

            if use_procs:
                for names_, px, ok in procs.chunks(todo):
                    good = np.nonzero(ok)[0]
                    a = 0                               # bulk placement of the chunk's decoded rows
                    while a < len(good):
                        take = min(batch - fill, len(good) - a)
                        sel = good[a:a + take]
                        src = px[:take] if take == len(names_) else px[sel]
                        bufs[slot][fill:fill + take] = torch.from_numpy(src)
                        names.extend(names_[i] for i in sel)
                        fill += take
                        a += take
                        if fill == batch:
                            flush()
                    for _ in range(len(names_) - len(good)):
                        print("#", end="", flush=True, file=out)
                        n_bad += 1
            else:
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
      finally:
        if procs is not None:
            procs.close()
    return n_ok, n_bad


def _nvjpeg_chunks(todo: List[str], chunk: int, decoder, dev, depth: int = 1):
    """Yield (names, pixels, status) per chunk of files, in order, decoded `depth` chunks ahead of the
    consumer.  JPEG files of a chunk are decoded by ONE call into the C ABI (clipb200.jpeg.Decoder:
    host threads + nvjpeg, no Python in the loop) into a rotating [chunk,224,224,3] device buffer;
    status[i] is 0 for a decoded file, -1 for a file that is not a JPEG (the caller decodes it on the
    CPU), anything else for a failure (see clipb200/jpeg.py)."""
    stages = [torch.empty((chunk, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(depth + 2)]
    # the consumer records `released[b]` once its copies out of stages[b] are queued; the decoder (which
    # writes on its own streams) waits for that before it overwrites the buffer
    released: List[Optional[torch.cuda.Event]] = [None] * len(stages)

    def prepare(k: int, names: List[str]):
        jidx = [i for i, tfn in enumerate(names) if tfn.lower().endswith((".jpg", ".jpeg"))]
        status = np.full(len(names), -1, dtype=np.int32)
        b = k % len(stages)
        px = stages[b]
        if released[b] is not None:
            released[b].synchronize()
        if jidx:
            if len(jidx) == len(names):
                _, st = decoder.decode_files(names, out=px)
                status[:] = st
            else:                                   # mixed chunk: decode the JPEGs to the front, note where they go
                tmp, st = decoder.decode_files([names[i] for i in jidx])
                px[torch.as_tensor(jidx, device=dev)] = tmp
                status[jidx] = st
        return names, px, status, b

    def release(b: int):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        released[b] = ev

    chunks = [todo[i:i + chunk] for i in range(0, len(todo), chunk)]
    with ThreadPoolExecutor(max_workers=1) as stage:
        pending = []
        for k, c in enumerate(chunks):
            pending.append(stage.submit(prepare, k, c))
            if len(pending) > depth:
                names, px, status, b = pending.pop(0).result()
                yield names, px, status
                release(b)
        for f in pending:
            names, px, status, b = f.result()
            yield names, px, status
            release(b)


_decoders: dict = {}


def _jpeg_decoder(device: int):
    """One clipb200.jpeg.Decoder per device for the life of the process: creating one (an nvjpeg handle, two
    back-end states, buffers and a stream per worker thread) costs ~0.3 s."""
    from . import jpeg
    key = (device, os.environ.get("CLIPB200_NVJPEG_THREADS", "0"))
    if key not in _decoders:
        _decoders[key] = jpeg.Decoder(device, int(key[1]))
    return _decoders[key]


def _device_pipeline(todo: List[str], model, batch: int, workers: int, decode: str, on_done, on_bad) -> None:
    """Embed `todo` on model.device: decode (nvjpeg through the C ABI, or Pillow in a thread pool) and
    resize on the GPU, batches of `batch` through the two-lane submit API.  Calls on_done(names, vecs)
    (vecs: float32 CPU tensor [n,512], L2-normalised) per finished batch and on_bad(name) per failed file,
    from the calling thread.  One call per device; several may run in parallel threads."""
    import ctypes as C
    from . import _native as N
    L = N.lib()
    dev = model.device
    nbuf = 3
    use_nvjpeg = decode == "nvjpeg"
    with torch.cuda.device(dev), ThreadPoolExecutor(max_workers=workers) as pool:
        dbuf = [torch.empty((batch, 224, 224, 3), dtype=torch.uint8, device=dev) for _ in range(nbuf)]
        dout = [torch.empty((batch, 512), dtype=torch.float32, device=dev) for _ in range(nbuf)]
        decoder = _jpeg_decoder(dev.index) if use_nvjpeg else None

        def to_device_224(item, dst):
            """item: a decoded uint8 array [h,w,3].  Writes [224,224,3] into dst."""
            src = torch.from_numpy(item).to(dev, non_blocking=True)
            if tuple(src.shape) == (224, 224, 3):
                dst.copy_(src)
            else:
                stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                N.check(L.cb_resize224_u8_device(C.c_void_p(src.data_ptr()), src.shape[0], src.shape[1],
                                                 C.c_void_p(dst.data_ptr()), stream))
                src.record_stream(torch.cuda.current_stream(dev))

        inflight = []
        slot, names, fill = 0, [], 0

        def drain():
            N.check(L.cb_clip_join(model.handle, model._stream()))
            for s_, nm_, n_ in inflight:
                on_done(nm_, dout[s_][:n_].cpu())
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
            for names_, px, status in _nvjpeg_chunks(todo, batch, decoder, dev):
                good = np.nonzero(status == 0)[0]
                a = 0                                   # decoded rows: bulk placement, one gather per batch slot
                while a < len(good):
                    take = min(batch - fill, len(good) - a)
                    sel = good[a:a + take]
                    if take == len(names_):
                        dbuf[slot][fill:fill + take].copy_(px[:take])
                    else:
                        dbuf[slot][fill:fill + take] = px[torch.as_tensor(sel, device=dev)]
                    names.extend(names_[i] for i in sel)
                    fill += take
                    a += take
                    if fill == batch:
                        flush()
                for i in np.nonzero(status != 0)[0]:
                    # not a JPEG, or nvjpeg would not take it (CMYK, arithmetic coding, a damaged stream Pillow
                    # may still read): Pillow decides; only an unreadable file is given up on at once
                    yield names_[i], (None if status[i] == 1 else _decode_full(names_[i]))

        stream_items = nvjpeg_items() if use_nvjpeg else zip(todo, pool.map(_decode_full, todo))
        for tfn, item in stream_items:
            ok = False
            if item is not None:
                try:
                    to_device_224(item, dbuf[slot][fill])
                    ok = True
                except KeyboardInterrupt:
                    raise
                except Exception:
                    ok = False
            if not ok:
                on_bad(tfn)
                continue
            names.append(tfn)
            fill += 1
            if fill == batch:
                flush()
        flush()
        drain()


def _embed_folders_gpu_resize(folders, env, model, batch, workers, out, decode) -> Tuple[int, int]:
    """GPU decode/resize driver.  `model` is one CLIPB200 or a list of them, one per device (SURVEY.md 8e:
    the files of a folder are split contiguously over the devices, each device runs its own pipeline in a
    thread with no collective, and this thread - the only LMDB writer - commits what they finish)."""
    import queue
    import threading
    models = list(model) if isinstance(model, (list, tuple)) else [model]
    fn_db = env.open_db(b"fn_db")
    skip_db = env.open_db(b"skip_db")
    batch = min([batch] + [m.max_image_batch for m in models])
    workers = workers or max(2, min(32, os.cpu_count() or 4) // len(models))
    n_ok = n_bad = 0

    def commit(names, vecs):
        nonlocal n_ok
        with env.begin(db=fn_db, write=True) as txn:
            for name, v in zip(names, vecs.numpy()):
                txn.put(name.encode(), v.astype("<f4", copy=False).tobytes())
                print(".", end="", flush=True, file=out)
                n_ok += 1

    def bad(name):
        nonlocal n_bad
        print("#", end="", flush=True, file=out)
        n_bad += 1

    for base_path in folders:
        print(f"CLIPing {base_path}...", file=out)
        todo = []
        with env.begin(db=skip_db) as st, env.begin(db=fn_db) as ft:
            for tfn in list_images(base_path):
                key = tfn.encode()
                if len(key) > 511 or st.get(key) is not None or ft.get(key) is not None:
                    continue
                todo.append(tfn)
        # one pipeline thread per device (also for a single GPU: decode/placement/submit then overlap with
        # this thread's LMDB commits, ~45 k puts/s of pure Python)
        R = len(models)
        per = -(-len(todo) // R) if todo else 0
        events: "queue.Queue" = queue.Queue(maxsize=4 * R)
        errors: List[BaseException] = []

        def run(r):
            try:
                _device_pipeline(todo[r * per:(r + 1) * per], models[r], batch, workers, decode,
                                 lambda nm, v: events.put(("ok", nm, v)), lambda nm: events.put(("bad", nm, None)))
            except BaseException as e:      # surfaced in the committing thread
                errors.append(e)
            finally:
                events.put(("end", None, None))

        threads = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(R)]
        for t in threads:
            t.start()
        live = R
        while live:
            kind, nm, v = events.get()
            if kind == "ok":
                commit(nm, v)
            elif kind == "bad":
                bad(nm)
            else:
                live -= 1
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
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
