"""The Pillow half of clip._transform, kept free of torch/CUDA imports so that decode worker
PROCESSES can load it in a fraction of a second (clipb200.clip re-exports both functions).

Reference: `transform(image)` at /root/reference/build-index.py:48 = openai/CLIP's
Compose([Resize(224, BICUBIC), CenterCrop(224), convert("RGB"), ToTensor(), Normalize(...)]);
everything after convert("RGB") is exact per-pixel arithmetic that the uint8 entry points of the
C ABI run on the GPU, so Pillow's pixels in == the reference transform's values out.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np


def resize_center_crop(image, n_px: int = 224):
    """Resize(n_px, bicubic) on the shorter side (torchvision rounding: longer side truncated),
    CenterCrop(n_px), convert("RGB").  Returns a PIL image."""
    from PIL import Image
    w, h = image.size
    if (w, h) != (n_px, n_px):
        if w <= h:
            nw, nh = n_px, int(n_px * h / w)
        else:
            nw, nh = int(n_px * w / h), n_px
        image = image.resize((nw, nh), Image.BICUBIC)
        left, top = int(round((nw - n_px) / 2.0)), int(round((nh - n_px) / 2.0))
        image = image.crop((left, top, left + n_px, top + n_px))
    return image.convert("RGB")


def image_to_u8(image, n_px: int = 224) -> np.ndarray:
    """PIL image -> uint8 [n_px, n_px, 3] ready for encode_image's uint8 path."""
    return np.array(resize_center_crop(image, n_px), dtype=np.uint8)


# ---- decode worker (runs in a child process) -------------------------------------------------

_attached: Dict[str, np.ndarray] = {}


def decode_into_shared(shm_name: str, rows: int, row0: int, files: List[str]) -> List[bool]:
    """Decode `files` with Pillow into rows row0.. of the shared uint8 [rows,224,224,3] block
    `shm_name`.  Returns one flag per file (False: unreadable / not an image; its row is untouched)."""
    from PIL import Image
    arr = _attached.get(shm_name)
    if arr is None:
        # the parent owns the POSIX shared-memory block; mapping its /dev/shm file keeps this process's
        # resource tracker out of it (SharedMemory(name=...) would try to unlink it again at exit)
        arr = np.memmap("/dev/shm/" + shm_name.lstrip("/"), dtype=np.uint8, mode="r+", shape=(rows, 224, 224, 3))
        _attached[shm_name] = arr
    ok = []
    for i, tfn in enumerate(files):
        try:
            with Image.open(tfn) as im:
                arr[row0 + i] = image_to_u8(im)
            ok.append(True)
        except KeyboardInterrupt:
            raise
        except Exception:
            ok.append(False)
    return ok


def _serve() -> None:
    """Worker loop of `python -m clipb200.pil_transform`: one JSON request per line on stdin
    ({"shm", "rows", "row0", "files"}), one JSON reply per line on stdout ({"ok": [...]}).  Plain pipes
    instead of multiprocessing: the worker imports numpy + Pillow only, whatever the parent's __main__ is."""
    import json
    import sys
    out = sys.stdout
    for line in sys.stdin:
        line = line.strip()
        if not line:
            continue
        req = json.loads(line)
        ok = decode_into_shared(req["shm"], int(req["rows"]), int(req["row0"]), req["files"])
        out.write(json.dumps({"ok": ok}) + "\n")
        out.flush()


if __name__ == "__main__":
    _serve()
