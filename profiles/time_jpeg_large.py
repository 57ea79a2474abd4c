"""Decode throughput of clipb200.jpeg.Decoder on multi-megapixel photos (what real folders hold), per nvjpeg
backend: 1 = Huffman on the host cores (default), 2 = GPU-assisted Huffman.  64 links to 4 unique files."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import numpy as np
import torch
from PIL import Image
from clipb200 import jpeg

tmp = tempfile.mkdtemp()
rng = np.random.default_rng(0)
paths = []
for w, h in ((224, 224), (320, 240), (640, 480), (1024, 768), (1920, 1080), (4000, 3000)):
    uniq = []
    for i in range(4):
        base = rng.integers(0, 256, (12, 16, 3), dtype=np.uint8)
        im = np.asarray(Image.fromarray(base).resize((w, h), Image.BICUBIC), dtype=np.float32)
        arr = np.clip(im + rng.normal(0, 6, im.shape), 0, 255).astype(np.uint8)
        p = os.path.join(tmp, f"u{w}_{i}.jpg")
        Image.fromarray(arr).save(p, quality=90)
        uniq.append(p)
    files = []
    for i in range(256):
        q = os.path.join(tmp, f"l{w}_{i:03d}.jpg")
        os.link(uniq[i % 4], q)
        files.append(q)
    paths.append(((w, h), files, os.path.getsize(uniq[0])))
dec = jpeg.Decoder(0)
for (w, h), files, size in paths:
    dec.decode_files(files[:16])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    px, st = dec.decode_files(files)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"backend {os.environ.get('CLIPB200_NVJPEG_BACKEND', 'auto')} threads {dec.threads}: {w}x{h} ({size / 1e6:.1f} MB files) "
          f"{len(files) / dt:.0f} images/s, {len(files) * w * h / dt / 1e6:.0f} Mpixel/s, ok={int((st == 0).sum())}/{len(files)}", flush=True)
