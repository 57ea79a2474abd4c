"""Decode worker processes of the index-time Pillow path (clipb200/indexer.py::_PilProcessPool): same
pixels as the in-process decode, failures flagged per file, shared-memory blocks released."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))


def test_process_pool_matches_in_process_decode(tmp_path):
    from PIL import Image
    from clipb200 import indexer
    rng = np.random.default_rng(3)
    files = []
    for i in range(70):
        h, w = [(224, 224), (300, 200), (224, 500), (97, 131)][i % 4]
        arr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        p = str(tmp_path / f"f{i:03d}{'.png' if i % 9 == 0 else '.jpg'}")
        Image.fromarray(arr).save(p, quality=85)
        files.append(p)
    gray = str(tmp_path / "gray.jpg")
    Image.fromarray(rng.integers(0, 256, (250, 250), dtype=np.uint8)).save(gray)
    broken = str(tmp_path / "broken.jpg")
    open(broken, "wb").write(b"nope")
    files = files[:20] + [broken] + files[20:] + [gray, str(tmp_path / "missing.jpg")]
    pool = indexer._PilProcessPool(nproc=3, chunk=32, depth=2)
    names_seen = []
    try:
        shm_names = [s.name for s in pool.slots]
        for names, px, ok in pool.chunks(files):
            assert px.shape == (32, 224, 224, 3) and len(ok) == len(names)
            for i, tfn in enumerate(names):
                want = indexer._decode(tfn)
                assert ok[i] == (want is not None), tfn
                if want is not None:
                    assert np.array_equal(px[i], want), tfn
            names_seen += names
    finally:
        pool.close()
    assert names_seen == files
    from multiprocessing import shared_memory
    for n in shm_names:                              # blocks are unlinked on close
        try:
            shared_memory.SharedMemory(name=n)
            raise AssertionError(f"shared block {n} still exists")
        except FileNotFoundError:
            pass
