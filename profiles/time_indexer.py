"""Wall-clock throughput of indexer.embed_folders over a folder of synthetic 224 px JPEGs (1000 unique
files hard-linked to the requested count), per decode path.  usage: time_indexer.py N mode [mode...]"""
import io
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from bench_configs import make_jpeg_folder
from clipb200 import clip, indexer, lmdb, weights

n = int(sys.argv[1])
modes = sys.argv[2:] or ["nvjpeg"]
tmp = tempfile.mkdtemp()
uniq = os.path.join(tmp, "uniq") + "/"
make_jpeg_folder(uniq, min(n, 1000))
folder = os.path.join(tmp, "photos") + "/"
os.makedirs(folder)
files = sorted(os.listdir(uniq))
for i in range(n):
    os.link(uniq + files[i % len(files)], folder + f"img_{i:07d}.jpg")
model = clip.CLIPB200(weights.synthetic_state_dict(0), device=0, max_image_batch=256, max_text_batch=1)
os.chdir(tmp)
for mode in modes:
    kw = {"decode": "nvjpeg"} if mode == "nvjpeg" else {}
    wenv = lmdb.open(f"w_{mode}.lmdb", map_size=1 << 30, max_dbs=4)
    indexer.embed_folders([uniq], wenv, model, out=io.StringIO(), **kw)       # warm-up
    env = lmdb.open(f"v_{mode}.lmdb", map_size=8 << 30, max_dbs=4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ok, bad = indexer.embed_folders([folder], env, model, out=io.StringIO(), **kw)
    dt = time.perf_counter() - t0
    print(f"{mode}: {ok} files ({bad} failed) in {dt:.2f} s = {ok / dt:.0f} images/s "
          f"(CLIPB200_NVJPEG_THREADS={os.environ.get('CLIPB200_NVJPEG_THREADS', 'default')})", flush=True)
shutil.rmtree(tmp, ignore_errors=True)
