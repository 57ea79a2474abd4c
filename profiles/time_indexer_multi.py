"""indexer.embed_folders from JPEG files with one replica per GPU (decode=nvjpeg).  usage: N ngpus"""
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

n, ng = int(sys.argv[1]), int(sys.argv[2])
tmp = tempfile.mkdtemp()
uniq = os.path.join(tmp, "uniq") + "/"
make_jpeg_folder(uniq, 1000)
folder = os.path.join(tmp, "photos") + "/"
os.makedirs(folder)
files = sorted(os.listdir(uniq))
for i in range(n):
    os.link(uniq + files[i % len(files)], folder + f"img_{i:07d}.jpg")
sd = weights.synthetic_state_dict(0)
os.chdir(tmp)
for g in sorted({1, ng}):
    models = [clip.CLIPB200(sd, device=d, max_image_batch=256, max_text_batch=1) for d in range(g)]
    m = models if g > 1 else models[0]
    indexer.embed_folders([uniq], lmdb.open(f"w{g}.lmdb", map_size=1 << 30, max_dbs=4), m, out=io.StringIO(), decode="nvjpeg")
    env = lmdb.open(f"v{g}.lmdb", map_size=8 << 30, max_dbs=4)
    t0 = time.perf_counter()
    ok, bad = indexer.embed_folders([folder], env, m, out=io.StringIO(), decode="nvjpeg")
    dt = time.perf_counter() - t0
    print(f"{g} GPU(s), {os.cpu_count()} host cores: {ok} files in {dt:.2f} s = {ok / dt:.0f} images/s", flush=True)
    del models, m
shutil.rmtree(tmp, ignore_errors=True)
