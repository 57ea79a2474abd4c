"""cProfile of the index-time pipeline (main thread) over a folder of synthetic JPEGs."""
import cProfile
import io
import os
import pstats
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from bench_configs import make_jpeg_folder
from clipb200 import clip, indexer, lmdb, weights

mode = sys.argv[1] if len(sys.argv) > 1 else "nvjpeg"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
tmp = tempfile.mkdtemp()
folder = os.path.join(tmp, "photos") + "/"
make_jpeg_folder(folder, n)
model = clip.CLIPB200(weights.synthetic_state_dict(0), device=0, max_image_batch=256, max_text_batch=1)
kw = {"decode": "nvjpeg"} if mode == "nvjpeg" else {}
os.chdir(tmp)
warm = os.path.join(tmp, "w") + "/"
os.makedirs(warm)
for fn in sorted(os.listdir(folder))[:64]:
    shutil.copy(folder + fn, warm + fn)
indexer.embed_folders([warm], lmdb.open("w.lmdb", map_size=1 << 30, max_dbs=4), model, out=io.StringIO(), **kw)
env = lmdb.open("vectors.lmdb", map_size=1 << 30, max_dbs=4)
pr = cProfile.Profile()
pr.enable()
indexer.embed_folders([folder], env, model, out=io.StringIO(), **kw)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
shutil.rmtree(tmp, ignore_errors=True)
