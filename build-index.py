#!/usr/bin/env python
"""build-index.py DIR/ [DIR/ ...] -- same command line, database names and output files as
CLI-P's builder, driven through clipb200's batched B200 pipeline (cli-p_b200/clipb200/indexer.py).
Environment: CLIP_WEIGHTS (checkpoint), CLIPB200_STORAGE (f32|f16), CLIPB200_DEVICES (0,1,...: the GPUs that
embed the files and hold the index shards),
CLIPB200_DECODE (pil = the reference's exact pixels, ~9 k files/s on large folders | nvjpeg = threaded GPU decode, ~30 k files/s,
stored vectors within cosine 0.999 of the pil path)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cli-p_b200"))

from clipb200 import clip, faiss, indexer, lmdb  # noqa: E402


def main(folders):
    devices = [int(t) for t in os.environ.get("CLIPB200_DEVICES", "").split(",") if t.strip()]
    if len(devices) > 1:      # one replica per GPU: the files of a folder are split over them (no collective)
        model = [clip.load("ViT-B/32", device=f"cuda:{d}", jit=False)[0] for d in devices]
    else:
        model, _ = clip.load("ViT-B/32", device=f"cuda:{devices[0]}" if devices else "cuda", jit=False)
    env = lmdb.open("vectors.lmdb", map_size=20 * 1024 ** 3, max_dbs=4)
    try:
        indexer.embed_folders(folders, env, model, decode=os.environ.get("CLIPB200_DECODE", "pil"))
    except KeyboardInterrupt:
        print("Interrupted!")          # like the reference, still build the index from what is stored
    indexer.build_index(env, faiss, index_path="images.index")
    env.close()


if __name__ == "__main__":
    main(sys.argv[1:])
