#!/usr/bin/env python
"""query-index.py -- same prompt, commands and printed lines as CLI-P's query loop, with the two hot
calls (encode_text, index.search) on the GPU through clipb200 (cli-p_b200/clipb200/repl.py).
Reads vectors.lmdb and images.index from the current directory, like the reference.
Environment: CLIP_WEIGHTS (checkpoint), CLIPB200_STORAGE (f32|f16), CLIPB200_DEVICES (0,1,...),
CLIPB200_NO_VIEWER=1 (print results only; also the behaviour when cv2 is not installed)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cli-p_b200"))

from clipb200 import clip, faiss, indexer, lmdb, repl  # noqa: E402


def main() -> int:
    model, _ = clip.load("ViT-B/32", device="cuda", jit=False)
    model.eval()
    env = lmdb.open("vectors.lmdb", map_size=20 * 1024 ** 3, max_dbs=4)
    index = faiss.read_index("images.index")
    index.nprobe = 32
    show = None if os.environ.get("CLIPB200_NO_VIEWER") else repl.opencv_viewer()
    session = repl.QuerySession(indexer.Searcher(env, index, model), index, show=show)
    try:
        while session.handle(input(repl.PROMPT)):
            pass
    except (EOFError, KeyboardInterrupt):
        print("Interrupted.")
    env.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
