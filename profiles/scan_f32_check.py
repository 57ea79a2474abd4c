"""Single-query scan over fp32-stored rows (the reference's own layout, 2048 B/row) vs fp16 rows:
achieved HBM GB/s of the whole search call (CUDA events, 100 searches)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import faiss, synth

dev = torch.device("cuda", 0)
for storage, rows in (("f32", 5_000_000), ("f16", 10_000_000)):
    index = faiss.IndexFlatIP(512, storage=storage, devices=[0])
    index.reserve(rows)
    step = 1 << 20
    for lo in range(0, rows, step):
        m = min(step, rows - lo)
        index.add_device(synth.device_unit_rows(m, 512, seed=lo, device=dev, dtype=torch.float32 if storage == "f32" else torch.float16))
    for nq in (1, 4, 8):
        q = synth.device_unit_rows(nq, 512, seed=7, device=dev, dtype=torch.float32)
        for _ in range(10):
            index.search_device(q, 100)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(100):
            index.search_device(q, 100)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 100
        bpr = 2048 if storage == "f32" else 1024
        passes = -(-nq // 4)
        print(f"{storage} rows={rows} nq={nq}: {ms:.3f} ms/search, {rows * bpr * passes / ms / 1e6:.0f} GB/s over {passes} pass(es)", flush=True)
    del index
    torch.cuda.empty_cache()
