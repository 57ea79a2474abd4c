"""Where the tensor-core batch path overtakes the streaming scan (10M x 512 fp16, k=100)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import faiss, synth

dev = torch.device("cuda", 0)
rows = 10_000_000
index = faiss.IndexFlatIP(512, storage="f16", devices=[0])
index.reserve(rows)
step = 1 << 20
for lo in range(0, rows, step):
    index.add_device(synth.device_unit_rows(min(step, rows - lo), 512, seed=lo, device=dev, dtype=torch.float16))


def t(nq, n=20):
    q = synth.device_unit_rows(nq, 512, seed=7, device=dev, dtype=torch.float32)
    for _ in range(3):
        index.search_device(q, 100)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        index.search_device(q, 100)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for nq in (2, 4, 5, 6, 8, 12, 16, 32, 64, 128, 256):
    os.environ["CLIPB200_BATCH_MIN_NQ"] = "100000"
    scan = t(nq)
    os.environ["CLIPB200_BATCH_MIN_NQ"] = "1"
    batch = t(nq)
    print(f"nq={nq:4d}: scan {scan:7.3f} ms   tensor-core batch {batch:7.3f} ms", flush=True)
