"""Small driver for ncu: a few batch-1024 searches over a 2M-row fp16 shard."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import faiss, synth
dev = torch.device("cuda", 0)
n = 2_000_000
index = faiss.IndexFlatIP(512, storage="f16", devices=[0])
index.reserve(n)
index.add_device(synth.device_unit_rows(n, 512, seed=1, device=dev, dtype=torch.float16))
q = synth.device_unit_rows(1024, 512, seed=2, device=dev, dtype=torch.float32)
for _ in range(3):
    D, I = index.search_device(q, 100)
torch.cuda.synchronize()
print("ok", D[0, :3].tolist())
