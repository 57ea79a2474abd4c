"""Single-query encode latency (the interactive path of query-index.py:107-108 and BASELINE
configs[3]): plain launches vs CUDA-graph replay.  Host wall clock around a synchronous call,
p50 over 200 calls after warm-up."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import numpy as np
import torch
from clipb200 import clip, weights

sd = weights.synthetic_state_dict(0)
m = clip.CLIPB200(sd, device=0, max_image_batch=32, max_text_batch=32)
g = torch.Generator().manual_seed(0)


def p50(fn, n=200):
    for _ in range(10):
        fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


for B in (1, 4, 16, 32):
    img = torch.randint(0, 256, (B, 224, 224, 3), generator=g, dtype=torch.uint8)
    ids = torch.zeros((B, 77), dtype=torch.int32)
    ids[:, 0] = 49406
    ids[:, 1:6] = torch.randint(1000, 40000, (B, 5), generator=g, dtype=torch.int32)
    ids[:, 6] = 49407
    img_np, ids_np = img.numpy(), ids.numpy()
    t_img = p50(lambda: m.encode_image_u8_host(img_np))
    t_txt = p50(lambda: m.encode_text_host(ids_np))
    print(f"{'graph' if not os.environ.get('CLIPB200_NO_GRAPH') else 'plain'} B={B:2d}: "
          f"encode_image (host u8 -> host f32) p50 {t_img:.3f} ms   encode_text (host ids -> host f32) p50 {t_txt:.3f} ms",
          flush=True)
