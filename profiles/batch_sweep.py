"""Sustained encode_image throughput vs batch size and lanes in flight (L2 residency of the
activation working set: ~0.69 MB per image per layer).  ~1 s per point, CUDA events."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import _native as N, weights as _weights
from clipb200.clip import CLIPB200

sd = _weights.synthetic_state_dict(0)
model = CLIPB200(sd, device=0, max_image_batch=256, max_text_batch=1)
L = N.lib()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
imgs = [torch.randint(0, 256, (256, 224, 224, 3), generator=g, device=dev, dtype=torch.uint8) for _ in range(4)]
outs = [torch.empty((256, 512), dtype=torch.float32, device=dev) for _ in range(2)]


def run(B, lanes, secs=1.0):
    i = [0]

    def step():
        j = i[0]
        i[0] += 1
        if lanes == 1:
            N.check(L.cb_clip_encode_image_u8_device(model.handle, B, C.c_void_p(imgs[j % 4].data_ptr()),
                                                     C.c_void_p(outs[0].data_ptr()), 1, model._stream()))
        else:
            N.check(L.cb_clip_submit_image_u8_device(model.handle, B, C.c_void_p(imgs[j % 4].data_ptr()),
                                                     C.c_void_p(outs[j % 2].data_ptr()), 1, model._stream()))

    def drain():
        if lanes == 2:
            N.check(L.cb_clip_join(model.handle, model._stream()))

    for _ in range(20):
        step()
    drain()
    torch.cuda.synchronize()
    # calibrate
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    drain()
    e1.record()
    torch.cuda.synchronize()
    n = max(20, int(secs * 1e3 / (e0.elapsed_time(e1) / 20)))
    e0.record()
    for _ in range(n):
        step()
    drain()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return B / ms * 1e3, ms


for B in (64, 96, 128, 160, 192, 256):
    for lanes in (1, 2):
        ips, ms = run(B, lanes)
        print(f"B={B:3d} lanes={lanes}: {ips:9.0f} img/s  {ms:.3f} ms/step", flush=True)
