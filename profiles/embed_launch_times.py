"""In-pipeline duration of every GEMM launch of one encode_image step (batch 256), measured ON THE DEVICE
(%globaltimer min at CTA entry / max at CTA exit), and the gaps between consecutive GEMMs (which contain
the attention / LayerNorm kernels and the launch gaps).  One batch in flight.

    python profiles/embed_launch_times.py [steps=20]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import _native as N, weights
from clipb200.clip import CLIPB200


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    B = 256
    L = N.lib()
    model = CLIPB200(weights.synthetic_state_dict(0), device=0, max_image_batch=B, max_text_batch=1)
    g = torch.Generator(device="cuda").manual_seed(0)
    imgs = [torch.randint(0, 256, (B, 224, 224, 3), generator=g, device="cuda", dtype=torch.uint8) for _ in range(4)]
    out = torch.empty((B, 512), device="cuda")

    def step(i):
        N.check(L.cb_clip_encode_image_u8_device(model.handle, B, C.c_void_p(imgs[i % 4].data_ptr()),
                                                 C.c_void_p(out.data_ptr()), 1, model._stream()))
    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    L.cb_clip_timing(model.handle, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    cap = 50 * steps
    ms, t0 = (C.c_double * cap)(), (C.c_double * cap)()
    n = C.c_int(0)
    N.check(L.cb_clip_timing_launches(model.handle, ms, t0, cap, C.byref(n)))
    L.cb_clip_timing(model.handle, 0)
    names = ["patch"] + ["qkv", "out_proj", "c_fc", "c_proj"] * 12 + ["proj"]
    per = {}
    gaps = {}
    for s in range(1, n.value // 50):          # skip the first step
        for j in range(50):
            i = s * 50 + j
            per.setdefault(names[j], []).append(ms[i] * 1e3)
            if j > 0:
                gaps.setdefault(names[j - 1] + "->" + names[j], []).append((t0[i] - t0[i - 1] - ms[i - 1]) * 1e3)
    tot = e0.elapsed_time(e1) / steps
    print(f"one-lane step {tot * 1e3:.1f} us (CUDA events around {steps} steps)")
    gsum = 0.0
    for k, v in per.items():
        v.sort()
        cnt = {"patch": 1, "proj": 1}.get(k, 12)
        gsum += sum(v) / len(v) * cnt
        print(f"  {k:9s} median {v[len(v) // 2]:7.1f} us  mean {sum(v) / len(v):7.1f}  min {v[0]:7.1f}  max {v[-1]:7.1f}   x{cnt} per step")
    print(f"  GEMM device time per step {gsum:.1f} us")
    for k, v in gaps.items():
        v.sort()
        print(f"  gap {k:18s} median {v[len(v) // 2]:7.1f} us  (kernels between: " +
              {"qkv->out_proj": "attention", "patch->qkv": "ln_pre", "c_proj->proj": "ln_post"}.get(k, "none") + ")")


if __name__ == "__main__":
    main()
