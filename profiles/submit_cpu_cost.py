"""CPU cost of queuing one batch (90 launches + tensor-map encodes) vs GPU time per batch."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import _native as N, clip, weights
L = N.lib()
sd = weights.synthetic_state_dict(0)
B = 256
m = clip.CLIPB200(sd, device=0, max_image_batch=B, max_text_batch=1)
img = torch.randint(0, 256, (B, 224, 224, 3), device="cuda", dtype=torch.uint8)
outs = [torch.empty((B, 512), device="cuda") for _ in range(2)]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def submit(i):
    N.check(L.cb_clip_submit_image_u8_device(m.handle, B, C.c_void_p(img.data_ptr()), C.c_void_p(outs[i % 2].data_ptr()), 1, st))
for i in range(6):
    submit(i)
N.check(L.cb_clip_join(m.handle, st)); torch.cuda.synchronize()
n = 40
t0 = time.perf_counter()
for i in range(n):
    submit(i)
t_cpu = time.perf_counter() - t0
N.check(L.cb_clip_join(m.handle, st)); torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"CPU time to queue one batch: {t_cpu / n * 1e3:.3f} ms; wall per batch incl. GPU: {t_all / n * 1e3:.3f} ms")
