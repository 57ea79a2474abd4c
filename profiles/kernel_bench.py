"""Stand-alone timing of the non-GEMM tower kernels at batch-256 shapes (CUDA events, 50 launches)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import _native as N

L = N.lib()
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


B = 256
rows = B * 50
x = torch.randn((rows, 768), device="cuda").half()
h = torch.empty_like(x)
g = torch.randn(768, device="cuda"); b = torch.randn(768, device="cuda")
print(f"layernorm 12800x768: {timeit(lambda: N.check(L.cb_layernorm_f16_device(p(x), p(h), p(g), p(b), rows, 768, 1, None, None, 0, st()))):.1f} us")
qkv = torch.randn((rows, 2304), device="cuda").half()
att = torch.empty((rows, 768), dtype=torch.float16, device="cuda")
print(f"attention B=256 L=50 h=12: {timeit(lambda: N.check(L.cb_attention_f16_device(p(qkv), p(att), B, 50, 12, 0, st()))):.1f} us")
img = torch.randint(0, 256, (B, 224, 224, 3), device="cuda", dtype=torch.uint8)
pat = torch.empty((B * 49, 3072), dtype=torch.float16, device="cuda")
print(f"preprocess_u8 B=256: {timeit(lambda: N.check(L.cb_preprocess_u8_device(p(img), p(pat), B, st()))):.1f} us")
rows_t = 1024 * 77
xt = torch.randn((rows_t, 512), device="cuda").half(); ht = torch.empty_like(xt)
gt = torch.randn(512, device="cuda"); bt = torch.randn(512, device="cuda")
print(f"layernorm 78848x512: {timeit(lambda: N.check(L.cb_layernorm_f16_device(p(xt), p(ht), p(gt), p(bt), rows_t, 512, 1, None, None, 0, st()))):.1f} us")
qkvt = torch.randn((rows_t, 1536), device="cuda").half(); attt = torch.empty((rows_t, 512), dtype=torch.float16, device="cuda")
print(f"attention B=1024 L=77 h=8 causal: {timeit(lambda: N.check(L.cb_attention_f16_device(p(qkvt), p(attt), 1024, 77, 8, 1, st()))):.1f} us")
