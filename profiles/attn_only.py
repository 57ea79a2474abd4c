"""Launch the vision attention kernel a few times (for ncu)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import _native as N

L = N.lib()
B = 256
qkv = torch.randn((B * 50, 2304), device="cuda").half()
att = torch.empty((B * 50, 768), dtype=torch.float16, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(6):
    N.check(L.cb_attention_f16_device(C.c_void_p(qkv.data_ptr()), C.c_void_p(att.data_ptr()), B, 50, 12, 0, st))
torch.cuda.synchronize()
print("ok")
