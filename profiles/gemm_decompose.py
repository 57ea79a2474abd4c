"""Where the tower GEMMs' time goes, per shape (run on the GPU box):
  full      the shipped kernel (tile picked by pick_config)
  nostore   gemm_debug=2: everything but the epilogue's TMA bulk stores
  mainloop  gemm_debug=1: TMA + MMA only (epilogue warps just release the accumulator)
  A-only    gemm_debug=3: full kernel, but W tiles are loaded for a cluster's first tile only (the
            operand traffic a W-stationary schedule would have; results are garbage)
  A-only+mainloop  gemm_debug=5: both of the above (is the mainloop bound by operand delivery?)
  cublas    torch's fp16 matmul (cuBLASLt) on the same operands, bias/activation/residual NOT included
CUDA events, 30 launches each after 5 warm-ups.  The gemm_debug probes corrupt results and exist only in
the experiments build:  python -m clipb200.build --experiments;  CLIPB200_LIB=<...>/libclipb200_exp.so python
profiles/gemm_decompose.py.  With the product library only `full` and `cublas` are reported."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import _native as N

SHAPES = [(12800, 2304, 768, 0, "vision qkv"), (12800, 768, 768, 2, "vision out_proj"),
          (12800, 3072, 768, 1, "vision c_fc"), (12800, 768, 3072, 2, "vision c_proj"),
          (12544, 768, 3072, 3, "patch embed"), (78848, 1536, 512, 0, "text qkv"),
          (78848, 512, 512, 2, "text out_proj"), (78848, 2048, 512, 1, "text c_fc"), (78848, 512, 2048, 2, "text c_proj")]
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (M, Nn, K, epi, name) in SHAPES:
    A = torch.randn((M, K), device="cuda").half()
    W = (torch.randn((Nn, K), device="cuda") * K ** -0.5).half()
    bias = torch.randn((Nn,), device="cuda")
    pos = torch.randn((50, Nn), device="cuda")
    out = torch.zeros((M + M // 49 + 2, Nn), dtype=torch.float16, device="cuda")
    res = out if epi == 2 else None
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        N.check(N.lib().cb_gemm_f16_device(M, Nn, K, p(A), p(W), p(bias), p(res), p(pos), p(out), Nn, epi, st))

    fl = 2 * M * Nn * K / 1e9
    line = f"{name:16s} M={M} N={Nn} K={K} epi={epi}:"
    HAVE_DBG = N.lib().cb_tuning_set(b"gemm_debug", -1) == 0
    variants = (("full", -1), ("nostore", 2), ("mainloop", 1), ("A-only", 3), ("A-only+mainloop", 5)) if HAVE_DBG else (("full", -1),)
    for label, dbg in variants:
        if HAVE_DBG:
            N.tuning_set("gemm_debug", dbg)
        ms = timeit(run)
        line += f"  {label} {ms * 1e3:6.1f}us {fl / ms:5.0f}TF"
    for bn in (128, 192, 256):
        if Nn % bn == 0:
            with N.tuning(gemm_bn=bn):
                ms = timeit(run)
            line += f"  bn{bn} {ms * 1e3:6.1f}us"
    if HAVE_DBG:
        N.tuning_set("gemm_debug", -1)
    Wt = W.t()
    o2 = torch.empty((M, Nn), dtype=torch.float16, device="cuda")
    ms = timeit(lambda: torch.matmul(A, Wt, out=o2))
    line += f"  cublas {ms * 1e3:6.1f}us {fl / ms:5.0f}TF"
    print(line, flush=True)
