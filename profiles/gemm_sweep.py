"""Time every (cluster size, N tile) instantiation of the tcgen05 GEMM on the tower shapes
(run on the GPU box; prints TFLOP/s)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import _native as N

SHAPES = [(12800, 2304, 768, 0), (12800, 768, 768, 2), (12800, 3072, 768, 1), (12800, 768, 3072, 2),
          (12544, 768, 3072, 3), (78848, 1536, 512, 0), (78848, 512, 512, 2), (78848, 2048, 512, 1), (78848, 512, 2048, 2)]
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
ONLY = None
if len(sys.argv) > 1:      # profiles/gemm_sweep.py <shape index> <ncta> <bn>  (for ncu captures)
    ONLY = (int(sys.argv[2]), int(sys.argv[3]))
    SHAPES = [SHAPES[int(sys.argv[1])]]
for (M, Nn, K, epi) in SHAPES:
    A = torch.randn((M, K), device="cuda").half()
    W = (torch.randn((Nn, K), device="cuda") * K ** -0.5).half()
    bias = torch.randn((Nn,), device="cuda")
    pos = torch.randn((50, Nn), device="cuda")
    out = torch.zeros((M + M // 49 + 2, Nn), dtype=torch.float16, device="cuda")
    res = out if epi == 2 else None
    line = f"M={M} N={Nn} K={K} epi={epi}: "
    for ncta in (1, 2):
        for bn in (128, 192, 256):
            if Nn % bn or (ONLY and ONLY != (ncta, bn)):
                continue
            os.environ["CLIPB200_GEMM_BN"] = str(bn)
            os.environ["CLIPB200_GEMM_NCTA"] = str(ncta)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            def run():
                N.check(N.lib().cb_gemm_f16_device(M, Nn, K, p(A), p(W), p(bias), p(res), p(pos), p(out), Nn, epi, st))
            for _ in range(3):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            line += f" [{ncta}cta bn{bn}: {ms * 1e3:6.1f}us {2 * M * Nn * K / ms / 1e9:6.0f}TF]"
    print(line, flush=True)
