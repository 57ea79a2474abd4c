"""Device time of the vision tower's attention at batch 256 (12 heads, L = 50): the tcgen05 pair kernel
against the mma.sync kernel (knob attn_tc = 0): CUDA events around a replayed CUDA graph of 48 launches
(no host launch cost in the number), once over eight rotating activations (472 MB: every launch reads HBM)
and once over one activation (59 MB: L2-hot).  In the real step the producer GEMM leaves ~40 % of qkv in L2.

    python profiles/attention_probe.py [B=256]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import _native as N


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    L, heads, W = 50, 12, 768
    dev = torch.device("cuda", 0)
    # eight rotating activations (59 MB each at B = 256): more than the 126 MB L2 between reuses
    bufs = [(torch.randn((B * L, 3 * W), device=dev) * 1.5).half() for _ in range(8)]
    out = torch.empty((B * L, W), dtype=torch.float16, device=dev)
    lib = N.lib()
    st = torch.cuda.current_stream().cuda_stream

    def run(n, rotate):
        nonlocal st
        for i in range(n):
            x = bufs[i % 8] if rotate else bufs[0]
            N.check(lib.cb_attention_f16_device(x.data_ptr(), out.data_ptr(), B, L, heads, 0, st))

    for name, knob in (("tcgen05 pair kernel", 1), ("mma.sync kernel", 0)):
        with N.tuning(attn_tc=knob):
            for rotate in (True, False):
                side = torch.cuda.Stream()
                with torch.cuda.stream(side):
                    st = side.cuda_stream
                    run(16, rotate)
                    side.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        st = torch.cuda.current_stream().cuda_stream
                        run(48, rotate)
                    g.replay()
                    side.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    g.replay()
                    e1.record()
                    e1.synchronize()
                us = e0.elapsed_time(e1) / 48 * 1e3
                mb = B * L * 4 * W * 2 / 1e6
                print(f"{name:20s} B={B} {'HBM-cold input' if rotate else 'L2-hot input  '}: {us:7.2f} us per launch "
                      f"({mb / us:.2f} TB/s of q,k,v,out)", flush=True)


def text():
    """The text tower's attention (L = 77, causal, 8 heads) at a batch of 256 prompts: mma.sync kernel."""
    B, L, heads, W = 256, 77, 8, 512
    dev = torch.device("cuda", 0)
    bufs = [(torch.randn((B * L, 3 * W), device=dev) * 1.5).half() for _ in range(4)]
    out = torch.empty((B * L, W), dtype=torch.float16, device=dev)
    lib = N.lib()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        def run(n):
            st = torch.cuda.current_stream().cuda_stream
            for i in range(n):
                N.check(lib.cb_attention_f16_device(bufs[i % 4].data_ptr(), out.data_ptr(), B, L, heads, 1, st))
        run(8)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            run(48)
        g.replay()
        side.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
    print(f"text tower, mma.sync kernel B={B} L=77 causal: {e0.elapsed_time(e1) / 48 * 1e3:7.2f} us per launch", flush=True)


if __name__ == "__main__":
    text()
    main()
