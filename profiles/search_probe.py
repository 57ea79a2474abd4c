"""Search latency per query-batch size over a fp16 shard (CUDA events, median of `reps`), for the
crossover between the streaming scan and the tensor-core batch path and for ncu launch lists.

    python profiles/search_probe.py [rows=10000000] [nq list, e.g. 1,2,3,4,8,16,128,256,1024] [scan|batch|auto]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
import torch
from clipb200 import _native, faiss, synth


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    nqs = [int(t) for t in (sys.argv[2] if len(sys.argv) > 2 else "1,2,3,4,8,16,64,128,256,512,1024").split(",")]
    modes = (sys.argv[3] if len(sys.argv) > 3 else "scan,batch").split(",")
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 7
    dev = torch.device("cuda", 0)
    index = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    index.reserve(n)
    index.add_device(synth.device_unit_rows(n, 512, seed=1000, device=dev, dtype=torch.float16))
    torch.cuda.synchronize()
    for nq in nqs:
        q = synth.device_unit_rows(nq, 512, seed=8, device=dev, dtype=torch.float32)
        line = f"nq={nq:5d}:"
        for mode in modes:
            if mode == "scan" and nq > 64:
                continue
            knob = {"scan": 1 << 30, "batch": 1, "auto": -1}[mode]
            with _native.tuning(batch_min_nq=knob):
                for _ in range(2):
                    index.search_device(q, 100)
                torch.cuda.synchronize()
                ts = []
                for _ in range(reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    index.search_device(q, 100)
                    e1.record()
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            line += f"  {mode} {ms:8.3f} ms"
            if mode != "scan":
                line += f" ({2.0 * nq * n * 512 / ms / 1e9:7.1f} TFLOP/s, {n * 1024 / ms / 1e6:6.0f} GB/s)"
        print(line, flush=True)
    # phases inside the single-query search kernel
    import ctypes as C
    L = _native.lib()
    h = index._shards[0].handle
    q = synth.device_unit_rows(1, 512, seed=8, device=dev, dtype=torch.float32)
    L.cb_flatip_timing(h, 1)
    for k in (100, 1000, 5000):
        acc = [0.0, 0.0, 0.0]
        for _ in range(5):
            with _native.tuning(batch_min_nq=1 << 30):
                index.search_device(q, k)
            torch.cuda.synchronize()
            ms3 = (C.c_double * 3)()
            L.cb_flatip_phase_times(h, ms3)
            acc = [a + b / 5 for a, b in zip(acc, ms3)]
        print(f"nq=1 k={k}: scan pass {acc[0] * 1e3:8.1f} us   decide {acc[1] * 1e3:7.1f} us   gather+sort+write {acc[2] * 1e3:7.1f} us")
    L.cb_flatip_timing(h, 0)


if __name__ == "__main__":
    main()
