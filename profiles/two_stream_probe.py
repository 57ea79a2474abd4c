"""Does running two model handles (two streams, two workspaces) concurrently raise aggregate
images/s?  (memory-bound LN/attention of one batch can hide under the other's GEMMs)"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))


def main():
    import torch
    from clipb200 import _native as N, clip, weights
    L = N.lib()
    sd = weights.synthetic_state_dict(0)
    B = 256
    nh = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    models = [clip.CLIPB200(sd, device=0, max_image_batch=B, max_text_batch=1) for _ in range(nh)]
    g = torch.Generator().manual_seed(0)
    host = [torch.randint(0, 256, (B, 224, 224, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(4)]
    outs = [torch.empty((B, 512)).pin_memory() for _ in range(4)]
    def run(n):
        for i in range(n):
            m = models[i % nh]
            N.check(L.cb_clip_submit_image_u8(m.handle, B, C.c_void_p(host[i % 4].data_ptr()), C.c_void_p(outs[i % 4].data_ptr()), 1))
        for m in models:
            N.check(L.cb_clip_sync(m.handle))
    run(8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 60
    run(n)
    dt = time.perf_counter() - t0
    print(f"{nh} handle(s): {n * B / dt:.0f} images/s end-to-end ({dt / n * 1e3:.3f} ms per batch)")


if __name__ == "__main__":
    main()
