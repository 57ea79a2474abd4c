#!/usr/bin/env python
"""bench_configs.py -- the remaining BASELINE.json configs (bench.py covers configs[1] and the
single-query half of configs[2]).  One JSON line per config; same timing hygiene as bench.py
(warm-up, CUDA events on the launching stream, max over ranks, inputs larger than L2).

  python bench_configs.py [--configs 3,4,5] [--rows-per-gpu N]
  torchrun --nproc-per-node 8 bench_configs.py --configs 3,5

  configs[2] exact IP search over 10M x 512 fp16 sharded over the GPUs, k=100:
             single-query latency (p50/p99 per query) and batch-1024 throughput
  configs[3] image-similarity query: encode_image of one query image + top-100 over 10M
             vectors, end-to-end p50 latency (device-resident image, and from host pixels)
  configs[4] 100M x 512 fp16 (12.5M rows per GPU on 8 GPUs), batched text queries:
             encode_text of 1024 token rows + search, k=100.  On fewer than 8 GPUs the database
             is rows_per_gpu x n_gpus ("scaled": stated in the line).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "cli-p_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import dist_env, load_peaks  # noqa: E402


def pct(xs, p):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(round(p / 100.0 * (len(xs) - 1))))]


def max_over_ranks(v, world):
    if world == 1:
        return v
    t = torch.tensor([v], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def build_index(rows_local, rank, local, storage="f16"):
    from clipb200 import faiss, sharded, synth
    dev = torch.device("cuda", local)
    index = faiss.IndexFlatIP(512, storage=storage, devices=[local])
    index.reserve(rows_local)
    step = 1 << 21
    for lo in range(0, rows_local, step):
        m = min(step, rows_local - lo)
        x = synth.device_unit_rows(m, 512, seed=1000 + rank * 1000 + lo // step, device=dev, dtype=torch.float16)
        index.add_device(x)
        del x
    torch.cuda.synchronize()
    ds = sharded.DistributedFlatIP(index=index, device=dev)
    ds.finalize()
    return index, ds


def per_call_ms(fn, n, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        out.append(e0.elapsed_time(e1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--rows-per-gpu", type=int, default=None)
    args = ap.parse_args()
    want = {int(c) for c in args.configs.split(",")}
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    from clipb200 import clip, sharded, synth, weights
    peaks = load_peaks()

    def emit(d):
        if rank == 0:
            d.update({"n_gpus": world, "data": "synthetic"})
            print(json.dumps(d), flush=True)

    if want & {3, 4}:
        N = 10_000_000
        lo, hi = sharded.shard_range(N, rank, world)
        index, ds = build_index(hi - lo, rank, local)
        q1 = synth.device_unit_rows(1, 512, seed=7, device=dev, dtype=torch.float32)
        if 3 in want:
            lat = per_call_ms(lambda: ds.search(q1, 100), 200, 20)
            p50, p99 = max_over_ranks(pct(lat, 50), world), max_over_ranks(pct(lat, 99), world)
            qb = synth.device_unit_rows(1024, 512, seed=8, device=dev, dtype=torch.float32)
            bt = per_call_ms(lambda: ds.search(qb, 100), 5, 2)
            bms = max_over_ranks(pct(bt, 50), world)
            emit({"config": "configs[2]: exact IP search over 10M x 512 fp16, k=100, sharded over the GPUs",
                  "single_query_latency_ms": {"p50": p50, "p99": p99},
                  "single_query_hbm_gbs_per_gpu": (hi - lo) * 1024 / (p50 / 1e3) / 1e9,
                  "hbm_peak_gbs": peaks["hbm_gbs"],
                  "batch1024": {"ms": bms, "queries_per_s": 1024 / (bms / 1e3),
                                "tflops_algorithmic_total": 2.0 * 1024 * N * 512 / (bms / 1e3) / 1e12}})
        if 4 in want:
            sd = weights.synthetic_state_dict(0)
            model = clip.CLIPB200(sd, device=local, max_image_batch=1, max_text_batch=1)
            g = torch.Generator(device=dev).manual_seed(5)
            img = torch.randint(0, 256, (1, 224, 224, 3), generator=g, device=dev, dtype=torch.uint8)
            img_host = img.cpu().pin_memory()

            def image_query(src):
                f = model.encode_image(src, normalize=True)
                return ds.search(f, 100)

            lat_dev = per_call_ms(lambda: image_query(img), 100, 10)
            # from host pixels, result ids back on the host: wall clock
            walls = []
            for i in range(110):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                D, I = image_query(img_host.to(dev, non_blocking=True))
                I.cpu()
                if i >= 10:
                    walls.append((time.perf_counter() - t0) * 1e3)
            emit({"config": "configs[3]: image-similarity query = encode_image(1 image) + top-100 over 10M vectors",
                  "latency_ms_device_resident": {"p50": max_over_ranks(pct(lat_dev, 50), world),
                                                 "p99": max_over_ranks(pct(lat_dev, 99), world)},
                  "latency_ms_end_to_end_host_pixels_to_host_ids": {"p50": max_over_ranks(pct(walls, 50), world),
                                                                    "p99": max_over_ranks(pct(walls, 99), world)}})
            del model
        del index, ds
        torch.cuda.empty_cache()

    if 5 in want:
        rows = args.rows_per_gpu or 12_500_000
        index, ds = build_index(rows, rank, local)
        sd = weights.synthetic_state_dict(0)
        NQ = 1024
        per = -(-NQ // world)
        model = clip.CLIPB200(sd, device=local, max_image_batch=1, max_text_batch=per)
        from clipb200.synth import synthetic_tokens
        ids = synthetic_tokens(NQ, seed=9).to(dev)

        def text_batch():
            # queries sharded over ranks for the encode, one all-gather of the 1024 x 512 embeddings,
            # then the database-sharded search
            mine = ids[rank * per:min(NQ, (rank + 1) * per)]
            f = model.encode_text(mine, normalize=True)
            if world > 1:
                pad = torch.zeros((per, 512), device=dev)
                pad[:f.shape[0]] = f
                allf = torch.empty((world * per, 512), device=dev)
                dist.all_gather_into_tensor(allf, pad)
                f = allf[:NQ]
            return ds.search(f.contiguous(), 100)

        ts = per_call_ms(text_batch, 5, 2)
        ms = max_over_ranks(pct(ts, 50), world)
        enc = per_call_ms(lambda: model.encode_text(ids[rank * per:min(NQ, (rank + 1) * per)], normalize=True), 5, 2)
        emit({"config": "configs[4]: batched text queries (encode_text + search, k=100) over a 512-d fp16 database "
                        f"of {rows * world:,} rows" + ("" if rows * world == 100_000_000 else " (scaled: 12.5M rows per GPU)"),
              "rows_total": rows * world, "nq": NQ,
              "ms_per_batch": ms, "queries_per_s": NQ / (ms / 1e3),
              "encode_text_ms": max_over_ranks(pct(enc, 50), world),
              "encode_text_queries_per_s_per_gpu": per / (pct(enc, 50) / 1e3)})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
