"""Worker for tests/test_sharded_gpu.py, launched with torch.distributed.run (one process per rank).

Every rank owns a contiguous shard of the same seeded database on its GPU (or, with --same-gpu, all
ranks share cuda:0 -- the cudaIpc mailbox path can then be exercised on a 1-GPU box).  Rank 0 checks
every answer against the CPU oracle and, bit for bit, against an unsharded index of the same rows.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cli-p_b200")):
    sys.path.insert(0, p)

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--same-gpu", action="store_true")
    ap.add_argument("--transports", default="p2p,nccl")
    ap.add_argument("--rows", type=int, default=40_001)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = 0 if args.same_gpu else int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")

    from oracle import flatip_ref as F
    from clipb200 import faiss, sharded, synth

    n = args.rows
    xb = synth.unit_rows(n, seed=21, clip_like=True)
    lo1 = sharded.shard_range(n, 1, world)[0]
    xb[lo1 - 1] = xb[17]                    # exact ties that straddle the shard boundary
    xb[lo1 + 3] = xb[17]
    xb[n - 1] = xb[17]
    xb16 = xb.astype(np.float16)
    xq = synth.unit_rows(40, seed=22, clip_like=True)
    xq[1] = xb[17]
    lo, hi = sharded.shard_range(n, rank, world)
    index = faiss.IndexFlatIP(512, storage="f16", devices=[local])
    index.add(xb[lo:hi])
    whole = None
    if rank == 0:
        whole = faiss.IndexFlatIP(512, storage="f16", devices=[local])
        whole.add(xb)
    checked = 0
    for transport in args.transports.split(","):
        ds = sharded.DistributedFlatIP(index=index, device=dev, transport=transport, mailbox_elems=4096)
        ds.finalize()
        assert ds.id_base == lo and ds.ntotal_global == n
        # nq 1..4: streaming scan; nq 20/40: tensor-core batch path (shards >= 8192 rows);
        # 40 x 200 > 4096 mailbox elements: the query batch is chunked over several slots
        for nq, k in ((1, 1), (1, 100), (3, 21), (4, 100), (20, 50), (40, 200), (2, 5000)):
            if k > 4096 and transport == "p2p":
                continue
            q = torch.from_numpy(xq[:nq]).to(dev)
            D, I = ds.search(q, k)
            torch.cuda.synchronize()
            if transport == "nccl":
                assert D is not None and I is not None        # the collective leaves the answer everywhere
            elif rank != 0:
                assert D is None and I is None
            if D is not None and rank == 0:
                D, I = D.cpu().numpy(), I.cpu().numpy()
                Dref, Iref = F.search(xq[:nq], xb16, k)
                ok, _, msg = F.ids_match_with_tolerance(Dref, Iref, D, I, gap=1e-5)
                assert ok, f"{transport} nq={nq} k={k}: {msg}"
                np.testing.assert_allclose(D, Dref, atol=1e-5, rtol=0)
                D1, I1 = whole.search(xq[:nq], k)
                assert (I == I1).all(), f"{transport} nq={nq} k={k}: sharded ids differ from the unsharded index"
                assert (D.view(np.uint32) == D1.view(np.uint32)).all()
                checked += 1
        # a burst of searches with no host synchronisation in between: the two mailbox slots are
        # reused, the back-pressure counter must hold the peers back
        q = torch.from_numpy(xq[:2]).to(dev)
        outs = [ds.search(q, 64) for _ in range(25)]
        torch.cuda.synchronize()
        if rank == 0:
            D0, I0 = outs[0]
            for D, I in outs[1:]:
                assert torch.equal(D, D0) and torch.equal(I, I0)
            D1, I1 = whole.search(xq[:2], 64)
            assert (I0.cpu().numpy() == I1).all()
            checked += 1
        if transport == "p2p":
            Dh, Ih = ds.search_host(xq[:3], 21)               # host-buffer entry point
            if rank == 0:
                D1, I1 = whole.search(xq[:3], 21)
                assert (Ih == I1).all() and (Dh.view(np.uint32) == D1.view(np.uint32)).all()
            else:
                assert Dh is None and Ih is None
            # the pipelined form: 31 different queries in flight two at a time, each into its own output
            qs = [torch.from_numpy(xq[i % 40:i % 40 + 1]).to(dev) for i in range(31)]
            outs = [ds.submit(qq, 64) for qq in qs]
            ds.join()
            torch.cuda.synchronize()
            if rank == 0:
                for i, (D, I) in enumerate(outs):
                    D1, I1 = whole.search(xq[i % 40:i % 40 + 1], 64)
                    assert (I.cpu().numpy() == I1).all(), f"submit #{i}: ids differ from the unsharded index"
                    assert (D.cpu().numpy().view(np.uint32) == D1.view(np.uint32)).all()
                checked += 1
            assert ds.p2p_error() == 0
        dist.barrier()
    if rank == 0:
        with open(args.out, "w") as fh:
            fh.write(f"ok {checked}\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
