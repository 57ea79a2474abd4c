"""CPU, world_size 2 over gloo: the host logic of the database-sharded search
(row partition, global id bases, packed single-collective gather) with the oracle
standing in for the per-rank GPU scan and the GPU merge kernel."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, k, nq, out_dir):
    for p in (ROOT, os.path.join(ROOT, "cli-p_b200")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import flatip_ref as F
    from clipb200 import sharded, synth

    xb = synth.unit_rows(n, seed=21, clip_like=True).astype(np.float16)
    xb[50] = xb[n - 1]              # a tie that straddles the two shards
    xq = synth.unit_rows(nq, seed=22, clip_like=True)
    lo, hi = sharded.shard_range(n, rank, world)
    shard = xb[lo:hi]

    def local_search(q, kk, D, I, id_base):
        d, i = F.search(q.numpy(), shard, kk)
        D.copy_(torch.from_numpy(d))
        I.copy_(torch.from_numpy(np.where(i >= 0, i + id_base, i)))

    def merge(gathered, R, nq_, kk, off_I, stride):
        g = gathered.view(R, stride)
        Ds = np.stack([g[r, :nq_ * kk * 4].view(torch.float32).view(nq_, kk).numpy() for r in range(R)])
        Is = np.stack([g[r, off_I:].view(torch.int64).view(nq_, kk).numpy() for r in range(R)])
        d, i = F.merge_topk(Ds, Is, kk)
        return torch.from_numpy(d), torch.from_numpy(i)

    s = sharded.DistributedFlatIP(local_search=local_search, merge=merge, device=torch.device("cpu"))
    s.finalize(n_local=hi - lo)
    assert s.id_base == lo and s.ntotal_global == n
    D, I = s.search(torch.from_numpy(xq), k)
    Dref, Iref = F.search(xq, xb, k)
    ok, _, msg = F.ids_match_with_tolerance(Dref, Iref, D.numpy(), I.numpy())
    assert ok, "sharded ids differ from unsharded: " + msg
    np.testing.assert_allclose(D.numpy(), Dref, atol=1e-6)   # BLAS block shape: ulp-level only
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,k,nq", [(1001, 10, 3), (64, 100, 1)])
def test_two_rank_sharded_search(tmp_path, n, k, nq):
    port = 29500 + (os.getpid() + n) % 2000
    mp.spawn(_worker, args=(2, port, n, k, nq, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0.npy") and os.path.exists(tmp_path / "ok1.npy")


def test_shard_range_covers_everything():
    from clipb200 import sharded
    for n in (0, 1, 7, 8, 1001, 10_000_000):
        for w in (1, 2, 3, 8):
            rs = [sharded.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            for a, b in zip(rs, rs[1:]):
                assert a[1] == b[0]


def test_packed_layout_is_aligned():
    from clipb200 import sharded
    for nq, k in ((1, 1), (1, 101), (3, 21), (1024, 100)):
        off, tot = sharded.packed_bytes(nq, k)
        assert off % 8 == 0 and tot % 8 == 0 and off >= nq * k * 4 and tot == off + nq * k * 8
