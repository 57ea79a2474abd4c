"""GPU parity of the SHARDED search, the code path that carries
`index.search(features, k)` (/root/reference/query-index.py:111) once the database is split over
GPUs: per-shard kernel chain with global ids -> peer stores into the root's mailbox (or one NCCL
all-gather) -> merge kernel.  Compared with the CPU oracle and, bit for bit, with the unsharded
index.  Three settings: logical shards of one process on one device; one process per rank sharing
cuda:0 through cudaIpc (runs on the driver's 1-GPU box); torchrun over 2 physical GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import flatip_ref as F
from clipb200 import synth


def _torchrun(tmp_path, nproc, extra, timeout=600):
    out = tmp_path / "ok.txt"
    port = 29600 + (os.getpid() * 7 + nproc) % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "_sharded_worker.py"), "--out", str(out)] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    assert out.exists() and out.read_text().startswith("ok"), r.stderr[-3000:]
    return out.read_text()


@pytest.mark.parametrize("R", [2, 5])
def test_logical_shards_with_ties_across_boundaries(R):
    """One process, R shards on one device, rows added in several calls (several id segments per
    shard): equals the oracle and the unsharded index bit for bit, including exact ties whose
    members sit in different shards and k larger than a shard."""
    from clipb200 import faiss
    n = 30_011
    xb = synth.unit_rows(n, seed=61, clip_like=True)
    xb[[1, 5_000, 9_999, 10_000, 10_001, 20_500, n - 1]] = xb[17]
    xq = synth.unit_rows(24, seed=62, clip_like=True)
    xq[2] = xb[17]
    one = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    many = faiss.IndexFlatIP(512, storage="f16", devices=[0] * R)
    for lo in range(0, n, 10_000):
        one.add(xb[lo:lo + 10_000])
        many.add(xb[lo:lo + 10_000])
    assert many.ntotal == n
    xb16 = xb.astype(np.float16)
    for nq, k in ((1, 1), (1, 100), (3, 21), (4, 8), (24, 100), (2, 7000), (1, n + 5)):
        D0, I0 = one.search(xq[:nq], k)
        D1, I1 = many.search(xq[:nq], k)
        assert (I0 == I1).all(), f"nq={nq} k={k}"
        assert (D0.view(np.uint32) == D1.view(np.uint32)).all()
        Dref, Iref = F.search(xq[:nq], xb16, k)
        ok, _, msg = F.ids_match_with_tolerance(Dref, Iref, D1, I1, gap=1e-5)
        assert ok, msg
    assert list(many.search(xq[2:3], 8)[1][0]) == [1, 17, 5_000, 9_999, 10_000, 10_001, 20_500, n - 1]
    np.testing.assert_array_equal(many.reconstruct_n(9_990, 30), one.reconstruct_n(9_990, 30))
    # device-resident entry point: same answer, no host round trip
    import torch
    D2, I2 = many.search_device(torch.from_numpy(xq[:3]).cuda(), 21)
    D3, I3 = one.search(xq[:3], 21)
    assert (I2.cpu().numpy() == I3).all() and (D2.cpu().numpy().view(np.uint32) == D3.view(np.uint32)).all()


def test_distributed_world_of_one():
    """DistributedFlatIP without a process group (world 1) is the plain index."""
    import torch
    from clipb200 import faiss, sharded
    xb = synth.unit_rows(9_000, seed=63)
    xq = synth.unit_rows(3, seed=64)
    index = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    index.add(xb)
    ds = sharded.DistributedFlatIP(index=index, device=torch.device("cuda", 0))
    ds.finalize()
    D, I = ds.search(torch.from_numpy(xq).cuda(), 33)
    Dref, Iref = F.search(xq, xb.astype(np.float16), 33)
    ok, _, msg = F.ids_match_with_tolerance(Dref, Iref, D.cpu().numpy(), I.cpu().numpy(), gap=1e-5)
    assert ok, msg


def test_two_ranks_share_one_gpu_through_the_ipc_mailbox(tmp_path):
    """torchrun --nproc-per-node 2, both ranks on cuda:0 (gloo carries the 64-byte IPC handle):
    the multi-process peer-delivery path -- cudaIpc mailbox, release/acquire counters, slot
    back-pressure, merge -- with answers checked on the driver's 1-GPU box."""
    _torchrun(tmp_path, 2, ["--backend", "gloo", "--same-gpu", "--transports", "p2p"])


def test_torchrun_two_physical_gpus(tmp_path):
    """The bench's multi-GPU code path (one process per GPU, NCCL process group): NVLink peer
    delivery and the NCCL all-gather transport both equal the oracle and the unsharded index."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    _torchrun(tmp_path, 2, ["--backend", "nccl", "--transports", "p2p,nccl"])


def test_torchrun_four_physical_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs")
    _torchrun(tmp_path, 4, ["--backend", "nccl", "--transports", "p2p,nccl", "--rows", "70001"])
