"""GPU parity: the CUDA flat-IP path, called through the C ABI (ctypes), against the
CPU oracle on the same seeded inputs.

Bar (BASELINE.json north star): ids identical to exact flat IP except where adjacent
reference scores differ by < 1e-5; recall@k >= 0.999; scores within 1e-5 (fp32
summation order differs from BLAS).  Both arms see the identical fp16-rounded rows.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import flatip_ref as F
from clipb200 import synth

SCORE_ATOL = 1e-5


def _check(D, I, Dref, Iref):
    ok, exempt, msg = F.ids_match_with_tolerance(Dref, Iref, D, I, gap=1e-5)
    assert ok, msg
    assert F.recall_at_k(Iref, I) >= 0.999
    valid = Iref >= 0
    np.testing.assert_allclose(D[valid], Dref[valid], atol=SCORE_ATOL, rtol=0)
    assert (D[~valid] == F.NEG_FLT_MAX).all() and (I[~valid] == -1).all()
    assert (np.diff(D, axis=1) <= 0).all(), "scores not sorted descending"


@pytest.fixture(scope="module")
def faiss():
    from clipb200 import faiss as f
    return f


@pytest.fixture(scope="module")
def db100k():
    xb = synth.unit_rows(100_000, seed=1000, clip_like=True)
    xb[500:504] = xb[17]                      # exact ties
    xb[99_999] = xb[17]
    return xb


@pytest.mark.parametrize("storage", ["f16", "f32"])
def test_search_matches_oracle(faiss, db100k, storage):
    xb_store = db100k.astype(np.float16) if storage == "f16" else db100k
    index = faiss.IndexFlatIP(512, storage=storage)
    index.add(db100k)
    assert index.ntotal == len(db100k)
    xq = synth.unit_rows(16, seed=7, clip_like=True)
    xq[3] = db100k[17]
    for nq in (1, 2, 3, 4, 5, 16):
        for k in (1, 21, 51, 100, 1000):
            if nq > 5 and k not in (21, 100):
                continue
            D, I = index.search(xq[:nq], k)
            assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (nq, k)
            Dref, Iref = F.search(xq[:nq], xb_store, k)
            _check(D, I, Dref, Iref)


def test_golden_fixture(faiss):
    import importlib.util
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(gdir, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    xb16, xq = m.flatip_inputs()
    g = np.load(os.path.join(gdir, "flatip_golden.npz"))
    index = faiss.IndexFlatIP(512, storage="f16")
    index.add(xb16.astype(np.float32))
    for k in (1, 21, 100):
        D, I = index.search(xq, k)
        _check(D, I, g[f"D{k}"], g[f"I{k}"])
    # exact duplicates of the query row come back in id order
    assert list(index.search(xq[5:6], 9)[1][0]) == [7, 100, 101, 102, 103, 104, 105, 106, 107]


@pytest.mark.parametrize("n", [1, 3, 7, 8, 9, 31, 1001])
def test_ragged_sizes_and_padding(faiss, n):
    xb = synth.unit_rows(n, seed=n)
    xq = synth.unit_rows(2, seed=77)
    for storage in ("f16", "f32"):
        index = faiss.IndexFlatIP(512, storage=storage)
        index.add(xb)
        store = xb.astype(np.float16) if storage == "f16" else xb
        for k in (1, n, n + 5, 100):
            D, I = index.search(xq, k)
            Dref, Iref = F.search(xq, store, k)
            _check(D, I, Dref, Iref)


def test_empty_index(faiss):
    index = faiss.IndexFlatIP(512)
    D, I = index.search(synth.unit_rows(3, seed=1), 7)
    assert (I == -1).all() and (D == F.NEG_FLT_MAX).all()


def test_all_equal_scores(faiss):
    """Every row ties: the k lowest ids win (exercises the exact-tie path)."""
    for n, k in ((5000, 100), (5000, 5000), (70_000, 33)):
        index = faiss.IndexFlatIP(512, storage="f16")
        index.add(np.zeros((n, 512), np.float32))
        D, I = index.search(synth.unit_rows(2, seed=3), k)
        assert (I == np.arange(k)[None, :]).all()
        assert (D == 0).all()


def test_many_duplicates_at_the_boundary(faiss):
    xb = synth.unit_rows(20_000, seed=5)
    xb[1000:3000] = xb[1000]                  # 2000 identical rows
    xq = xb[1000:1001].copy()
    index = faiss.IndexFlatIP(512, storage="f32")
    index.add(xb)
    D, I = index.search(xq, 50)
    assert list(I[0]) == list(range(1000, 1050))


def test_large_k_global_sort(faiss):
    xb = synth.unit_rows(20_000, seed=9)
    xq = synth.unit_rows(2, seed=10)
    index = faiss.IndexFlatIP(512, storage="f16")
    index.add(xb)
    for k in (4096, 5000, 20_000):
        D, I = index.search(xq, k)
        Dref, Iref = F.search(xq, xb.astype(np.float16), k)
        _check(D, I, Dref, Iref)


def test_incremental_add_reconstruct_and_file_roundtrip(faiss, tmp_path):
    xb = synth.unit_rows(3000, seed=2)
    index = faiss.IndexFlatIP(512, storage="f32")
    for lo in range(0, 3000, 700):
        index.add(xb[lo:lo + 700])
    assert index.ntotal == 3000
    np.testing.assert_array_equal(index.reconstruct_n(0, 3000), xb)
    np.testing.assert_array_equal(index.reconstruct(1234), xb[1234])
    # the reference's construction (build-index.py:80-81,96,99): quantizer, IVF wrapper, train, add
    quantizer = faiss.IndexFlatIP(512, storage="f32")
    ivf = faiss.IndexIVFFlat(quantizer, 512, 100, faiss.METRIC_INNER_PRODUCT)
    with pytest.raises(AssertionError):
        ivf.add(xb)                                      # faiss refuses add() before train() too
    ivf.train(xb)
    ivf.add(xb)
    assert ivf.ntotal == 3000 and quantizer.ntotal == 0   # the quantizer is not the row store
    ivf.nprobe = 32
    path = str(tmp_path / "images.index")
    faiss.write_index(ivf, path)
    assert open(path, "rb").read(4) == b"IwFl"           # faiss's own IndexIVFFlat container
    back = faiss.read_index(path)
    assert isinstance(back, faiss.IndexIVFFlat) and back.nprobe == 32 and back.ntotal == 3000
    xq = synth.unit_rows(2, seed=4)
    D0, I0 = ivf.search(xq, 21)
    D1, I1 = back.search(xq, 21)
    assert (I0 == I1).all() and (D0 == D1).all()
    # plain flat index -> "IxFI"; fp16 storage on the way back in
    fpath = str(tmp_path / "flat.index")
    faiss.write_index(index, fpath)
    assert open(fpath, "rb").read(4) == b"IxFI"
    flat = faiss.read_index(fpath)
    assert isinstance(flat, faiss.IndexFlatIP) and flat.ntotal == 3000
    np.testing.assert_array_equal(flat.reconstruct_n(0, 3000), xb)
    D2, I2 = flat.search(xq, 21)
    assert (I0 == I2).all() and (D0 == D2).all()
    half = faiss.read_index(fpath, storage="f16")
    assert half.storage == "f16" and half.ntotal == 3000
    # a k-means-partitioned file as the reference's build-index.py writes it (nlist 100, ids scattered
    # over the lists) flattens back into add order
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_faiss_io import ivf_bytes
    assign = np.random.default_rng(0).integers(0, 100, size=3000)
    kpath = str(tmp_path / "kmeans.index")
    open(kpath, "wb").write(ivf_bytes(xb, assign, 100, 32))
    km = faiss.read_index(kpath)
    assert isinstance(km, faiss.IndexIVFFlat) and km.nlist == 100 and km.nprobe == 32
    D3, I3 = km.search(xq, 21)
    assert (I0 == I3).all() and (D0 == D3).all()
    # fp16 storage rounds rows exactly like numpy
    h = faiss.IndexFlatIP(512, storage="f16")
    h.add(xb)
    np.testing.assert_array_equal(h.reconstruct_n(0, 3000), xb.astype(np.float16).astype(np.float32))


def test_wrapper_argument_errors(faiss):
    index = faiss.IndexFlatIP(512)
    index.add(synth.unit_rows(10, seed=1))
    with pytest.raises(AssertionError):
        index.search(np.zeros((1, 100), np.float32), 5)
    with pytest.raises(AssertionError):
        index.search(np.zeros((1, 512), np.float32), 0)
    with pytest.raises(TypeError):
        index.add(np.zeros((1, 512), np.float64))


def test_merge_kernel_matches_oracle(faiss):
    import torch
    rng = np.random.default_rng(0)
    R, nq, k = 8, 3, 100
    Ds = np.sort(rng.standard_normal((R, nq, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    Is = rng.permutation(R * nq * k).reshape(R, nq, k).astype(np.int64)
    Ds[2, :, 60:] = F.NEG_FLT_MAX
    Is[2, :, 60:] = -1                      # a short shard
    Ds[5, 1, :] = Ds[4, 1, :]               # cross-shard exact ties
    for kk in (100, 37):      # per-shard lists and the merged list have the same length k
        Dk, Ik = Ds[:, :, :kk].copy(), Is[:, :, :kk].copy()
        D, I = faiss.merge_topk_device(torch.from_numpy(Dk).cuda(), torch.from_numpy(Ik).cuda(), kk)
        Dref, Iref = F.merge_topk(Dk, Ik, kk)
        assert (I.cpu().numpy() == Iref).all() and (D.cpu().numpy() == Dref).all()
    # very large R*k goes through the global-memory sort
    R, nq, k = 8, 2, 2000
    Ds = np.sort(rng.standard_normal((R, nq, k)).astype(np.float32), axis=2)[:, :, ::-1].copy()
    Is = rng.permutation(R * nq * k).reshape(R, nq, k).astype(np.int64)
    D, I = faiss.merge_topk_device(torch.from_numpy(Ds).cuda(), torch.from_numpy(Is).cuda(), k)
    Dref, Iref = F.merge_topk(Ds, Is, k)
    assert (I.cpu().numpy() == Iref).all() and (D.cpu().numpy() == Dref).all()


@pytest.mark.parametrize("R", [2, 3, 8])
def test_logical_shards_equal_single_shard_bitwise(faiss, db100k, R):
    """R logical shards on one device (shard/merge must be testable on
    1 GPU): the merged answer is bit-identical to the unsharded answer."""
    xq = synth.unit_rows(4, seed=7, clip_like=True)
    xq[1] = db100k[17]
    one = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    many = faiss.IndexFlatIP(512, storage="f16", devices=[0] * R)
    n = 50_001
    for lo in range(0, n, 20_000):                      # several add() calls -> several segments
        one.add(db100k[lo:min(lo + 20_000, n)])
        many.add(db100k[lo:min(lo + 20_000, n)])
    assert many.ntotal == one.ntotal == n
    for k in (1, 21, 100):
        D0, I0 = one.search(xq, k)
        D1, I1 = many.search(xq, k)
        assert (I0 == I1).all(), "sharded ids differ from single-shard ids"
        assert (D0.view(np.uint32) == D1.view(np.uint32)).all()
    np.testing.assert_array_equal(many.reconstruct_n(19_990, 30), one.reconstruct_n(19_990, 30))


def test_pipelined_submit_equals_blocking_search(faiss, db100k):
    """cb_flatip_submit_search_device / cb_flatip_join (two lanes, half the SMs' residency each): a stream of
    different queries returns exactly what the blocking search returns, for single queries (scan kernel) and
    small batches (tensor-core path)."""
    import torch
    from clipb200 import sharded
    index = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    index.add(db100k)
    ds = sharded.DistributedFlatIP(index=index, device=torch.device("cuda", 0))
    ds.finalize()
    xq = synth.unit_rows(64, seed=71, clip_like=True)
    for nq, k in ((1, 100), (1, 7), (3, 21), (32, 50)):
        qs = [torch.from_numpy(xq[i:i + nq].copy()).cuda() for i in range(0, 64 - nq + 1, max(nq, 3))]
        outs = [ds.submit(q, k) for q in qs]
        ds.join()
        torch.cuda.synchronize()
        for q, (D, I) in zip(qs, outs):
            D0, I0 = index.search(q.cpu().numpy(), k)
            assert (I.cpu().numpy() == I0).all()
            assert (D.cpu().numpy().view(np.uint32) == D0.view(np.uint32)).all()


def test_full_size_properties():
    """BASELINE config 3 scale on one GPU (10M x 512 fp16 = 10.24 GB): size-independent
    properties -- sorted, unique ids, returned scores re-derive from the rows, and exactly
    rank-many rows beat the k-th score."""
    import torch
    from clipb200 import faiss
    n, k = 10_000_000, 100
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("not enough free HBM for the 10M-row case")
    dev = torch.device("cuda", 0)
    rows = synth.device_unit_rows(n, 512, seed=1000, device=dev, dtype=torch.float16)
    index = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    index.reserve(n)
    index.add_device(rows)
    q = synth.device_unit_rows(3, 512, seed=7, device=dev, dtype=torch.float32)
    D, I = index.search_device(q, k)
    torch.cuda.synchronize()
    D, I = D.cpu(), I.cpu()
    assert (D[:, 1:] <= D[:, :-1]).all()
    for qi in range(3):
        assert len(set(I[qi].tolist())) == k
        got = rows[I[qi].to(dev)].float() @ q[qi]
        assert torch.allclose(got.cpu(), D[qi], atol=SCORE_ATOL, rtol=0)
        # threshold scan with an independent fp32 matmul: how many rows beat the k-th score?
        kth = D[qi, -1].item()
        above = 0
        for lo in range(0, n, 1 << 20):
            s = rows[lo:lo + (1 << 20)].float() @ q[qi]
            above += int((s > kth + SCORE_ATOL).sum())
        assert above <= k - 1, f"{above} rows beat the returned k-th score"
    # host-pointer entry point returns the same answer
    Dh, Ih = index.search(q.cpu().numpy(), k)
    assert (Ih == I.numpy()).all() and (Dh == D.numpy()).all()


def test_two_physical_gpus_equal_one(db100k):
    """Real multi-device sharding inside one process (the REPL's mode): needs >= 2 GPUs."""
    import torch
    from clipb200 import faiss
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    xq = synth.unit_rows(3, seed=7, clip_like=True)
    one = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    two = faiss.IndexFlatIP(512, storage="f16", devices=[0, 1])
    one.add(db100k)
    two.add(db100k)
    for k in (1, 100):
        D0, I0 = one.search(xq, k)
        D1, I1 = two.search(xq, k)
        assert (I0 == I1).all() and (D0.view(np.uint32) == D1.view(np.uint32)).all()
