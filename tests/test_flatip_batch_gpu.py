"""GPU parity for the tensor-core query-batch path (tcgen05 GEMM of the fp16-rounded queries with a
fused threshold filter + exact fp32 re-scoring), through the C ABI, against the CPU oracle.  Same
bar as the scan path: ids identical except near-ties (< 1e-5), recall >= 0.999, scores within 1e-5;
and against the scan path itself the answer is bit-identical."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import flatip_ref as F
from clipb200 import synth


def _check(D, I, Dref, Iref):
    ok, exempt, msg = F.ids_match_with_tolerance(Dref, Iref, D, I, gap=1e-5)
    assert ok, msg
    assert F.recall_at_k(Iref, I) >= 0.999
    valid = Iref >= 0
    np.testing.assert_allclose(D[valid], Dref[valid], atol=1e-5, rtol=0)
    assert (np.diff(D, axis=1) <= 0).all()


def _stats(index):
    from clipb200 import _native as N
    a, b = C.c_int64(0), C.c_int64(0)
    N.check(N.lib().cb_flatip_batch_stats(index._shards[0].handle, C.byref(a), C.byref(b)))
    return a.value, b.value


@pytest.fixture(scope="module")
def setup():
    from clipb200 import faiss
    xb = synth.unit_rows(100_003, seed=31, clip_like=True)
    xb[700:704] = xb[17]
    xb[100_002] = xb[17]
    index = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    index.add(xb)
    xq = synth.unit_rows(1024, seed=32, clip_like=True)
    xq[5] = xb[17]
    return index, xb.astype(np.float16), xq


@pytest.mark.parametrize("nq", [16, 100, 128, 129, 256, 300, 1024, 1100])
def test_batch_path_matches_oracle(setup, nq):
    index, xb16, xq = setup
    before, _ = _stats(index)
    for k in ((1, 21, 100, 1000) if nq <= 300 else (100,)):
        D, I = index.search(xq[:nq], k)
        Dref, Iref = F.search(xq[:nq], xb16, k)
        _check(D, I, Dref, Iref)
    after, rescued = _stats(index)
    assert after > before, "the tensor-core batch path did not run"
    assert rescued == 0


@pytest.mark.parametrize("nq,k", [(64, 50), (200, 100), (17, 1)])
def test_batch_equals_scan_path_bitwise(setup, nq, k):
    """The batch path re-scores its survivors with the scan kernel's summation order: ids AND
    scores are bit-identical between the two paths."""
    from clipb200 import _native
    index, xb16, xq = setup
    q = np.tile(xq, (2, 1))[:nq]
    D1, I1 = index.search(q, k)
    with _native.tuning(batch_min_nq=100000):
        D0, I0 = index.search(q, k)
    assert (I0 == I1).all()
    assert (D0.view(np.uint32) == D1.view(np.uint32)).all()


def test_queries_outside_fp16_range_are_still_exact():
    """A query the tensor cores cannot hold (|q_i| > 65504 after rounding) takes the exact
    on-device selection instead of the filter."""
    from clipb200 import faiss
    xb = synth.unit_rows(20_000, seed=51)
    xq = synth.unit_rows(24, seed=52)
    xq[3] *= 1e7
    index = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    index.add(xb)
    index.add(xb)                                        # 40k rows: two row ranges per query
    xb2 = np.concatenate([xb, xb]).astype(np.float16)
    D, I = index.search(xq, 21)
    Dref, Iref = F.search(xq, xb2, 21)
    ok, _, msg = F.ids_match_with_tolerance(Dref[3:4], Iref[3:4], D[3:4], I[3:4], gap=1e-5 * 1e7)
    assert ok, msg
    np.testing.assert_allclose(D[3], Dref[3], rtol=1e-5)
    keep = [i for i in range(24) if i != 3]
    _check(D[keep], I[keep], Dref[keep], Iref[keep])
    assert _stats(index)[1] >= 1                         # the out-of-range query went through the exact path


def test_adversarial_order_is_rescued_exactly():
    """Rows sorted by ascending score for every query direction: each row beats the running
    threshold, the candidate lists overflow, and the per-query blocks must re-select exactly
    on the device (no host round trip, no fallback launch)."""
    from clipb200 import faiss
    rng = np.random.default_rng(0)
    u = rng.standard_normal(512).astype(np.float32)
    u /= np.linalg.norm(u)
    n = 60_000
    t = np.linspace(-0.9, 0.9, n, dtype=np.float32)
    noise = synth.unit_rows(n, seed=3) * 0.01
    xb = (t[:, None] * u[None, :] + noise).astype(np.float32)
    xq = np.repeat(u[None, :], 32, axis=0) + synth.unit_rows(32, seed=4) * 0.001
    xq = xq.astype(np.float32)
    index = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    index.add(xb)
    D, I = index.search(xq, 100)
    Dref, Iref = F.search(xq, xb.astype(np.float16), 100)
    _check(D, I, Dref, Iref)
    ran, rescued = _stats(index)
    assert ran >= 1 and rescued >= 1, "expected the on-device rescue to trigger"


def test_small_and_ragged_shards():
    from clipb200 import faiss
    for n in (8192, 8193, 12_345):
        xb = synth.unit_rows(n, seed=n)
        xq = synth.unit_rows(40, seed=n + 1)
        index = faiss.IndexFlatIP(512, storage="f16", devices=[0])
        index.add(xb)
        for k in (1, 100):
            D, I = index.search(xq, k)
            Dref, Iref = F.search(xq, xb.astype(np.float16), k)
            _check(D, I, Dref, Iref)


def test_batch_over_logical_shards_equals_single(setup):
    from clipb200 import faiss
    index, xb16, xq = setup
    many = faiss.IndexFlatIP(512, storage="f16", devices=[0, 0, 0])
    many.add(xb16.astype(np.float32))
    D0, I0 = index.search(xq[:48], 100)
    D1, I1 = many.search(xq[:48], 100)
    assert (I0 == I1).all()
    assert (D0.view(np.uint32) == D1.view(np.uint32)).all()


def test_batch_path_on_a_second_device():
    """Kernel attributes (opt-in shared memory) are per device: the tensor-core path must also
    work on cuda:1 in a process that already used cuda:0."""
    import torch
    from clipb200 import faiss
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    xb = synth.unit_rows(20_000, seed=41)
    xq = synth.unit_rows(32, seed=42)
    a = faiss.IndexFlatIP(512, storage="f16", devices=[0])
    b = faiss.IndexFlatIP(512, storage="f16", devices=[1])
    a.add(xb)
    b.add(xb)
    D0, I0 = a.search(xq, 50)
    D1, I1 = b.search(xq, 50)
    assert (I0 == I1).all() and (D0.view(np.uint32) == D1.view(np.uint32)).all()
