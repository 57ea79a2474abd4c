"""Property tests (hypothesis) of the top-k semantics: oracle self-consistency on the CPU,
and the CUDA path against the oracle on adversarially structured inputs on the GPU."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import flatip_ref as F


@st.composite
def score_rows(draw):
    n = draw(st.integers(1, 300))
    nq = draw(st.integers(1, 3))
    # few distinct values -> many exact ties
    vals = draw(st.lists(st.floats(-2, 2, allow_nan=False, width=32), min_size=1, max_size=6))
    idx = draw(st.lists(st.integers(0, len(vals) - 1), min_size=n * nq, max_size=n * nq))
    S = np.array([vals[i] for i in idx], dtype=np.float32).reshape(nq, n)
    k = draw(st.integers(1, n + 5))
    return S, k


@given(score_rows())
@settings(max_examples=150, deadline=None)
def test_oracle_topk_is_the_sorted_prefix(case):
    S, k = case
    D, I = F.topk_from_scores(S, k)
    nq, n = S.shape
    for q in range(nq):
        order = sorted(range(n), key=lambda i: (-float(S[q, i] + np.float32(0.0)), i))[:k]
        assert list(I[q, :len(order)]) == order
        assert (I[q, len(order):] == -1).all() and (D[q, len(order):] == F.NEG_FLT_MAX).all()
        assert (np.diff(D[q, :len(order)]) <= 0).all()


@given(score_rows(), st.integers(2, 5))
@settings(max_examples=60, deadline=None)
def test_oracle_sharded_merge_equals_unsharded(case, R):
    S, k = case
    nq, n = S.shape
    D, I = F.topk_from_scores(S, k)
    per = -(-n // R)
    Ds, Is = [], []
    for r in range(R):
        lo, hi = min(r * per, n), min((r + 1) * per, n)
        if hi > lo:
            d, i = F.topk_from_scores(S[:, lo:hi], k, id_base=lo)
        else:
            d = np.full((nq, k), F.NEG_FLT_MAX, np.float32)
            i = np.full((nq, k), -1, np.int64)
        Ds.append(d)
        Is.append(i)
    Dm, Im = F.merge_topk(np.stack(Ds), np.stack(Is), k)
    assert (Im == I).all() and (Dm == D).all()


@pytest.mark.gpu
@given(st.integers(1, 4000), st.integers(1, 4), st.integers(1, 64), st.integers(0, 2 ** 31 - 1),
       st.sampled_from(["f16", "f32"]), st.integers(1, 12))
@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
def test_cuda_search_on_tie_heavy_databases(n, nq, k, seed, storage, distinct):
    """Databases made of a handful of distinct rows: almost every comparison is an exact tie,
    so the result is decided by the id tie-break at every rank."""
    from clipb200 import faiss, synth
    rng = np.random.default_rng(seed)
    proto = synth.unit_rows(distinct, seed=seed % 1000).astype(np.float16).astype(np.float32)
    xb = proto[rng.integers(0, distinct, n)]
    xq = proto[rng.integers(0, distinct, nq)] + 0.0
    index = faiss.IndexFlatIP(512, storage=storage, devices=[0])
    index.add(xb)
    D, I = index.search(xq, k)
    Dref, Iref = F.search(xq, xb.astype(np.float16) if storage == "f16" else xb, k)
    # identical rows give bit-identical scores on both sides, so ids must match exactly wherever the
    # oracle's own scores separate neighbours by >= 1e-5, and tie groups must be the k lowest ids
    ok, _, msg = F.ids_match_with_tolerance(Dref, Iref, D, I)
    assert ok, msg
    valid = Iref >= 0
    np.testing.assert_allclose(D[valid], Dref[valid], atol=1e-5, rtol=0)
    for q in range(nq):
        ids = I[q][I[q] >= 0]
        assert len(set(ids.tolist())) == len(ids)
