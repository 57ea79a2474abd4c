"""CPU: the oracle restatements agree with each other and with the committed fixtures."""
import os

import numpy as np
import pytest

from oracle import flatip_ref as F
from clipb200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "flatip_golden.npz")


def _golden_inputs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(GOLDEN), "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.flatip_inputs()


def test_numpy_oracle_matches_golden():
    xb, xq = _golden_inputs()
    g = np.load(GOLDEN)
    for k in (1, 21, 100):
        D, I = F.search(xq, xb, k)
        ok, _, msg = F.ids_match_with_tolerance(g[f"D{k}"], g[f"I{k}"], D, I)
        assert ok, msg
        np.testing.assert_allclose(D, g[f"D{k}"], atol=1e-6)


def test_c_oracle_matches_golden(oracle_c):
    xb, xq = _golden_inputs()
    g = np.load(GOLDEN)
    for k in (1, 21, 100):
        D, I = oracle_c(xb, xq, k)
        ok, _, msg = F.ids_match_with_tolerance(g[f"D{k}"], g[f"I{k}"], D, I)
        assert ok, msg
        np.testing.assert_allclose(D, g[f"D{k}"], atol=1e-6)


def test_golden_has_exact_ties_in_id_order():
    g = np.load(GOLDEN)
    I = g["I21"][5]          # query 5 == row 7, duplicated at rows 100..107
    assert list(I[:9]) == [7, 100, 101, 102, 103, 104, 105, 106, 107]


@pytest.mark.parametrize("dtype", [np.float16, np.float32])
@pytest.mark.parametrize("k", [1, 5, 64])
def test_numpy_vs_c(oracle_c, dtype, k):
    xb = synth.unit_rows(3000, seed=3).astype(dtype)
    xq = synth.unit_rows(4, seed=4)
    D, I = F.search(xq, xb, k)
    D2, I2 = oracle_c(xb, xq, k)
    ok, _, msg = F.ids_match_with_tolerance(D, I, D2, I2)
    assert ok, msg
    np.testing.assert_allclose(D, D2, atol=1e-6)


def test_padding_when_k_exceeds_n(oracle_c):
    xb = synth.unit_rows(5, seed=1)
    xq = synth.unit_rows(2, seed=2)
    for fn in (lambda: F.search(xq, xb, 8), lambda: oracle_c(xb, xq, 8)):
        D, I = fn()
        assert (I[:, 5:] == -1).all() and (D[:, 5:] == F.NEG_FLT_MAX).all()
        assert sorted(I[0, :5]) == [0, 1, 2, 3, 4]
        assert (np.diff(D[:, :5], axis=1) <= 0).all()


def test_all_equal_scores_take_lowest_ids(oracle_c):
    xb = np.zeros((300, 512), np.float32)
    xq = synth.unit_rows(1, seed=2)
    for D, I in (F.search(xq, xb, 10), oracle_c(xb, xq, 10)):
        assert list(I[0]) == list(range(10))
        assert (D == 0).all()


def test_blocked_search_equals_direct():
    xb = synth.unit_rows(5000, seed=5).astype(np.float16)
    xq = synth.unit_rows(3, seed=6)
    D, I = F.search(xq, xb, 50)
    Db, Ib = F.search(xq, xb, 50, block=777)
    # BLAS may sum in a different order for a different block shape: ulp-level only
    ok, _, msg = F.ids_match_with_tolerance(D, I, Db, Ib)
    assert ok, msg
    np.testing.assert_allclose(D, Db, atol=1e-6)


def test_merge_topk_equals_unsharded():
    xb = synth.unit_rows(4001, seed=8).astype(np.float16)
    xq = synth.unit_rows(3, seed=9)
    k = 40
    D, I = F.search(xq, xb, k)
    Ds, Is = [], []
    R = 3
    per = -(-len(xb) // R)
    for r in range(R):
        d, i = F.search(xq, xb[r * per:(r + 1) * per], k)
        i = np.where(i >= 0, i + r * per, i)
        Ds.append(d)
        Is.append(i)
    Dm, Im = F.merge_topk(np.stack(Ds), np.stack(Is), k)
    ok, _, msg = F.ids_match_with_tolerance(D, I, Dm, Im)
    assert ok, msg
    np.testing.assert_allclose(D, Dm, atol=1e-6)


def test_tolerance_rule():
    D = np.array([[0.9, 0.5, 0.499995, 0.1]], np.float32)
    I = np.array([[1, 2, 3, 4]])
    assert F.ids_match_with_tolerance(D, I, D, np.array([[1, 3, 2, 4]]))[0]
    assert not F.ids_match_with_tolerance(D, I, D, np.array([[2, 1, 3, 4]]))[0]


def test_numpy_oracle_matches_scikit_learn_brute_force():
    """A third, independent exact k-NN: scikit-learn's brute-force neighbours under the cosine metric.
    On unit-norm rows cosine distance = 1 - <q, x>, so its neighbour order is the inner-product order."""
    from sklearn.neighbors import NearestNeighbors
    xb = synth.unit_rows(20_000, seed=11, clip_like=True)
    xq = synth.unit_rows(8, seed=12, clip_like=True)
    for k in (1, 21, 100):
        D, I = F.search(xq, xb, k)
        nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(xb)
        dist, idx = nn.kneighbors(xq)
        ok, _, msg = F.ids_match_with_tolerance(D, I, (1.0 - dist).astype(np.float32), idx.astype(np.int64))
        assert ok, msg
        np.testing.assert_allclose(D, 1.0 - dist, atol=2e-6)
