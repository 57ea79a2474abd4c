"""ORACLE (test infrastructure, not product code) -- exact inner-product top-k.

CPU restatement of what `faiss.IndexFlatIP(512).search(x, k)` computes, i.e. the
call the reference makes at /root/reference/query-index.py:111 (`D, I =
index.search(features, k + offset + 1)`) on the index it builds at
/root/reference/build-index.py:80-81,99,107.

The arithmetic lives in a third-party dependency that is NOT vendored in the
reference (setup.sh:12 clones facebookresearch/faiss at HEAD, un-pinned, CPU
build).  Published semantics of faiss `IndexFlat::search` ->
`knn_inner_product` restated here:

  * scores are fp32 inner products  <q, x_i>
  * the k largest per query, sorted descending
  * ids are positions in add() order (int64)
  * unfilled slots (k > ntotal): I = -1, D = -FLT_MAX (-3.4028235e38)
  * ties: the earlier id is kept (strict `>` is needed to displace)

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
this path and faiss is not installable offline, so this
oracle is pinned only against an independent plain-C restatement of faiss's
heap algorithm (oracle/flatip_ref.c) and committed fixtures made by
tests/golden/make_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm
may import this module.
"""
from __future__ import annotations

import numpy as np

NEG_FLT_MAX = np.float32(-3.4028234663852886e38)


def scores_f32(xq: np.ndarray, xb: np.ndarray) -> np.ndarray:
    """fp32 score matrix (nq, n).  `xb` may be fp16 storage: it is up-cast to
    fp32 first, so both arms see the identical fp16-rounded database values
    (hard part 2)."""
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    xb32 = np.ascontiguousarray(xb).astype(np.float32, copy=False)
    return xq @ xb32.T


def topk_from_scores(S: np.ndarray, k: int, id_base: int = 0):
    """k largest per row of S, ordered by (-score, id); padded like faiss."""
    nq, n = S.shape
    D = np.full((nq, k), NEG_FLT_MAX, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    kk = min(k, n)
    if kk == 0:
        return D, I
    ids = np.arange(n, dtype=np.int64)
    for qi in range(nq):
        s = S[qi] + np.float32(0.0)          # -0.0 -> +0.0 so zero ties order by id
        if kk < n:
            # candidate superset: everything >= the kk-th largest value
            thr = np.partition(s, n - kk)[n - kk]
            cand = ids[s >= thr]
        else:
            cand = ids
        order = np.lexsort((cand, -s[cand].astype(np.float64)))[:kk]
        sel = cand[order]
        D[qi, :kk] = s[sel]
        I[qi, :kk] = sel + id_base
    return D, I


def search(xq: np.ndarray, xb: np.ndarray, k: int, block: int = 1 << 18):
    """IndexFlatIP.search restatement; database streamed in row blocks so a
    10M-row check does not need an (nq, 10M) matrix."""
    xq = np.ascontiguousarray(xq, dtype=np.float32)
    assert xq.ndim == 2 and xb.ndim == 2 and xq.shape[1] == xb.shape[1]
    assert k > 0
    n = xb.shape[0]
    if n <= block:
        return topk_from_scores(scores_f32(xq, xb), k)
    Ds, Is = [], []
    for lo in range(0, n, block):
        d, i = topk_from_scores(scores_f32(xq, xb[lo:lo + block]), k, id_base=lo)
        Ds.append(d)
        Is.append(i)
    return merge_topk(np.stack(Ds), np.stack(Is), k)


def merge_topk(Ds: np.ndarray, Is: np.ndarray, k: int):
    """Merge R per-shard results (R, nq, k) by (-score, id); -1 ids sort last.
    This is the cross-GPU merge the sharded search performs."""
    R, nq, kk = Ds.shape
    D = np.full((nq, k), NEG_FLT_MAX, dtype=np.float32)
    I = np.full((nq, k), -1, dtype=np.int64)
    for qi in range(nq):
        d = Ds[:, qi, :].reshape(-1)
        i = Is[:, qi, :].reshape(-1)
        valid = i >= 0
        d, i = d[valid], i[valid]
        order = np.lexsort((i, -d.astype(np.float64)))[:k]
        D[qi, :len(order)] = d[order]
        I[qi, :len(order)] = i[order]
    return D, I


def ids_match_with_tolerance(D_ref, I_ref, D_got, I_got, gap: float = 1e-5):
    """North-star parity rule: ids identical except where adjacent reference
    scores differ by < gap.  Returns (ok, n_exempt_positions, message)."""
    D_ref = np.asarray(D_ref); I_ref = np.asarray(I_ref)
    D_got = np.asarray(D_got); I_got = np.asarray(I_got)
    if I_ref.shape != I_got.shape:
        return False, 0, f"shape {I_got.shape} != {I_ref.shape}"
    exempt = 0
    nq, k = I_ref.shape
    for qi in range(nq):
        bad = np.nonzero(I_ref[qi] != I_got[qi])[0]
        for j in bad:
            # position j may differ only if some neighbouring reference score is
            # within `gap` of it (a near-tie whose order fp32 summation may flip),
            # or the k-th boundary is a near-tie with the first excluded element.
            s = D_ref[qi, j]
            near = False
            if j > 0 and abs(D_ref[qi, j - 1] - s) < gap:
                near = True
            if j + 1 < k and abs(D_ref[qi, j + 1] - s) < gap:
                near = True
            if j == k - 1 and abs(D_got[qi, j] - s) < gap:
                near = True
            if not near:
                return False, exempt, (
                    f"query {qi} rank {j}: id {I_got[qi, j]} (score {D_got[qi, j]!r}) "
                    f"!= ref id {I_ref[qi, j]} (score {s!r}), no near-tie")
            exempt += 1
    return True, exempt, "ok"


def recall_at_k(I_ref, I_got) -> float:
    hit = tot = 0
    for a, b in zip(np.asarray(I_ref), np.asarray(I_got)):
        a = a[a >= 0]
        hit += len(np.intersect1d(a, b[b >= 0]))
        tot += len(a)
    return hit / max(tot, 1)
