"""faiss index-file codec (clipb200/faiss_io.py), CPU only.

The expected bytes below are assembled field by field from faiss's published serialiser
layout (index_write.cpp: write_index_header, WRITEXBVECTOR, write_ivf_header,
write_direct_map, write_InvertedLists), independently of the module under test.  faiss itself
is not installable here, so this pins the codec against the format description only
("parity unpinned" in DESIGN.md)."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cli-p_b200"))
from clipb200 import faiss_io as fio  # noqa: E402


def hdr(d, ntotal, metric=0, trained=1):
    return struct.pack("<i", d) + struct.pack("<q", ntotal) + struct.pack("<qq", 1 << 20, 1 << 20) + \
        struct.pack("<B", trained) + struct.pack("<i", metric)


def flat_bytes(x, metric=0):
    n, d = x.shape
    return (b"IxFI" if metric == 0 else b"IxF2") + hdr(d, n, metric) + struct.pack("<Q", n * d) + x.astype("<f4").tobytes()


def ivf_bytes(x, assign, nlist, nprobe, sparse=False, direct_map=0):
    """IwFl file with rows x distributed over `nlist` lists by assign[i]; ids = row numbers."""
    n, d = x.shape
    cent = np.arange(nlist * d, dtype=np.float32).reshape(nlist, d)
    out = b"IwFl" + hdr(d, n) + struct.pack("<QQ", nlist, nprobe) + flat_bytes(cent)
    if direct_map == 0:
        out += struct.pack("<BQ", 0, 0)
    else:   # Array map: one packed (list, offset) entry per id - contents are irrelevant to the reader
        out += struct.pack("<BQ", 1, n) + np.zeros(n, dtype="<i8").tobytes()
    out += b"ilar" + struct.pack("<QQ", nlist, d * 4)
    lists = [[i for i in range(n) if assign[i] == li] for li in range(nlist)]
    if sparse:
        pairs = []
        for li, ids in enumerate(lists):
            if ids:
                pairs += [li, len(ids)]
        out += b"sprs" + struct.pack("<Q", len(pairs)) + np.array(pairs, dtype="<u8").tobytes()
    else:
        out += b"full" + struct.pack("<Q", nlist) + np.array([len(l) for l in lists], dtype="<u8").tobytes()
    for ids in lists:
        if ids:
            out += x[ids].astype("<f4").tobytes() + np.array(ids, dtype="<i8").tobytes()
    return out


@pytest.fixture
def rows():
    return np.random.default_rng(5).standard_normal((37, 8)).astype(np.float32)


def test_flat_file_is_byte_exact_and_parses(tmp_path, rows):
    p = str(tmp_path / "flat.index")
    fio.write_flat(p, 8, 37, lambda lo, hi: rows[lo:hi], step=10)
    assert open(p, "rb").read() == flat_bytes(rows)
    ix = fio.parse(p)
    assert (ix.kind, ix.d, ix.ntotal, ix.metric) == ("flat", 8, 37, 0)
    np.testing.assert_array_equal(ix.rows(0, 37), rows)
    np.testing.assert_array_equal(np.concatenate(list(ix.iter_rows(step=7))), rows)
    np.testing.assert_array_equal(ix.rows(5, 9), rows[5:9])


@pytest.mark.parametrize("sparse", [False, True])
@pytest.mark.parametrize("direct_map", [0, 1])
def test_ivf_file_flattens_back_to_add_order(tmp_path, rows, sparse, direct_map):
    rng = np.random.default_rng(2)
    nlist = 5
    assign = rng.integers(0, nlist, size=37)
    assign[assign == 3] = 1                      # list 3 stays empty
    p = str(tmp_path / "ivf.index")
    open(p, "wb").write(ivf_bytes(rows, assign, nlist, 32, sparse=sparse, direct_map=direct_map))
    ix = fio.parse(p)
    assert (ix.kind, ix.d, ix.ntotal, ix.nlist, ix.nprobe) == ("ivf", 8, 37, nlist, 32)
    np.testing.assert_array_equal(ix.rows(0, 37), rows)
    np.testing.assert_array_equal(np.concatenate(list(ix.iter_rows(step=6))), rows)


def test_single_list_ivf_writer_layout(tmp_path, rows):
    p = str(tmp_path / "ivf1.index")
    fio.write_ivf_single_list(p, 8, 37, lambda lo, hi: rows[lo:hi], nprobe=32, step=9)
    b = open(p, "rb").read()
    # the same file assembled by hand: nlist 1, centroid = mean row, "full" table, codes then ids
    cent = (rows.astype(np.float64).sum(axis=0) / 37).astype("<f4")[None, :]
    want = b"IwFl" + hdr(8, 37) + struct.pack("<QQ", 1, 32) + flat_bytes(cent) + struct.pack("<BQ", 0, 0) + \
        b"ilar" + struct.pack("<QQ", 1, 32) + b"full" + struct.pack("<QQ", 1, 37) + \
        rows.astype("<f4").tobytes() + np.arange(37, dtype="<i8").tobytes()
    assert b == want
    ix = fio.parse(p)
    assert (ix.kind, ix.nlist, ix.nprobe, ix.ntotal) == ("ivf", 1, 32, 37)
    np.testing.assert_array_equal(ix.rows(0, 37), rows)


def test_empty_indexes(tmp_path):
    p = str(tmp_path / "e.index")
    fio.write_flat(p, 512, 0, lambda lo, hi: np.zeros((0, 512), np.float32))
    ix = fio.parse(p)
    assert ix.ntotal == 0 and list(ix.iter_rows()) == []
    fio.write_ivf_single_list(p, 512, 0, lambda lo, hi: np.zeros((0, 512), np.float32))
    ix = fio.parse(p)
    assert ix.kind == "ivf" and ix.ntotal == 0 and ix.nlist == 1


def test_malformed_files_raise(tmp_path, rows):
    p = str(tmp_path / "bad.index")
    good = flat_bytes(rows)
    open(p, "wb").write(good[:-5])
    with pytest.raises(fio.FaissFormatError, match="truncated"):
        fio.parse(p)
    open(p, "wb").write(b"IxPQ" + good[4:])
    with pytest.raises(fio.FaissFormatError, match="not supported"):
        fio.parse(p)
    open(p, "wb").write(good[:4] + hdr(8, 36) + good[4 + len(hdr(8, 36)):])
    with pytest.raises(fio.FaissFormatError, match="header says"):
        fio.parse(p)
    # ids that are not 0..n-1 (add_with_ids) cannot be mapped onto add order
    bad = bytearray(ivf_bytes(rows, np.zeros(37, dtype=int), 1, 1))
    bad[-8:] = struct.pack("<q", 1000)
    open(p, "wb").write(bytes(bad))
    with pytest.raises(fio.FaissFormatError, match="ids outside"):
        fio.parse(p)
    dup = bytearray(ivf_bytes(rows, np.zeros(37, dtype=int), 1, 1))
    dup[-8:] = struct.pack("<q", 0)
    open(p, "wb").write(bytes(dup))
    with pytest.raises(fio.FaissFormatError, match="do not cover"):
        fio.parse(p)


def test_l2_metric_header_and_metric_arg(tmp_path, rows):
    p = str(tmp_path / "l2.index")
    open(p, "wb").write(flat_bytes(rows, metric=1))
    assert fio.parse(p).metric == 1
    # metric_type > 1 carries a float metric_arg after the header
    b = b"IxFl" + hdr(8, 37, metric=4) + struct.pack("<f", 3.0) + struct.pack("<Q", 37 * 8) + rows.tobytes()
    open(p, "wb").write(b)
    ix = fio.parse(p)
    assert ix.metric == 4
    np.testing.assert_array_equal(ix.rows(0, 37), rows)


def test_random_ivf_partitions_round_trip(tmp_path):
    """Property: whatever the k-means partition (empty lists, one list, every row its own list, full or
    sparse size table), the flattened rows come back in add order."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(n=st.integers(0, 120), d=st.sampled_from([4, 16]), nlist=st.integers(1, 14), sparse=st.booleans(),
           seed=st.integers(0, 2 ** 16))
    def check(n, d, nlist, sparse, seed):
        rng = np.random.default_rng(seed)
        x = rng.standard_normal((n, d)).astype(np.float32)
        assign = rng.integers(0, max(1, nlist // (1 + seed % 3)), size=n)      # leaves some lists empty
        p = str(tmp_path / "h.index")
        open(p, "wb").write(ivf_bytes(x, assign, nlist, 1 + seed % 100, sparse=sparse))
        ix = fio.parse(p)
        assert (ix.kind, ix.d, ix.ntotal, ix.nlist) == ("ivf", d, n, nlist)
        got = np.concatenate(list(ix.iter_rows(step=17))) if n else np.zeros((0, d), np.float32)
        np.testing.assert_array_equal(got, x)

    check()
