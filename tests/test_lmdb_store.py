"""CPU: the LMDB-format store.  No liblmdb exists in this image, so the file
is checked by this repo's own independent reader plus structural invariants of the format
(page accounting, sorted keys, overflow placement) -- parity unpinned."""
import os
import struct

import numpy as np
import pytest

from clipb200 import lmdb


def _fill(env, n, seed=0):
    rng = np.random.default_rng(seed)
    fn_db = env.open_db(b"fn_db")
    skip_db = env.open_db(b"skip_db")
    want = {}
    for i in range(n):
        key = f"/photos/album_{i % 7}/img_{i:06d}.jpg".encode()
        vec = rng.standard_normal(512).astype(np.float32)
        with env.begin(db=fn_db, write=True) as txn:            # build-index.py:42-51
            if txn.get(key) is not None:
                continue
            txn.put(key, vec.tobytes())
        want[key] = vec.tobytes()
    return fn_db, skip_db, want


def test_round_trip_in_reference_usage_pattern(tmp_path):
    path = str(tmp_path / "vectors.lmdb")
    env = lmdb.open(path, map_size=1024 * 1024 * 1024 * 20, max_dbs=4)
    fn_db, skip_db, want = _fill(env, 300)
    idx_db = env.open_db(b"idx_db")
    with env.begin(db=fn_db) as txn:                              # build-index.py:68-89
        assert txn.stat()["entries"] == 300
        cur = txn.cursor()
        assert cur.first()
        seen = []
        for i, (k, v) in enumerate(cur):
            seen.append(k)
            assert np.frombuffer(v, dtype=np.float32).shape == (512,)
            with env.begin(db=idx_db, write=True) as itxn:
                itxn.put(f"{i}".encode(), k, dupdata=False, overwrite=True)
        assert seen == sorted(want)                               # memcmp key order defines the ids
    with env.begin(db=skip_db) as txn:
        assert txn.get(b"/nope") is None
    env.close()
    assert os.path.exists(os.path.join(path, "data.mdb")) and os.path.exists(os.path.join(path, "lock.mdb"))

    env2 = lmdb.open(path, map_size=1024 * 1024 * 1024 * 20, max_dbs=4)  # query-index.py:25-27
    idx2, fn2 = env2.open_db(b"idx_db"), env2.open_db(b"fn_db")
    with env2.begin(db=idx2) as txn:
        key = txn.get(b"17")
    assert key == sorted(want)[17]
    with env2.begin(db=fn2) as txn:
        assert txn.get(key) == want[key]
        assert np.array_equal(np.frombuffer(txn.get(key), dtype=np.float32).reshape((1, 512)),
                              np.frombuffer(want[key], dtype=np.float32).reshape((1, 512)))
    env2.close()


def test_file_structure_matches_lmdb_format(tmp_path):
    path = str(tmp_path / "v.lmdb")
    env = lmdb.open(path, map_size=1 << 30, max_dbs=4)
    _, _, want = _fill(env, 2500, seed=1)
    env.close()
    raw = open(os.path.join(path, "data.mdb"), "rb").read()
    assert len(raw) % 4096 == 0
    # both meta pages: P_META, magic, version 1, page size 4096 in the free DB's md_pad
    txnids = []
    for pg in (0, 1):
        p = raw[pg * 4096:(pg + 1) * 4096]
        pgno, _pad, flags = struct.unpack_from("<QHH", p, 0)
        magic, version = struct.unpack_from("<II", p, 16)
        assert pgno == pg and flags == 0x08 and magic == 0xBEEFC0DE and version == 1
        assert struct.unpack_from("<I", p, 16 + 24)[0] == 4096
        last_pg, txnid = struct.unpack_from("<QQ", p, 16 + 24 + 96)
        assert last_pg == len(raw) // 4096 - 1
        txnids.append(txnid)
    assert abs(txnids[0] - txnids[1]) == 1
    # page accounting: every page is a meta, branch, leaf or overflow page claimed by exactly one DB
    dbs, meta = lmdb.read_file(os.path.join(path, "data.mdb"))
    assert dbs[b"fn_db"] == want and dbs[b"skip_db"] == {}
    buf = memoryview(raw)
    total = 2
    main = meta["main"]
    total += main[3] + main[4] + main[5]
    for key, val, fl in lmdb._walk(buf, main[7]):
        assert fl & lmdb.F_SUBDATA
        sub = struct.unpack(lmdb.DB_FMT, val)
        total += sub[3] + sub[4] + sub[5]
        if key == b"fn_db":
            assert sub[6] == 2500 and sub[5] == 2500          # one overflow page per 2048-byte value
            assert sub[2] >= 2                                   # needs branch pages
    assert total == len(raw) // 4096
    # every leaf/branch page: sorted keys, lower/upper sane, >= 2 keys on branch pages
    for pg in range(2, len(raw) // 4096):
        p = raw[pg * 4096:(pg + 1) * 4096]
        flags, lower, upper = struct.unpack_from("<HHH", p, 10)
        if flags & 0x04:
            continue
        if not flags & 0x03:
            continue                                             # continuation of an overflow run
        n = (lower - 16) // 2
        assert 16 <= lower <= upper <= 4096
        ptrs = struct.unpack_from(f"<{n}H", p, 16)
        keys = []
        for off in ptrs:
            ks = struct.unpack_from("<H", p, off + 6)[0]
            keys.append(bytes(p[off + 8:off + 8 + ks]))
        if flags & 0x01:
            assert n >= 2 and keys[0] == b""
            keys = keys[1:]
        assert keys == sorted(keys)


def test_large_values_small_values_and_limits(tmp_path):
    env = lmdb.open(str(tmp_path / "x.lmdb"), map_size=1 << 30, max_dbs=2)
    db = env.open_db(b"d")
    with env.begin(db=db, write=True) as txn:
        txn.put(b"small", b"1")                                   # skip_db style value (build-index.py:61)
        txn.put(b"big", bytes(range(256)) * 40)                   # 10240 bytes -> 3 overflow pages
        txn.put(b"k" * 511, b"v")
        with pytest.raises(lmdb.BadValsizeError):
            txn.put(b"k" * 512, b"v")                              # liblmdb: MDB_BAD_VALSIZE
        assert txn.put(b"small", b"2", overwrite=False) is False
    env.close()
    dbs, _ = lmdb.read_file(str(tmp_path / "x.lmdb" / "data.mdb"))
    assert dbs[b"d"][b"small"] == b"1" and dbs[b"d"][b"big"] == bytes(range(256)) * 40
    assert dbs[b"d"][b"k" * 511] == b"v"


def test_aborted_transaction_leaves_no_trace(tmp_path):
    env = lmdb.open(str(tmp_path / "a.lmdb"), map_size=1 << 30, max_dbs=2)
    db = env.open_db(b"d")
    try:
        with env.begin(db=db, write=True) as txn:
            txn.put(b"a", b"1")
            raise KeyError("boom")
    except KeyError:
        pass
    with env.begin(db=db) as txn:
        assert txn.get(b"a") is None
    with env.begin(db=db, write=True) as txn:
        txn.put(b"a", b"2")
        assert txn.get(b"a") == b"2"
        assert txn.delete(b"a") and txn.get(b"a") is None
    env.close()


def test_three_level_tree(tmp_path):
    path = str(tmp_path / "big.lmdb")
    env = lmdb.open(path, map_size=1 << 32, max_dbs=2, sync_every=10 ** 9)
    db = env.open_db(b"idx_db")
    with env.begin(db=db, write=True) as txn:
        for i in range(60_000):
            txn.put(f"{i}".encode(), f"/some/fairly/long/path/prefix/for/the/photo/library/img_{i:08d}.jpeg".encode())
    env.close()
    dbs, meta = lmdb.read_file(os.path.join(path, "data.mdb"))
    assert len(dbs[b"idx_db"]) == 60_000
    assert dbs[b"idx_db"][b"59999"].endswith(b"img_00059999.jpeg")
    for key, val, fl in lmdb._walk(memoryview(open(os.path.join(path, "data.mdb"), "rb").read()), meta["main"][7]):
        if key == b"idx_db":
            assert struct.unpack(lmdb.DB_FMT, val)[2] == 3       # depth
