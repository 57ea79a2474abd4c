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


def _check_pages(raw):
    """Structural invariants of every reachable branch / leaf page of a data.mdb image."""
    seen, _ = _reachable(raw)
    for pg in sorted(p for p, owner in seen.items() if not isinstance(owner, tuple)):
        p = raw[pg * 4096:(pg + 1) * 4096]
        flags, lower, upper = struct.unpack_from("<HHH", p, 10)
        assert flags & 0x03 and not flags & 0x04
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


def _reachable(raw):
    """Pages reachable from the newest meta page, as {pgno: owner}; asserts no page is claimed twice."""
    buf = memoryview(raw)
    meta = lmdb._read_meta(buf)
    seen = {}

    def tree(root, owner):
        if root == lmdb.P_INVALID:
            return
        stack = [root]
        while stack:
            pg = stack.pop()
            assert pg not in seen, f"page {pg} claimed by {seen.get(pg)} and {owner}"
            seen[pg] = owner
            p = buf[pg * 4096:(pg + 1) * 4096]
            flags, lower = struct.unpack_from("<HH", p, 10)
            n = (lower - 16) // 2
            for off in struct.unpack_from(f"<{n}H", p, 16):
                lo, hi, fl, ks = struct.unpack_from("<HHHH", p, off)
                if flags & 0x01:
                    stack.append(lo | (hi << 16) | (fl << 32))
                elif fl & lmdb.F_BIGDATA:
                    ov = struct.unpack_from("<Q", p, off + 8 + ks)[0]
                    npg = (16 - 1 + (lo | (hi << 16))) // 4096 + 1
                    for j in range(npg):
                        assert ov + j not in seen
                        seen[ov + j] = (owner, "overflow")

    tree(meta["main"][7], b"")
    for key, val, fl in lmdb._walk(buf, meta["main"][7]):
        if fl & lmdb.F_SUBDATA:
            tree(struct.unpack(lmdb.DB_FMT, val)[7], key)
    return seen, meta


def test_file_structure_matches_lmdb_format(tmp_path):
    path = str(tmp_path / "v.lmdb")
    env = lmdb.open(path, map_size=1 << 30, max_dbs=4, flush_records=300)
    _, _, want = _fill(env, 2500, seed=1)
    env.sync()
    garbage = env.info()["garbage_pages"]
    assert env.info()["pending_records"] == 0
    env.close()
    raw = open(os.path.join(path, "data.mdb"), "rb").read()
    assert len(raw) % 4096 == 0
    # both meta pages: P_META, magic, version 1, page size 4096 in the free DB's md_pad; the newer one
    # (larger txnid, in page txnid & 1) describes the whole file, the older one the previous commit
    metas = []
    for pg in (0, 1):
        p = raw[pg * 4096:(pg + 1) * 4096]
        pgno, _pad, flags = struct.unpack_from("<QHH", p, 0)
        magic, version = struct.unpack_from("<II", p, 16)
        assert pgno == pg and flags == 0x08 and magic == 0xBEEFC0DE and version == 1
        assert struct.unpack_from("<I", p, 16 + 24)[0] == 4096
        metas.append(struct.unpack_from("<QQ", p, 16 + 24 + 96))       # (last_pg, txnid)
    new, old = sorted(metas, key=lambda m: -m[1])
    assert new[1] == old[1] + 1 and metas[new[1] & 1] == new
    assert new[0] == len(raw) // 4096 - 1 and old[0] <= new[0]
    dbs, meta = lmdb.read_file(os.path.join(path, "data.mdb"))
    assert dbs[b"fn_db"] == want and dbs[b"skip_db"] == {}
    _check_pages(raw)
    # copy-on-write: every page is either reachable from the newest meta (claimed once) or was freed by a
    # later commit; the MDB_db page counters match what is reachable
    seen, _ = _reachable(raw)
    assert len(seen) + 2 + garbage == len(raw) // 4096
    counts = {}
    for pg, owner in seen.items():
        owner = owner[0] if isinstance(owner, tuple) else owner
        counts[owner] = counts.get(owner, 0) + 1
    for key, val, fl in lmdb._walk(memoryview(raw), meta["main"][7]):
        sub = struct.unpack(lmdb.DB_FMT, val)
        assert sub[3] + sub[4] + sub[5] == counts.get(key, 0)
        if key == b"fn_db":
            assert sub[6] == 2500 and sub[5] == 2500          # one overflow page per 2048-byte value
            assert sub[2] >= 2                                   # needs branch pages
    assert meta["main"][3] + meta["main"][4] == counts[b""]
    # a compacted copy wastes nothing
    env = lmdb.open(path, map_size=1 << 30, max_dbs=4)
    env.copy(str(tmp_path / "packed.lmdb"), compact=True)
    env.close()
    raw2 = open(str(tmp_path / "packed.lmdb" / "data.mdb"), "rb").read()
    seen2, _ = _reachable(raw2)
    assert len(seen2) + 2 == len(raw2) // 4096 < len(raw) // 4096
    assert lmdb.read_file(str(tmp_path / "packed.lmdb" / "data.mdb"))[0][b"fn_db"] == want
    _check_pages(raw2)


def test_commits_survive_a_crash_and_torn_flushes_do_not(tmp_path):
    """Copy-on-write: data pages are written and fsync'ed before the meta page flips.  Cutting the file
    anywhere inside the last flush (data written, meta not yet) must leave the previous commit intact."""
    path = str(tmp_path / "c.lmdb")
    env = lmdb.open(path, map_size=1 << 30, max_dbs=2, sync=True)          # sync=True: liblmdb's durability
    db = env.open_db(b"fn_db")
    vals = {f"/p/{i:04d}.jpg".encode(): bytes([i % 251]) * 2048 for i in range(40)}
    for k, v in list(vals.items())[:30]:
        with env.begin(db=db, write=True) as txn:
            txn.put(k, v)
    data = os.path.join(path, "data.mdb")
    before = open(data, "rb").read()
    with env.begin(db=db, write=True) as txn:
        for k, v in list(vals.items())[30:]:
            txn.put(k, v)
    after = open(data, "rb").read()
    env.close()
    assert lmdb.read_file(data)[0][b"fn_db"] == vals
    # a crash after the data pages but before the meta flip = the new pages appended to the old file
    torn = bytearray(after)
    torn[:8192] = before[:8192]
    open(data, "wb").write(torn)
    assert lmdb.read_file(data)[0][b"fn_db"] == dict(list(vals.items())[:30])
    env = lmdb.open(path, map_size=1 << 30, max_dbs=2)                     # and the store opens and extends from there
    db = env.open_db(b"fn_db")
    with env.begin(db=db, write=True) as txn:
        txn.put(b"/p/zzzz.jpg", b"x")
    env.close()
    got = lmdb.read_file(data)[0][b"fn_db"]
    assert len(got) == 31 and got[b"/p/zzzz.jpg"] == b"x"


def test_second_writer_is_refused_readers_are_not(tmp_path):
    path = str(tmp_path / "w.lmdb")
    env = lmdb.open(path, map_size=1 << 30, max_dbs=2)
    db = env.open_db(b"d")
    with env.begin(db=db, write=True) as txn:
        txn.put(b"a", b"1")
    env.sync()
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # a second process: opening read-write and READING is fine (query-index.py beside build-index.py);
    # its first WRITE is refused while this process holds the writer lock
    code = ("import sys; sys.path.insert(0, %r); from clipb200 import lmdb\n"
            "e = lmdb.open(%r, map_size=1 << 30, max_dbs=2)\n"
            "d = e.open_db(b'd')\n"
            "with e.begin(db=d) as t: print(t.get(b'a'))\n"
            "try:\n"
            "    with e.begin(db=d, write=True) as t: t.put(b'b', b'2')\n"
            "    e.sync(); print('wrote')\n"
            "except lmdb.Error as ex:\n    print('refused')\n") % (os.path.join(root, "cli-p_b200"), path)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.stdout.split() == ["b'1'", "refused"], out.stdout + out.stderr
    # this (reading) side sees what the writer commits later, at its next begin()
    reader = lmdb.open(path, map_size=1 << 30, max_dbs=2)
    rdb = reader.open_db(b"d")
    with env.begin(db=db, write=True) as txn:
        txn.put(b"c", b"3")
    env.sync()
    with reader.begin(db=rdb) as txn:
        assert txn.get(b"c") == b"3"
    reader.close()
    env.close()
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.stdout.split() == ["b'1'", "wrote"], out.stdout + out.stderr
    assert lmdb.read_file(os.path.join(path, "data.mdb"))[0][b"d"] == {b"a": b"1", b"b": b"2", b"c": b"3"}


def test_random_operations_against_a_dict_model(tmp_path):
    """Puts, overwrites, deletes, reopen, cursors from arbitrary keys: the tree always equals a dict."""
    rng = np.random.default_rng(7)
    path = str(tmp_path / "r.lmdb")
    model = {}
    env = lmdb.open(path, map_size=1 << 30, max_dbs=2, flush_records=97)
    db = env.open_db(b"d")
    for round_ in range(6):
        for _ in range(1500):
            k = f"k{rng.integers(0, 3000):05d}".encode() * int(rng.integers(1, 4))
            op = rng.integers(0, 10)
            with env.begin(db=db, write=True) as txn:
                if op < 6:
                    v = bytes(rng.integers(0, 256, size=int(rng.choice([1, 40, 700, 2048, 5000])), dtype=np.uint8))
                    txn.put(k, v)
                    model[k] = v
                elif op < 9:
                    assert txn.delete(k) == (k in model)
                    model.pop(k, None)
                else:
                    assert txn.get(k) == model.get(k)
        with env.begin(db=db) as txn:
            assert txn.stat()["entries"] == len(model)
            assert list(txn.cursor()) == sorted(model.items())
            probe = f"k{rng.integers(0, 3000):05d}".encode()
            cur = txn.cursor()
            exp = [kv for kv in sorted(model.items()) if kv[0] >= probe]
            assert cur.set_range(probe) == bool(exp)
            if exp:
                assert cur.item() == exp[0]
        if round_ % 2:
            env.close()
            env = lmdb.open(path, map_size=1 << 30, max_dbs=2, flush_records=97)
            db = env.open_db(b"d")
    env.close()
    raw = open(os.path.join(path, "data.mdb"), "rb").read()
    assert lmdb.read_file(os.path.join(path, "data.mdb"))[0][b"d"] == model
    _reachable(raw)                                              # no page claimed twice
    _check_pages(raw)


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
    env = lmdb.open(path, map_size=1 << 32, max_dbs=2, flush_records=20_000)
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


def test_store_carries_the_identity_of_the_weights_that_filled_it(tmp_path):
    """A fresh vectors.lmdb is stamped with the checkpoint's fingerprint; another checkpoint is refused
    (resume-by-key would never recompute the old rows); a store that already has rows but no stamp
    (the reference's, or an older clipb200's) is left alone."""
    import types

    import pytest
    from clipb200 import indexer, lmdb
    env = lmdb.open(str(tmp_path / "vectors.lmdb"), map_size=1 << 30, max_dbs=4)
    a, b = types.SimpleNamespace(weights_id="aaaaaaaaaaaaaaaa"), types.SimpleNamespace(weights_id="bbbbbbbbbbbbbbbb")
    indexer.check_weights_stamp(env, a)
    with env.begin(db=env.open_db(indexer.META_DB)) as txn:
        assert bytes(txn.get(b"weights")) == b"aaaaaaaaaaaaaaaa"
    indexer.check_weights_stamp(env, [a, a])                  # same weights, list of per-GPU models
    with pytest.raises(RuntimeError, match="fresh vectors.lmdb"):
        indexer.check_weights_stamp(env, b)
    env.close()

    env = lmdb.open(str(tmp_path / "old.lmdb"), map_size=1 << 30, max_dbs=4)
    fn_db = env.open_db(b"fn_db")
    with env.begin(db=fn_db, write=True) as txn:
        txn.put(b"x.jpg", b"\0" * 2048)
    indexer.check_weights_stamp(env, b)                       # unstamped store with rows: not claimed
    with pytest.raises(lmdb.Error):
        env.open_db(indexer.META_DB, create=False)            # and nothing was added to it
    env.close()
