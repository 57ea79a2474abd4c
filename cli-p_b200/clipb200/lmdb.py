"""Drop-in for the subset of `import lmdb` (py-lmdb) that CLI-P uses, reading and writing
LMDB's own on-disk format (`vectors.lmdb/data.mdb` + `lock.mdb`).

Reference call sites (under /root/reference):
  env = lmdb.open('vectors.lmdb', map_size=20 GiB, max_dbs=4)       build-index.py:22, query-index.py:25
  fn_db = env.open_db(b"fn_db"); skip_db; idx_db                    build-index.py:23-24,66; query-index.py:26-27
  with env.begin(db=fn_db, write=True) as txn: txn.get / txn.put    build-index.py:36-51,60-61,87-88
  txn.stat()['entries'], txn.cursor(), cursor.first(), iteration    build-index.py:68-76,83
  env.close()                                                       build-index.py:113

Neither py-lmdb nor liblmdb is present in this image, so this module implements the file
format itself [UPSTREAM liblmdb mdb.c, data version 1]: 4096-byte pages, two meta pages
(magic 0xBEEFC0DE; the one with the larger txnid wins), B+tree branch/leaf pages with 16-byte
headers and 2-byte node offsets, 8-byte node headers, values that do not fit in a node
(node > 2038 bytes: every 2048-byte embedding) on overflow pages, named databases as
F_SUBDATA records of the main database, memcmp key order, 511-byte key limit.
PARITY UNPINNED: files are verified only by this module's own independent reader
(tests/test_lmdb_store.py); when a real `lmdb` module is importable it is used instead.

Write model: a write transaction updates the in-memory B+tree image; commits are made
durable by rewriting data.mdb with a bulk-loaded, fully packed tree (write to a temp file +
atomic rename).  Rewrites are amortised (at most ~0.5 % of the entries, or `sync_every`
commits, may be pending) and forced by env.sync() / env.close(); the reference's
one-commit-per-image loop (build-index.py:42-51) therefore stays O(N log N) instead of O(N^2).
"""
from __future__ import annotations

import builtins
import os
import struct
from typing import Dict, Iterator, List, Optional, Tuple

def _find_real_lmdb():
    """A real py-lmdb wins when one is installed somewhere other than our alias directory."""
    import importlib.machinery
    import importlib.util
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in sys.path:
        ap = os.path.abspath(p or ".")
        if ap == here:
            continue
        try:
            spec = importlib.machinery.PathFinder.find_spec("lmdb", [ap])
        except Exception:
            spec = None
        if spec is not None and spec.origin and "clipb200" not in spec.origin:
            try:
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                return mod
            except Exception:
                return None
    return None


_real = _find_real_lmdb()

PAGE = 4096
HDR = 16
MAGIC = 0xBEEFC0DE
VERSION = 1
P_BRANCH, P_LEAF, P_OVERFLOW, P_META = 0x01, 0x02, 0x04, 0x08
F_BIGDATA, F_SUBDATA = 0x01, 0x02
MDB_INTEGERKEY = 0x08
P_INVALID = 0xFFFFFFFFFFFFFFFF
NODE_MAX = (((PAGE - HDR) // 2) & ~1) - 2        # 2038
MAX_KEY = 511
DB_FMT = "<IHHQQQQQ"                             # md_pad, md_flags, md_depth, branch, leaf, overflow, entries, root


class Error(Exception):
    pass


class BadValsizeError(Error):
    pass


class MapFullError(Error):
    pass


# =====================================================================================
# reader
# =====================================================================================

def _page(buf: memoryview, pgno: int) -> memoryview:
    off = pgno * PAGE
    if off + PAGE > len(buf):
        raise Error(f"page {pgno} beyond end of file")
    return buf[off:off + PAGE]


def _read_meta(buf: memoryview) -> dict:
    best = None
    for pg in (0, 1):
        if len(buf) < (pg + 1) * PAGE:
            continue
        p = _page(buf, pg)
        flags = struct.unpack_from("<H", p, 10)[0]
        magic, version = struct.unpack_from("<II", p, HDR)
        if not (flags & P_META) or magic != MAGIC:
            continue
        if version != VERSION:
            raise Error(f"unsupported LMDB data version {version}")
        _addr, mapsize = struct.unpack_from("<QQ", p, HDR + 8)
        free_db = struct.unpack_from(DB_FMT, p, HDR + 24)
        main_db = struct.unpack_from(DB_FMT, p, HDR + 24 + 48)
        last_pg, txnid = struct.unpack_from("<QQ", p, HDR + 24 + 96)
        m = {"mapsize": mapsize, "psize": free_db[0], "free": free_db, "main": main_db, "last_pg": last_pg,
             "txnid": txnid}
        if best is None or txnid > best["txnid"]:
            best = m
    if best is None:
        raise Error("no valid LMDB meta page")
    if best["psize"] != PAGE:
        raise Error(f"page size {best['psize']} not supported (only {PAGE})")
    return best


def _walk(buf: memoryview, root: int) -> Iterator[Tuple[bytes, bytes, int]]:
    """In-order (key, value, node flags) of the tree rooted at `root`."""
    if root == P_INVALID:
        return
    stack = [root]
    # explicit DFS keeping child order
    def visit(pgno):
        p = _page(buf, pgno)
        flags, lower = struct.unpack_from("<HH", p, 10)
        n = (lower - HDR) // 2
        ptrs = struct.unpack_from(f"<{n}H", p, HDR)
        if flags & P_BRANCH:
            for off in ptrs:
                lo, hi, fl, ks = struct.unpack_from("<HHHH", p, off)
                yield from visit(lo | (hi << 16) | (fl << 32))
        elif flags & P_LEAF:
            for off in ptrs:
                lo, hi, fl, ks = struct.unpack_from("<HHHH", p, off)
                dsize = lo | (hi << 16)
                key = bytes(p[off + 8:off + 8 + ks])
                if fl & F_BIGDATA:
                    ov = struct.unpack_from("<Q", p, off + 8 + ks)[0]
                    start = ov * PAGE + HDR
                    val = bytes(buf[start:start + dsize])
                else:
                    val = bytes(p[off + 8 + ks:off + 8 + ks + dsize])
                yield key, val, fl
        else:
            raise Error(f"page {pgno}: unexpected flags {flags:#x}")
    yield from visit(stack[0])


def read_file(path: str) -> Tuple[Dict[bytes, Dict[bytes, bytes]], dict]:
    """Parse data.mdb -> ({db name: {key: value}}, meta).  The unnamed main database's own
    plain records (none in CLI-P) are returned under the name b''."""
    with builtins.open(path, "rb") as fh:
        raw = fh.read()
    buf = memoryview(raw)
    meta = _read_meta(buf)
    dbs: Dict[bytes, Dict[bytes, bytes]] = {b"": {}}
    for key, val, fl in _walk(buf, meta["main"][7]):
        if fl & F_SUBDATA:
            sub = struct.unpack(DB_FMT, val)
            dbs[key] = {k: v for k, v, _ in _walk(buf, sub[7])}
        else:
            dbs[b""][key] = val
    return dbs, meta


# =====================================================================================
# writer (bulk load of sorted records)
# =====================================================================================

class _Out:
    def __init__(self):
        self.pages: List[bytes] = [b"", b""]      # meta pages patched last

    def alloc(self, n: int = 1) -> int:
        pg = len(self.pages)
        self.pages.extend([b""] * n)
        return pg


def _node_size(ks: int, ds: int) -> int:
    sz = 8 + ks + ds
    if sz > NODE_MAX:
        sz = 8 + ks + 8                           # value goes to overflow pages
    return (sz + 1) & ~1


def _build_page(pgno: int, flags: int, nodes: List[bytes]) -> bytes:
    page = bytearray(PAGE)
    upper = PAGE
    ptrs = []
    for nd in nodes:
        upper -= len(nd)
        page[upper:upper + len(nd)] = nd
        ptrs.append(upper)
    lower = HDR + 2 * len(nodes)
    assert lower <= upper, "page overfull"
    struct.pack_into("<QHHHH", page, 0, pgno, 0, flags, lower, upper)
    struct.pack_into(f"<{len(ptrs)}H", page, HDR, *ptrs)
    return bytes(page)


def _write_tree(out: _Out, items: List[Tuple[bytes, bytes, int]]) -> tuple:
    """items: sorted (key, value, node flags).  Returns the MDB_db tuple."""
    if not items:
        return (0, 0, 0, 0, 0, 0, 0, P_INVALID)
    n_leaf = n_branch = n_over = 0
    level: List[Tuple[bytes, int]] = []           # (first key, pgno) per page of the current level
    cur: List[bytes] = []
    cur_first: Optional[bytes] = None
    free = PAGE - HDR

    def flush_leaf():
        nonlocal cur, cur_first, free, n_leaf
        if not cur:
            return
        pg = out.alloc()
        out.pages[pg] = _build_page(pg, P_LEAF, cur)
        level.append((cur_first, pg))
        n_leaf += 1
        cur, cur_first, free = [], None, PAGE - HDR

    for key, val, fl in items:
        if not (0 < len(key) <= MAX_KEY):
            raise BadValsizeError(f"key length {len(key)} outside 1..{MAX_KEY}")
        ks, ds = len(key), len(val)
        big = 8 + ks + ds > NODE_MAX
        if big:
            npages = (HDR - 1 + ds) // PAGE + 1
            ov = out.alloc(npages)
            blob = bytearray(npages * PAGE)
            struct.pack_into("<QHHI", blob, 0, ov, 0, P_OVERFLOW, npages)
            blob[HDR:HDR + ds] = val
            for i in range(npages):
                out.pages[ov + i] = bytes(blob[i * PAGE:(i + 1) * PAGE])
            n_over += npages
            body = struct.pack("<HHHH", ds & 0xFFFF, ds >> 16, fl | F_BIGDATA, ks) + key + struct.pack("<Q", ov)
        else:
            body = struct.pack("<HHHH", ds & 0xFFFF, ds >> 16, fl, ks) + key + val
        if len(body) & 1:
            body += b"\0"
        need = len(body) + 2
        if need > free:
            flush_leaf()
        if cur_first is None:
            cur_first = key
        cur.append(body)
        free -= need
    flush_leaf()

    def branch_node(pg: int, kbytes: bytes) -> bytes:
        b = struct.pack("<HHHH", pg & 0xFFFF, (pg >> 16) & 0xFFFF, (pg >> 32) & 0xFFFF, len(kbytes)) + kbytes
        return b + (b"\0" if len(b) & 1 else b"")

    depth = 1
    while len(level) > 1:
        # greedy grouping of children into branch pages; the first key of a page is implicit (empty)
        groups: List[List[Tuple[bytes, int]]] = [[]]
        free = PAGE - HDR
        for k, pg in level:
            need = len(branch_node(pg, k if groups[-1] else b"")) + 2
            if need > free:
                groups.append([])
                free = PAGE - HDR
                need = len(branch_node(pg, b"")) + 2
            groups[-1].append((k, pg))
            free -= need
        if len(groups) > 1 and len(groups[-1]) < 2:       # liblmdb expects >= 2 keys per branch page
            groups[-1].insert(0, groups[-2].pop())
        nxt: List[Tuple[bytes, int]] = []
        for grp in groups:
            bp = out.alloc()
            nodes = [branch_node(pg, b"" if i == 0 else k) for i, (k, pg) in enumerate(grp)]
            out.pages[bp] = _build_page(bp, P_BRANCH, nodes)
            nxt.append((grp[0][0], bp))
            n_branch += 1
        level = nxt
        depth += 1
    return (0, 0, depth, n_branch, n_leaf, n_over, len(items), level[0][1])


def write_file(path: str, dbs: Dict[bytes, Dict[bytes, bytes]], mapsize: int, txnid: int) -> None:
    out = _Out()
    main_items: List[Tuple[bytes, bytes, int]] = [(k, v, 0) for k, v in dbs.get(b"", {}).items()]
    for name in dbs:
        if name == b"":
            continue
        rec = _write_tree(out, [(k, v, 0) for k, v in sorted(dbs[name].items())])
        main_items.append((name, struct.pack(DB_FMT, *rec), F_SUBDATA))
    main_items.sort(key=lambda t: t[0])
    main = _write_tree(out, main_items)
    last_pg = len(out.pages) - 1
    if (last_pg + 1) * PAGE > mapsize:
        raise MapFullError("environment map_size reached")
    for pg in (0, 1):
        page = bytearray(PAGE)
        struct.pack_into("<QHHHH", page, 0, pg, 0, P_META, 0, 0)
        struct.pack_into("<IIQQ", page, HDR, MAGIC, VERSION, 0, mapsize)
        struct.pack_into(DB_FMT, page, HDR + 24, PAGE, MDB_INTEGERKEY, 0, 0, 0, 0, 0, P_INVALID)
        struct.pack_into(DB_FMT, page, HDR + 24 + 48, *main)
        # the newer meta lives in page (txnid & 1); the other carries txnid - 1 and the same tree
        t = txnid if (txnid & 1) == pg else max(txnid - 1, 0)
        struct.pack_into("<QQ", page, HDR + 24 + 96, last_pg, t)
        out.pages[pg] = bytes(page)
    tmp = path + ".tmp"
    with builtins.open(tmp, "wb") as fh:
        fh.write(b"".join(out.pages))
        fh.flush()
        os.fsync(fh.fileno())
    os.replace(tmp, path)


# =====================================================================================
# py-lmdb shaped API
# =====================================================================================

class _Database:
    def __init__(self, name: bytes):
        self.name = name


class Cursor:
    def __init__(self, txn: "Transaction", db: _Database):
        self._txn, self._db = txn, db
        self._keys: List[bytes] = []
        self._i = -1

    def _snapshot(self):
        self._keys = sorted(self._txn._table(self._db))

    def first(self) -> bool:
        self._snapshot()
        self._i = 0
        return bool(self._keys)

    def next(self) -> bool:
        self._i += 1
        return self._i < len(self._keys)

    def key(self) -> bytes:
        return self._keys[self._i] if 0 <= self._i < len(self._keys) else b""

    def value(self) -> bytes:
        return self._txn._table(self._db).get(self.key(), b"")

    def item(self):
        return self.key(), self.value()

    def __iter__(self):
        # py-lmdb: iterating a positioned cursor starts at the current record;
        # an unpositioned one starts at the first record
        if self._i < 0:
            self._snapshot()
            self._i = 0
        tab = self._txn._table(self._db)
        while self._i < len(self._keys):
            k = self._keys[self._i]
            yield k, tab[k]
            self._i += 1

    iternext = __iter__


class Transaction:
    def __init__(self, env: "Environment", db: Optional[_Database], write: bool):
        self._env, self._db, self._write = env, db or env._main, write
        self._pending: Dict[bytes, Dict[bytes, Optional[bytes]]] = {}
        self._done = False

    def _table(self, db: Optional[_Database]) -> Dict[bytes, bytes]:
        name = (db or self._db).name
        base = self._env._dbs.setdefault(name, {})
        pend = self._pending.get(name)
        if not pend:
            return base
        merged = dict(base)
        for k, v in pend.items():
            if v is None:
                merged.pop(k, None)
            else:
                merged[k] = v
        return merged

    def get(self, key: bytes, default=None, db: Optional[_Database] = None):
        name = (db or self._db).name
        pend = self._pending.get(name)
        if pend is not None and key in pend:
            v = pend[key]
            return default if v is None else v
        return self._env._dbs.get(name, {}).get(key, default)

    def put(self, key: bytes, value: bytes, dupdata: bool = True, overwrite: bool = True, append: bool = False,
            db: Optional[_Database] = None) -> bool:
        if not self._write:
            raise Error("put() on a read-only transaction")
        key, value = bytes(key), bytes(value)
        if not (0 < len(key) <= MAX_KEY):
            raise BadValsizeError(f"key length {len(key)} outside 1..{MAX_KEY}")
        name = (db or self._db).name
        if not overwrite and self.get(key, db=db) is not None:
            return False
        self._pending.setdefault(name, {})[key] = value
        return True

    def delete(self, key: bytes, value: bytes = b"", db: Optional[_Database] = None) -> bool:
        if not self._write:
            raise Error("delete() on a read-only transaction")
        if self.get(key, db=db) is None:
            return False
        self._pending.setdefault((db or self._db).name, {})[bytes(key)] = None
        return True

    def stat(self, db: Optional[_Database] = None) -> dict:
        return {"psize": PAGE, "entries": len(self._table(db)), "depth": 0, "branch_pages": 0, "leaf_pages": 0,
                "overflow_pages": 0}

    def cursor(self, db: Optional[_Database] = None) -> Cursor:
        return Cursor(self, db or self._db)

    def commit(self) -> None:
        if self._done:
            return
        self._done = True
        if self._write:
            changed = 0
            for name, pend in self._pending.items():
                tab = self._env._dbs.setdefault(name, {})
                for k, v in pend.items():
                    if v is None:
                        tab.pop(k, None)
                    else:
                        tab[k] = v
                    changed += 1
            self._env._committed(changed)

    def abort(self) -> None:
        self._done = True
        self._pending.clear()

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is None:
            self.commit()
        else:
            self.abort()
        return False


class Environment:
    def __init__(self, path: str, map_size: int = 10485760, max_dbs: int = 0, readonly: bool = False,
                 subdir: bool = True, create: bool = True, sync_every: Optional[int] = None, **_ignored):
        self._dir = path
        self._map_size = int(map_size)
        self._max_dbs = max_dbs
        self._readonly = readonly
        self._sync_every = sync_every
        if subdir:
            if create and not readonly:
                os.makedirs(path, exist_ok=True)
            self._data = os.path.join(path, "data.mdb")
            lock = os.path.join(path, "lock.mdb")
        else:
            self._data, lock = path, path + "-lock"
        self._dbs: Dict[bytes, Dict[bytes, bytes]] = {b"": {}}
        self._txnid = 0
        if os.path.exists(self._data) and os.path.getsize(self._data) >= 2 * PAGE:
            self._dbs, meta = read_file(self._data)
            self._txnid = meta["txnid"]
            self._map_size = max(self._map_size, meta["mapsize"])
        elif not readonly and create:
            write_file(self._data, self._dbs, self._map_size, 0)
        else:
            raise Error(f"{self._data}: No such file or directory")
        if not readonly and not os.path.exists(lock):
            builtins.open(lock, "ab").close()          # liblmdb (re)initialises the lock file on first open
        self._main = _Database(b"")
        self._dirty = 0
        self._closed = False

    def open_db(self, key: Optional[bytes] = None, txn=None, create: bool = True, **_ignored) -> _Database:
        if key is None:
            return self._main
        key = bytes(key)
        if key not in self._dbs:
            if not create or self._readonly:
                raise Error(f"named database {key!r} not found")
            if self._max_dbs and len(self._dbs) - 1 >= self._max_dbs:
                raise Error("max_dbs reached (MDB_DBS_FULL)")
            self._dbs[key] = {}
            self._committed(1)
        return _Database(key)

    def begin(self, db: Optional[_Database] = None, parent=None, write: bool = False, buffers: bool = False) -> Transaction:
        if write and self._readonly:
            raise Error("write transaction on a read-only environment")
        return Transaction(self, db, write)

    def _committed(self, changed: int) -> None:
        if not changed:
            return
        self._txnid += 1
        self._dirty += 1
        n = sum(len(t) for t in self._dbs.values())
        limit = self._sync_every if self._sync_every is not None else max(1, n // 256)
        if self._dirty >= limit:
            self.sync()

    def sync(self, force: bool = False) -> None:
        if self._dirty and not self._readonly:
            write_file(self._data, self._dbs, self._map_size, self._txnid)
            self._dirty = 0

    def stat(self) -> dict:
        return {"psize": PAGE, "entries": len(self._dbs.get(b"", {})) + len(self._dbs) - 1}

    def info(self) -> dict:
        return {"map_size": self._map_size, "last_txnid": self._txnid}

    def path(self) -> str:
        return self._dir

    def close(self) -> None:
        if not self._closed:
            self.sync()
            self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def open(path: str, **kwargs):
    """lmdb.open(...): the real binding when importable, this implementation otherwise."""
    if _real is not None and not os.environ.get("CLIPB200_OWN_LMDB"):
        return _real.open(path, **kwargs)
    return Environment(path, **kwargs)
