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

Write model (liblmdb's own): copy-on-write.  A flush is one LMDB write transaction: every
page on the path of a changed record is copied to a fresh page at the END of the file, values
go to new overflow pages, the new pages are written and fsync'ed, then the meta page
(txnid & 1) is flipped and fsync'ed.  Nothing is rewritten, RAM holds only the pages of the
flush in progress, and a crash leaves the previous meta page -- i.e. the previous commit --
intact.  Reads walk the B+tree through an mmap of data.mdb.

Two deliberate differences from liblmdb, both documented in README / INTEGRATION:
  * Commits are GROUPED: committed transactions accumulate in a small in-memory buffer that is
    flushed when it holds `flush_records` records (default 4096), is `flush_seconds` old
    (default 1 s), on cursor()/stat()/sync()/close(), or at every commit with sync=True.  A
    kill or power loss can therefore lose up to ~1 s of commits (liblmdb: none).  The reference
    commits one image per transaction (build-index.py:42-51); at index-time rates of 10^4
    images/s a per-commit fsync would be the bottleneck.
  * Freed pages are not recycled (no free-list records are written), so the file grows by the
    copied pages; `env.copy(path, compact=True)` writes a packed copy.  A real liblmdb can
    read and extend the file.
The first write of an environment takes an exclusive flock on lock.mdb (held until close): a
second WRITING process is refused instead of silently overwriting the first (liblmdb would
serialise them on its own mutex); environments that only read never lock and see the newest
commit at every begin().
"""
from __future__ import annotations

import bisect
import builtins
import mmap
import os
import struct
import time
from typing import Dict, Iterator, List, Optional, Tuple


def _find_real_lmdb():
    """A real py-lmdb wins when one is installed somewhere other than our alias directory.  It is
    imported the normal way (importlib.import_module, registered in sys.modules, so its own
    `import lmdb.cpython` resolves) with the alias directory taken off sys.path for the duration."""
    import importlib
    import importlib.machinery
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    others = [p for p in sys.path if os.path.abspath(p or ".") != here]
    try:
        spec = importlib.machinery.PathFinder.find_spec("lmdb", others)
    except Exception:
        spec = None
    if spec is None or not spec.origin or "clipb200" in spec.origin:
        return None
    saved_path, saved_mod = list(sys.path), sys.modules.pop("lmdb", None)
    try:
        sys.path[:] = others
        mod = importlib.import_module("lmdb")
        sys.modules["_clipb200_real_lmdb"] = mod
        return mod
    except Exception as e:                                   # a broken install is reported, not ignored
        print(f"clipb200.lmdb: a real lmdb package exists at {spec.origin} but failed to import ({e}); "
              "using the built-in store", file=sys.stderr)
        return None
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "lmdb" or k.startswith("lmdb.")]:
            if saved_mod is not None and k == "lmdb":
                continue
            sys.modules.setdefault("_clipb200_real_" + k, sys.modules[k])
        if saved_mod is not None:
            sys.modules["lmdb"] = saved_mod


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
# bulk writer (packed copy of sorted records; streams, O(tree height) memory per level)
# =====================================================================================

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


def _leaf_node(key: bytes, val: bytes, fl: int, ov_pgno: Optional[int]) -> bytes:
    ds = len(val)
    if ov_pgno is not None:
        body = struct.pack("<HHHH", ds & 0xFFFF, ds >> 16, fl | F_BIGDATA, len(key)) + key + struct.pack("<Q", ov_pgno)
    else:
        body = struct.pack("<HHHH", ds & 0xFFFF, ds >> 16, fl, len(key)) + key + val
    return body + (b"\0" if len(body) & 1 else b"")


def _branch_node(pg: int, kbytes: bytes) -> bytes:
    b = struct.pack("<HHHH", pg & 0xFFFF, (pg >> 16) & 0xFFFF, (pg >> 32) & 0xFFFF, len(kbytes)) + kbytes
    return b + (b"\0" if len(b) & 1 else b"")


def _overflow_pages(first: int, val: bytes) -> List[bytes]:
    npages = (HDR - 1 + len(val)) // PAGE + 1
    blob = bytearray(npages * PAGE)
    struct.pack_into("<QHHI", blob, 0, first, 0, P_OVERFLOW, npages)
    blob[HDR:HDR + len(val)] = val
    return [bytes(blob[i * PAGE:(i + 1) * PAGE]) for i in range(npages)]


class _Bulk:
    """Writes pages 2.. of a fresh file in order; records are fed sorted, one tree at a time."""

    def __init__(self, fh):
        self.fh = fh
        self.next_pg = 2
        fh.seek(2 * PAGE)

    def _emit(self, pages: List[bytes]) -> int:
        first = self.next_pg
        self.fh.write(b"".join(pages))
        self.next_pg += len(pages)
        return first

    def tree(self, items: Iterator[Tuple[bytes, bytes, int]]) -> tuple:
        n_leaf = n_branch = n_over = n_items = 0
        level: List[Tuple[bytes, int]] = []
        cur: List[bytes] = []
        cur_first: Optional[bytes] = None
        free = PAGE - HDR

        def flush_leaf():
            nonlocal cur, cur_first, free, n_leaf
            if cur:
                pg = self.next_pg
                self._emit([_build_page(pg, P_LEAF, cur)])
                level.append((cur_first, pg))
                n_leaf += 1
                cur, cur_first, free = [], None, PAGE - HDR

        for key, val, fl in items:
            if not (0 < len(key) <= MAX_KEY):
                raise BadValsizeError(f"key length {len(key)} outside 1..{MAX_KEY}")
            n_items += 1
            ov = None
            if 8 + len(key) + len(val) > NODE_MAX:
                pages = _overflow_pages(self.next_pg, val)
                ov = self._emit(pages)
                n_over += len(pages)
            body = _leaf_node(key, val, fl, ov)
            if len(body) + 2 > free:
                flush_leaf()
            if cur_first is None:
                cur_first = key
            cur.append(body)
            free -= len(body) + 2
        flush_leaf()
        if not level:
            return (0, 0, 0, 0, 0, 0, 0, P_INVALID)
        depth = 1
        while len(level) > 1:
            groups: List[List[Tuple[bytes, int]]] = [[]]
            free = PAGE - HDR
            for k, pg in level:
                need = len(_branch_node(pg, k if groups[-1] else b"")) + 2
                if need > free:
                    groups.append([])
                    free = PAGE - HDR
                    need = len(_branch_node(pg, b"")) + 2
                groups[-1].append((k, pg))
                free -= need
            if len(groups) > 1 and len(groups[-1]) < 2:       # liblmdb expects >= 2 keys per branch page
                groups[-1].insert(0, groups[-2].pop())
            nxt: List[Tuple[bytes, int]] = []
            for grp in groups:
                bp = self.next_pg
                self._emit([_build_page(bp, P_BRANCH, [_branch_node(pg, b"" if i == 0 else k)
                                                      for i, (k, pg) in enumerate(grp)])])
                nxt.append((grp[0][0], bp))
                n_branch += 1
            level = nxt
            depth += 1
        return (0, 0, depth, n_branch, n_leaf, n_over, n_items, level[0][1])


def _meta_page(pg: int, mapsize: int, main: tuple, last_pg: int, txnid: int) -> bytes:
    page = bytearray(PAGE)
    struct.pack_into("<QHHHH", page, 0, pg, 0, P_META, 0, 0)
    struct.pack_into("<IIQQ", page, HDR, MAGIC, VERSION, 0, mapsize)
    struct.pack_into(DB_FMT, page, HDR + 24, PAGE, MDB_INTEGERKEY, 0, 0, 0, 0, 0, P_INVALID)
    struct.pack_into(DB_FMT, page, HDR + 24 + 48, *main)
    struct.pack_into("<QQ", page, HDR + 24 + 96, last_pg, txnid)
    return bytes(page)


def write_packed(path: str, named: Dict[bytes, Iterator[Tuple[bytes, bytes]]], mapsize: int, txnid: int,
                 main_plain: Optional[Dict[bytes, bytes]] = None) -> None:
    """A fresh, fully packed data.mdb: `named` maps database names to SORTED (key, value) iterators."""
    tmp = path + ".tmp"
    with builtins.open(tmp, "wb") as fh:
        bulk = _Bulk(fh)
        main_items: List[Tuple[bytes, bytes, int]] = [(k, v, 0) for k, v in (main_plain or {}).items()]
        for name in sorted(named):
            rec = bulk.tree((k, v, 0) for k, v in named[name])
            main_items.append((name, struct.pack(DB_FMT, *rec), F_SUBDATA))
        main_items.sort(key=lambda t: t[0])
        main = bulk.tree(iter(main_items))
        last_pg = bulk.next_pg - 1
        if (last_pg + 1) * PAGE > mapsize:
            raise MapFullError("environment map_size reached")
        fh.seek(0)
        # the newer meta lives in page (txnid & 1); the other carries txnid - 1 and the same tree
        for pg in (0, 1):
            fh.write(_meta_page(pg, mapsize, main, max(last_pg, 1), txnid if (txnid & 1) == pg else max(txnid - 1, 0)))
        fh.flush()
        os.fsync(fh.fileno())
    os.replace(tmp, path)


def write_file(path: str, dbs: Dict[bytes, Dict[bytes, bytes]], mapsize: int, txnid: int) -> None:
    """Packed file from in-memory tables (small stores, tests)."""
    write_packed(path, {n: iter(sorted(t.items())) for n, t in dbs.items() if n != b""}, mapsize, txnid,
                 main_plain=dbs.get(b"", {}))


# =====================================================================================
# copy-on-write B+tree over data.mdb
# =====================================================================================
_DEPTH, _BRANCH, _LEAF, _OVER, _ENTRIES, _ROOT = 2, 3, 4, 5, 6, 7
_CAP = PAGE - HDR


class _Pg:
    """A decoded branch / leaf page: parallel lists of keys and encoded nodes (key order)."""
    __slots__ = ("pgno", "flags", "keys", "nodes", "used")

    def __init__(self, pgno: int, flags: int, keys: List[bytes], nodes: List[bytes]):
        self.pgno, self.flags, self.keys, self.nodes = pgno, flags, keys, nodes
        self.used = sum(len(n) for n in nodes) + 2 * len(nodes)

    @property
    def is_branch(self) -> bool:
        return bool(self.flags & P_BRANCH)

    @staticmethod
    def decode(buf, pgno: int) -> "_Pg":
        off0 = pgno * PAGE
        flags, lower = struct.unpack_from("<HH", buf, off0 + 10)
        if not flags & (P_BRANCH | P_LEAF):
            raise Error(f"page {pgno}: unexpected flags {flags:#x}")
        n = (lower - HDR) // 2
        ptrs = struct.unpack_from(f"<{n}H", buf, off0 + HDR)
        keys, nodes = [], []
        leaf = bool(flags & P_LEAF)
        for off in ptrs:
            lo, hi, fl, ks = struct.unpack_from("<HHHH", buf, off0 + off)
            if leaf:
                size = 8 + ks + (8 if fl & F_BIGDATA else (lo | (hi << 16)))
            else:
                size = 8 + ks
            size += size & 1
            nodes.append(bytes(buf[off0 + off:off0 + off + size]))
            keys.append(bytes(buf[off0 + off + 8:off0 + off + 8 + ks]))
        return _Pg(pgno, flags & (P_BRANCH | P_LEAF), keys, nodes)

    def encode(self) -> bytes:
        return _build_page(self.pgno, self.flags, self.nodes)

    def child(self, i: int) -> int:
        lo, hi, fl = struct.unpack_from("<HHH", self.nodes[i], 0)
        return lo | (hi << 16) | (fl << 32)

    def set_child(self, i: int, pgno: int) -> None:
        self.nodes[i] = _branch_node(pgno, self.keys[i])

    def child_index(self, key: bytes) -> int:
        # the first key of a branch page is implicit (-inf); the child covering `key` is the last one
        # whose separator is <= key
        return max(0, bisect.bisect_right(self.keys, key, 1) - 1)

    def find(self, key: bytes) -> Tuple[int, bool]:
        i = bisect.bisect_left(self.keys, key)
        return i, i < len(self.keys) and self.keys[i] == key

    def insert(self, i: int, key: bytes, node: bytes) -> None:
        self.keys.insert(i, key)
        self.nodes.insert(i, node)
        self.used += len(node) + 2

    def replace(self, i: int, node: bytes) -> None:
        self.used += len(node) - len(self.nodes[i])
        self.nodes[i] = node

    def remove(self, i: int) -> None:
        self.used -= len(self.nodes[i]) + 2
        del self.keys[i]
        del self.nodes[i]


def _node_value(buf_page, node: bytes, ks: int):
    """(is_big, inline value or overflow pgno, data size, node flags) of an encoded leaf node."""
    lo, hi, fl = struct.unpack_from("<HHH", node, 0)
    ds = lo | (hi << 16)
    if fl & F_BIGDATA:
        return True, struct.unpack_from("<Q", node, 8 + ks)[0], ds, fl
    return False, node[8 + ks:8 + ks + ds], ds, fl


class _Store:
    """The file: meta, mmap for reads, and the copy-on-write flush."""

    def __init__(self, path: str, map_size: int, readonly: bool, create: bool):
        self.path, self.readonly = path, readonly
        self.map_size = int(map_size)
        fresh = not (os.path.exists(path) and os.path.getsize(path) >= 2 * PAGE)
        if fresh:
            if readonly or not create:
                raise Error(f"{path}: No such file or directory")
            write_packed(path, {}, self.map_size, 0)
        self.fh = builtins.open(path, "rb" if readonly else "r+b")
        self.mm: Optional[mmap.mmap] = None
        self.garbage = 0                             # pages made unreachable by flushes of this session
        self.reload()

    def disk_txnid(self) -> int:
        """txnid of the newest meta page as it is on disk now (two 8-byte reads)."""
        best = -1
        for pg in (0, 1):
            self.fh.seek(pg * PAGE + HDR)
            magic = struct.unpack("<I", self.fh.read(4))[0]
            self.fh.seek(pg * PAGE + HDR + 24 + 96 + 8)
            t = struct.unpack("<Q", self.fh.read(8))[0]
            if magic == MAGIC:
                best = max(best, t)
        return best

    def reload(self) -> None:
        """(Re)read the newest meta page and the named-database records: another process may have committed."""
        self._remap()
        meta = _read_meta(memoryview(self.mm))
        self.main = list(meta["main"])
        self.last_pg = max(int(meta["last_pg"]), 1)
        self.txnid = int(meta["txnid"])
        self.map_size = max(self.map_size, int(meta["mapsize"]))
        self.recs: Dict[bytes, list] = {}            # named database -> MDB_db fields (mutable)
        self.plain: Dict[bytes, bytes] = {}          # plain records of the main database (none in CLI-P)
        for key, val, fl in self._iter_tree(self.main[_ROOT]):
            if fl & F_SUBDATA:
                self.recs[key] = list(struct.unpack(DB_FMT, val))
            else:
                self.plain[key] = val

    def _remap(self) -> None:
        if self.mm is not None:
            self.mm.close()
        self.fh.flush()
        self.mm = mmap.mmap(self.fh.fileno(), 0, access=mmap.ACCESS_READ)

    def close(self) -> None:
        if self.mm is not None:
            self.mm.close()
            self.mm = None
        self.fh.close()

    # ---- reads ------------------------------------------------------------------------------
    def _value(self, node: bytes, ks: int) -> bytes:
        big, v, ds, _ = _node_value(None, node, ks)
        if not big:
            return bytes(v)
        start = v * PAGE + HDR
        return bytes(self.mm[start:start + ds])

    def get(self, root: int, key: bytes) -> Optional[bytes]:
        """Point lookup: binary search over each page's node pointers straight in the mmap (no page decode)."""
        if root == P_INVALID:
            return None
        mm, pgno = self.mm, root
        while True:
            base = pgno * PAGE
            flags, lower = struct.unpack_from("<HH", mm, base + 10)
            n = (lower - HDR) // 2
            branch = bool(flags & P_BRANCH)
            lo, hi = (1, n) if branch else (0, n)     # a branch page's first key is implicit
            while lo < hi:                            # first node whose key is > key (branch) / >= key (leaf)
                mid = (lo + hi) >> 1
                off = base + struct.unpack_from("<H", mm, base + HDR + 2 * mid)[0]
                ks = struct.unpack_from("<H", mm, off + 6)[0]
                k = mm[off + 8:off + 8 + ks]
                if (k <= key) if branch else (k < key):
                    lo = mid + 1
                else:
                    hi = mid
            if branch:
                off = base + struct.unpack_from("<H", mm, base + HDR + 2 * (lo - 1))[0]
                a, b, c = struct.unpack_from("<HHH", mm, off)
                pgno = a | (b << 16) | (c << 32)
                continue
            if lo >= n:
                return None
            off = base + struct.unpack_from("<H", mm, base + HDR + 2 * lo)[0]
            a, b, fl, ks = struct.unpack_from("<HHHH", mm, off)
            if mm[off + 8:off + 8 + ks] != key:
                return None
            ds = a | (b << 16)
            if fl & F_BIGDATA:
                start = struct.unpack_from("<Q", mm, off + 8 + ks)[0] * PAGE + HDR
                return bytes(mm[start:start + ds])
            return bytes(mm[off + 8 + ks:off + 8 + ks + ds])

    def _iter_tree(self, root: int, start: Optional[bytes] = None) -> Iterator[Tuple[bytes, bytes, int]]:
        """In-order (key, value, node flags) from the first key >= start."""
        if root == P_INVALID:
            return
        stack: List[Tuple[_Pg, int]] = []
        pg = _Pg.decode(self.mm, root)
        while pg.is_branch:                          # descend to the starting leaf
            i = pg.child_index(start) if start is not None else 0
            stack.append((pg, i))
            pg = _Pg.decode(self.mm, pg.child(i))
        i = pg.find(start)[0] if start is not None else 0
        while True:
            while i < len(pg.keys):
                k = pg.keys[i]
                fl = struct.unpack_from("<H", pg.nodes[i], 4)[0]
                yield k, self._value(pg.nodes[i], len(k)), fl
                i += 1
            # next leaf
            while stack and stack[-1][1] + 1 >= len(stack[-1][0].keys):
                stack.pop()
            if not stack:
                return
            parent, j = stack.pop()
            stack.append((parent, j + 1))
            pg = _Pg.decode(self.mm, parent.child(j + 1))
            while pg.is_branch:
                stack.append((pg, 0))
                pg = _Pg.decode(self.mm, pg.child(0))
            i = 0

    # ---- one copy-on-write write transaction ---------------------------------------------------
    def flush(self, pending: Dict[bytes, Dict[bytes, Optional[bytes]]], new_dbs: List[bytes], sync: bool = True) -> None:
        if self.readonly:
            raise Error("flush on a read-only environment")
        dirty: Dict[int, _Pg] = {}
        raw: Dict[int, bytes] = {}                   # overflow pages (already encoded)
        state = {"next": self.last_pg + 1, "freed": 0}

        def alloc(n: int = 1) -> int:
            pg = state["next"]
            state["next"] += n
            if state["next"] * PAGE > self.map_size:
                raise MapFullError("environment map_size reached")
            return pg

        def page(pgno: int) -> _Pg:
            return dirty.get(pgno) or _Pg.decode(self.mm, pgno)

        def cow(pgno: int) -> _Pg:
            pg = dirty.get(pgno)
            if pg is None:
                pg = _Pg.decode(self.mm, pgno)
                pg.pgno = alloc()
                dirty[pg.pgno] = pg
                state["freed"] += 1
            return pg

        def new_page(flags: int) -> _Pg:
            pg = _Pg(alloc(), flags, [], [])
            dirty[pg.pgno] = pg
            return pg

        def descend(rec: list, key: bytes):
            root = cow(rec[_ROOT])
            rec[_ROOT] = root.pgno
            path: List[Tuple[_Pg, int]] = []
            pg = root
            while pg.is_branch:
                i = pg.child_index(key)
                child = cow(pg.child(i))
                if child.pgno != pg.child(i):
                    pg.set_child(i, child.pgno)
                path.append((pg, i))
                pg = child
            return path, pg

        def split(rec: list, path, pg: _Pg, at: int) -> None:
            n = len(pg.keys)
            if not pg.is_branch and at == n - 1:
                s = n - 1                            # append at the right edge: the full page stays full
            else:
                half, acc, s = pg.used // 2, 0, 1
                for j in range(n):
                    acc += len(pg.nodes[j]) + 2
                    if acc >= half:
                        s = j + 1
                        break
                s = min(max(s, 2 if pg.is_branch else 1), n - (2 if pg.is_branch else 1))
            right = new_page(pg.flags)
            right.keys, right.nodes = pg.keys[s:], pg.nodes[s:]
            del pg.keys[s:], pg.nodes[s:]
            for p_ in (pg, right):
                p_.used = sum(len(x) for x in p_.nodes) + 2 * len(p_.nodes)
            sep = right.keys[0]
            if pg.is_branch:
                right.used -= len(right.nodes[0])
                right.nodes[0] = _branch_node(right.child(0), b"")
                right.keys[0] = b""
                right.used += len(right.nodes[0])
                rec[_BRANCH] += 1
            else:
                rec[_LEAF] += 1
            assert pg.used <= _CAP and right.used <= _CAP, "split produced an overfull page"
            if not path:
                root = new_page(P_BRANCH)
                root.insert(0, b"", _branch_node(pg.pgno, b""))
                root.insert(1, sep, _branch_node(right.pgno, sep))
                rec[_ROOT] = root.pgno
                rec[_DEPTH] += 1
                rec[_BRANCH] += 1
                return
            parent, i = path.pop()
            parent.insert(i + 1, sep, _branch_node(right.pgno, sep))
            if parent.used > _CAP:
                split(rec, path, parent, i + 1)

        def old_overflow(rec: list, node: bytes, ks: int) -> None:
            big, v, ds, _ = _node_value(None, node, ks)
            if big:
                npg = (HDR - 1 + ds) // PAGE + 1
                rec[_OVER] -= npg
                state["freed"] += npg

        def put(rec: list, key: bytes, val: bytes, fl: int = 0) -> None:
            ov = None
            if 8 + len(key) + len(val) > NODE_MAX:
                pages = _overflow_pages(0, val)
                ov = alloc(len(pages))
                pages = _overflow_pages(ov, val)
                for j, b in enumerate(pages):
                    raw[ov + j] = b
                rec[_OVER] += len(pages)
            node = _leaf_node(key, val, fl, ov)
            if rec[_ROOT] == P_INVALID:
                leaf = new_page(P_LEAF)
                leaf.insert(0, key, node)
                rec[_ROOT], rec[_DEPTH], rec[_LEAF], rec[_ENTRIES] = leaf.pgno, 1, 1, 1
                return
            path, leaf = descend(rec, key)
            i, found = leaf.find(key)
            if found:
                old_overflow(rec, leaf.nodes[i], len(key))
                leaf.replace(i, node)
            else:
                leaf.insert(i, key, node)
                rec[_ENTRIES] += 1
            if leaf.used > _CAP:
                split(rec, path, leaf, i)

        def delete(rec: list, key: bytes) -> None:
            if rec[_ROOT] == P_INVALID:
                return
            path, leaf = descend(rec, key)
            i, found = leaf.find(key)
            if not found:
                return
            old_overflow(rec, leaf.nodes[i], len(key))
            leaf.remove(i)
            rec[_ENTRIES] -= 1
            pg = leaf
            # an emptied page is unlinked from its parent; a branch left with one child is replaced by it
            while not pg.keys:
                rec[_BRANCH if pg.is_branch else _LEAF] -= 1
                dirty.pop(pg.pgno, None)
                if not path:
                    rec[_ROOT], rec[_DEPTH] = P_INVALID, 0
                    return
                parent, j = path.pop()
                parent.remove(j)
                if j == 0 and parent.keys:           # the new first child's separator becomes implicit
                    parent.used -= len(parent.nodes[0])
                    parent.nodes[0] = _branch_node(parent.child(0), b"")
                    parent.keys[0] = b""
                    parent.used += len(parent.nodes[0])
                pg = parent
            while pg.is_branch and len(pg.keys) == 1:
                only = pg.child(0)
                rec[_BRANCH] -= 1
                dirty.pop(pg.pgno, None)
                if not path:
                    rec[_ROOT] = only
                    rec[_DEPTH] -= 1
                    return
                parent, j = path.pop()
                parent.set_child(j, only)
                pg = parent

        touched = set(new_dbs)
        for name in new_dbs:
            self.recs.setdefault(name, [0, 0, 0, 0, 0, 0, 0, P_INVALID])
        for name, table in pending.items():
            if not table:
                continue
            if name == b"":
                for k, v in table.items():
                    if v is None:
                        self.plain.pop(k, None)
                    else:
                        self.plain[k] = v
                touched.add(b"")
                continue
            rec = self.recs.setdefault(name, [0, 0, 0, 0, 0, 0, 0, P_INVALID])
            touched.add(name)
            for k in sorted(table):                  # sorted: consecutive keys share their (cached, dirty) path
                v = table[k]
                if v is None:
                    delete(rec, k)
                else:
                    put(rec, k, v)
        if not touched:
            return
        # the main database: one F_SUBDATA record per named database that changed
        main = self.main
        for name in sorted(touched):
            if name == b"":
                continue
            put(main, name, struct.pack(DB_FMT, *self.recs[name]), F_SUBDATA)
        if b"" in touched:
            for k, v in pending.get(b"", {}).items():
                if v is None:
                    delete(main, k)
                else:
                    put(main, k, v)
        # data pages first ...
        first, last = self.last_pg + 1, state["next"] - 1
        if last >= first:
            blob = bytearray((last - first + 1) * PAGE)
            for pgno, pg in dirty.items():
                blob[(pgno - first) * PAGE:(pgno - first + 1) * PAGE] = pg.encode()
            for pgno, b in raw.items():
                blob[(pgno - first) * PAGE:(pgno - first + 1) * PAGE] = b
            self.fh.seek(first * PAGE)
            self.fh.write(blob)
            self.fh.flush()
            if sync:
                os.fsync(self.fh.fileno())
        # ... then the meta page flips
        self.txnid += 1
        self.last_pg = max(last, self.last_pg)
        self.fh.seek((self.txnid & 1) * PAGE)
        self.fh.write(_meta_page(self.txnid & 1, self.map_size, tuple(main), self.last_pg, self.txnid))
        self.fh.flush()
        if sync:
            os.fsync(self.fh.fileno())
        self.garbage += state["freed"]
        self._remap()


# =====================================================================================
# py-lmdb shaped API
# =====================================================================================

class _Database:
    def __init__(self, name: bytes):
        self.name = name


class Cursor:
    """Forward cursor over a snapshot of the tree (pending commits are flushed first)."""

    def __init__(self, txn: "Transaction", db: _Database):
        self._txn, self._db = txn, db
        self._it: Optional[Iterator] = None
        self._cur: Optional[Tuple[bytes, bytes]] = None

    def _open(self, start: Optional[bytes] = None):
        env = self._txn._env
        self._txn._apply_own()
        env._flush()
        rec = env._store.recs.get(self._db.name)
        root = rec[_ROOT] if rec else P_INVALID
        if self._db.name == b"":
            root = env._store.main[_ROOT]
        self._it = ((k, v) for k, v, fl in env._store._iter_tree(root, start) if not fl & F_SUBDATA)

    def first(self) -> bool:
        self._open()
        return self.next()

    def set_range(self, key: bytes) -> bool:
        self._open(bytes(key))
        return self.next()

    def next(self) -> bool:
        if self._it is None:
            self._open()
        self._cur = next(self._it, None)
        return self._cur is not None

    def key(self) -> bytes:
        return self._cur[0] if self._cur else b""

    def value(self) -> bytes:
        return self._cur[1] if self._cur else b""

    def item(self):
        return self.key(), self.value()

    def __iter__(self):
        # py-lmdb: iterating a positioned cursor starts at the current record;
        # an unpositioned one starts at the first record
        if self._it is None:
            self._open()
            self._cur = next(self._it, None)
        while self._cur is not None:
            yield self._cur
            self._cur = next(self._it, None)

    iternext = __iter__


class Transaction:
    def __init__(self, env: "Environment", db: Optional[_Database], write: bool):
        self._env, self._db, self._write = env, db or env._main, write
        self._pending: Dict[bytes, Dict[bytes, Optional[bytes]]] = {}
        self._done = False

    def _apply_own(self) -> None:
        """A cursor / stat inside a write transaction sees the transaction's own writes: they are moved
        into the environment's buffer (this shim has a single writer; abort after that is not undone)."""
        if self._pending:
            self._env._absorb(self._pending)
            self._pending = {}

    def get(self, key: bytes, default=None, db: Optional[_Database] = None):
        name = (db or self._db).name
        key = bytes(key)
        for layer in (self._pending.get(name), self._env._pending.get(name)):
            if layer is not None and key in layer:
                v = layer[key]
                return default if v is None else v
        v = self._env._tree_get(name, key)
        return default if v is None else v

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
        self._apply_own()
        self._env._flush()
        name = (db or self._db).name
        rec = self._env._store.main if name == b"" else self._env._store.recs.get(name, [0] * 8)
        return {"psize": PAGE, "entries": rec[_ENTRIES], "depth": rec[_DEPTH], "branch_pages": rec[_BRANCH],
                "leaf_pages": rec[_LEAF], "overflow_pages": rec[_OVER]}

    def cursor(self, db: Optional[_Database] = None) -> Cursor:
        return Cursor(self, db or self._db)

    def commit(self) -> None:
        if self._done:
            return
        self._done = True
        if self._write and self._pending:
            self._env._absorb(self._pending)
            self._pending = {}
            self._env._committed()

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
                 subdir: bool = True, create: bool = True, sync: bool = False, flush_records: int = 4096,
                 flush_seconds: float = 1.0, sync_every: Optional[int] = None, lock: bool = True, **_ignored):
        self._dir = path
        self._max_dbs = max_dbs
        self._readonly = readonly
        self._sync_each_commit = bool(sync)
        self._flush_records = int(sync_every) if sync_every is not None else int(flush_records)
        self._flush_seconds = float(flush_seconds)
        if subdir:
            if create and not readonly:
                os.makedirs(path, exist_ok=True)
            self._data = os.path.join(path, "data.mdb")
            lockp = os.path.join(path, "lock.mdb")
        else:
            self._data, lockp = path, path + "-lock"
        self._lockp, self._use_lock = lockp, bool(lock)
        self._lock_fh = None
        self._locked = False
        self._store = _Store(self._data, map_size, readonly, create)
        if not readonly and not os.path.exists(lockp):
            builtins.open(lockp, "ab").close()        # liblmdb (re)initialises the lock file on first open
        self._main = _Database(b"")
        self._pending: Dict[bytes, Dict[bytes, Optional[bytes]]] = {}
        self._new_dbs: List[bytes] = []
        self._n_pending = 0
        self._oldest = 0.0
        self._closed = False

    # ---- buffer of committed, not yet flushed transactions ---------------------------------------
    def _absorb(self, pending: Dict[bytes, Dict[bytes, Optional[bytes]]]) -> None:
        for name, tab in pending.items():
            dst = self._pending.setdefault(name, {})
            for k, v in tab.items():
                if k not in dst:
                    self._n_pending += 1
                dst[k] = v
        if self._n_pending and not self._oldest:
            self._oldest = time.monotonic()

    def _committed(self) -> None:
        if (self._sync_each_commit or self._n_pending >= self._flush_records
                or (self._oldest and time.monotonic() - self._oldest >= self._flush_seconds)):
            self._flush()

    def _take_write_lock(self) -> None:
        """Single writer: the first flush takes an exclusive flock on lock.mdb and keeps it until close().
        An environment that only reads (query-index.py opens read-write, as the reference does, and never
        writes) never takes it, so a query session can run beside a builder.  If another process committed
        since this environment was opened, its view is reloaded before anything is written."""
        if self._locked:
            return
        if self._use_lock:
            self._lock_fh = builtins.open(self._lockp, "ab")
            try:
                import fcntl
                fcntl.flock(self._lock_fh.fileno(), fcntl.LOCK_EX | fcntl.LOCK_NB)
            except ImportError:
                pass
            except OSError:
                self._lock_fh.close()
                self._lock_fh = None
                raise Error(f"{self._lockp}: another process is writing to this environment; two builders would "
                            "overwrite each other's commits (README.md:49-51 of the reference warns about "
                            "exactly this) -- wait for it to finish") from None
        self._locked = True
        if self._store.disk_txnid() != self._store.txnid:
            self._store.reload()

    def _flush(self) -> None:
        if self._readonly or not (self._n_pending or self._new_dbs):
            return
        self._take_write_lock()
        self._store.flush(self._pending, self._new_dbs, sync=True)
        self._pending, self._new_dbs, self._n_pending, self._oldest = {}, [], 0, 0.0

    def _tree_get(self, name: bytes, key: bytes) -> Optional[bytes]:
        if name == b"":
            return self._store.plain.get(key)
        rec = self._store.recs.get(name)
        return self._store.get(rec[_ROOT], key) if rec else None

    # ---- py-lmdb surface ---------------------------------------------------------------------------
    def open_db(self, key: Optional[bytes] = None, txn=None, create: bool = True, **_ignored) -> _Database:
        if key is None:
            return self._main
        key = bytes(key)
        if key not in self._store.recs and key not in self._new_dbs:
            if not create or self._readonly:
                raise Error(f"named database {key!r} not found")
            if self._max_dbs and len(self._store.recs) + len(self._new_dbs) >= self._max_dbs:
                raise Error("max_dbs reached (MDB_DBS_FULL)")
            self._new_dbs.append(key)
            self._flush()                             # like liblmdb: creating a database is a committed write
        return _Database(key)

    def begin(self, db: Optional[_Database] = None, parent=None, write: bool = False, buffers: bool = False) -> Transaction:
        if write and self._readonly:
            raise Error("write transaction on a read-only environment")
        # a transaction of an environment that is not the writer sees the newest commit on disk, as in liblmdb
        if not self._locked and not self._n_pending and self._store.disk_txnid() != self._store.txnid:
            self._store.reload()
        return Transaction(self, db, write)

    def sync(self, force: bool = False) -> None:
        self._flush()

    def stat(self) -> dict:
        m = self._store.main
        return {"psize": PAGE, "entries": m[_ENTRIES], "depth": m[_DEPTH], "branch_pages": m[_BRANCH],
                "leaf_pages": m[_LEAF], "overflow_pages": m[_OVER]}

    def info(self) -> dict:
        return {"map_size": self._store.map_size, "last_txnid": self._store.txnid, "last_pgno": self._store.last_pg,
                "pending_records": self._n_pending, "garbage_pages": self._store.garbage}

    def path(self) -> str:
        return self._dir

    def copy(self, path: str, compact: bool = False, txn=None) -> None:
        """env.copy(path, compact=True): a packed copy (every page reachable, none wasted)."""
        self._flush()
        os.makedirs(path, exist_ok=True)
        st = self._store
        named = {n: ((k, v) for k, v, _ in st._iter_tree(r[_ROOT])) for n, r in st.recs.items()}
        write_packed(os.path.join(path, "data.mdb"), named, st.map_size, st.txnid, main_plain=dict(st.plain))
        builtins.open(os.path.join(path, "lock.mdb"), "ab").close()

    def close(self) -> None:
        if not self._closed:
            try:
                self._flush()
            finally:
                self._closed = True
                self._store.close()
                if self._lock_fh:
                    self._lock_fh.close()              # releases the flock
                    self._lock_fh = None

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
