"""faiss index-file codec for the two index kinds CLI-P touches.

The reference writes `images.index` with `faiss.write_index(IndexIVFFlat)` (build-index.py:109)
and opens it with `faiss.read_index` (query-index.py:29).  This module reads and writes that
on-disk format so index files travel in both directions between faiss and clipb200:

  "IxFI" / "IxF2"  IndexFlat (inner product / L2)
  "IwFl"           IndexIVFFlat: ivf header, nested quantizer index, direct map, "ilar"
                   array inverted lists ("full" or "sprs" size table), per list codes then ids

Layout restated from faiss's published serialiser (faiss/impl/index_write.cpp,
index_read.cpp; faiss is NOT vendored by the reference, setup.sh:12-18 installs HEAD, and it
is not installable here - so this codec is **parity unpinned**: checked against a byte-level
fixture written by hand from the format description and by round trips, never against faiss
itself).  Everything is little-endian; size_t/idx_t are 8 bytes:

  index header   int32 d | int64 ntotal | int64 dummy(1<<20) | int64 dummy | u8 is_trained |
                 int32 metric_type | [float32 metric_arg if metric_type > 1]
  IxFI           fourcc | header | u64 count_of_floats | float32[count]
  IwFl           fourcc | header | u64 nlist | u64 nprobe | <quantizer index> |
                 u8 direct_map_type | u64 n | int64[n] | [u64 m | (int64,int64)[m] if type == 2] |
                 "ilar" | u64 nlist | u64 code_size | "full" u64 nlist u64[nlist] sizes
                                                    | "sprs" u64 2m u64[2m] (list, size) pairs |
                 for every non-empty list: u8[size*code_size] codes, int64[size] ids

Pure numpy, no GPU: rows are handed over in chunks so 20 GB files stream through.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Callable, Iterator, List, Optional, Tuple

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
_DUMMY = 1 << 20
_MAX_VEC = 1 << 40          # faiss's own sanity bound in READVECTOR


class FaissFormatError(RuntimeError):
    pass


@dataclass
class FlatSection:
    d: int
    ntotal: int
    metric: int
    is_trained: bool
    data_offset: int                 # byte offset of the float32 rows in the file


@dataclass
class IVFSection:
    d: int
    ntotal: int
    metric: int
    is_trained: bool
    nlist: int
    nprobe: int
    quantizer: FlatSection
    code_size: int
    # per non-empty list: (list number, size, byte offset of its codes, byte offset of its ids)
    lists: List[Tuple[int, int, int, int]] = field(default_factory=list)


class _Reader:
    def __init__(self, mm: np.ndarray, path: str):
        self.mm, self.pos, self.path = mm, 0, path

    def take(self, n: int) -> bytes:
        if self.pos + n > self.mm.shape[0]:
            raise FaissFormatError(f"{self.path}: truncated index file (wanted {n} bytes at offset {self.pos})")
        b = self.mm[self.pos:self.pos + n].tobytes()
        self.pos += n
        return b

    def skip(self, n: int) -> int:
        if n < 0 or self.pos + n > self.mm.shape[0]:
            raise FaissFormatError(f"{self.path}: truncated index file (wanted {n} bytes at offset {self.pos})")
        at = self.pos
        self.pos += n
        return at

    def unpack(self, fmt: str):
        return struct.unpack("<" + fmt, self.take(struct.calcsize("<" + fmt)))

    def u64(self) -> int:
        return self.unpack("Q")[0]

    def vec_size(self) -> int:
        n = self.u64()
        if n >= _MAX_VEC:
            raise FaissFormatError(f"{self.path}: implausible vector length {n} at offset {self.pos - 8}")
        return n

    def fourcc(self) -> bytes:
        return self.take(4)


def _read_header(r: _Reader):
    d, ntotal, _d0, _d1, trained, metric = r.unpack("iqqqBi")
    if d <= 0 or ntotal < 0:
        raise FaissFormatError(f"{r.path}: bad index header (d={d}, ntotal={ntotal})")
    if metric > 1:
        r.take(4)                    # metric_arg
    return d, ntotal, metric, bool(trained)


def _read_flat(r: _Reader, cc: bytes) -> FlatSection:
    d, ntotal, metric, trained = _read_header(r)
    count = r.vec_size()
    if count != ntotal * d:
        raise FaissFormatError(f"{r.path}: {cc.decode()} holds {count} floats, header says {ntotal} x {d}")
    off = r.skip(count * 4)
    return FlatSection(d, ntotal, metric, trained, off)


def _read_section(r: _Reader):
    cc = r.fourcc()
    if cc in (b"IxFI", b"IxF2", b"IxFl"):
        return _read_flat(r, cc)
    if cc == b"IwFl":
        d, ntotal, metric, trained = _read_header(r)
        nlist, nprobe = r.unpack("QQ")
        quant = _read_section(r)
        if not isinstance(quant, FlatSection):
            raise FaissFormatError(f"{r.path}: IVF quantizer is not a flat index")
        dm_type = r.unpack("B")[0]
        r.skip(r.vec_size() * 8)                         # direct-map array
        if dm_type == 2:
            r.skip(r.vec_size() * 16)                    # hashtable pairs
        ilcc = r.fourcc()
        sec = IVFSection(d, ntotal, metric, trained, nlist, nprobe, quant, d * 4)
        if ilcc == b"il00":
            return sec
        if ilcc != b"ilar":
            raise FaissFormatError(f"{r.path}: inverted-list kind {ilcc!r} is not supported (only 'ilar')")
        il_nlist, code_size = r.unpack("QQ")
        if il_nlist != nlist or code_size != d * 4:
            raise FaissFormatError(f"{r.path}: inverted lists (nlist={il_nlist}, code_size={code_size}) do not "
                                   f"match an IVFFlat of nlist={nlist}, d={d}")
        kind = r.fourcc()
        n = r.vec_size()
        raw = np.frombuffer(r.take(n * 8), dtype="<u8")
        if kind == b"full":
            if n != nlist:
                raise FaissFormatError(f"{r.path}: 'full' size table has {n} entries for {nlist} lists")
            sizes = [(i, int(s)) for i, s in enumerate(raw) if s]
        elif kind == b"sprs":
            if n % 2:
                raise FaissFormatError(f"{r.path}: odd 'sprs' size table")
            sizes = [(int(raw[2 * j]), int(raw[2 * j + 1])) for j in range(n // 2) if raw[2 * j + 1]]
        else:
            raise FaissFormatError(f"{r.path}: unknown list-size encoding {kind!r}")
        for (li, sz) in sizes:
            c_off = r.skip(sz * code_size)
            i_off = r.skip(sz * 8)
            sec.lists.append((li, sz, c_off, i_off))
        if sum(s for _, s, _, _ in sec.lists) != ntotal:
            raise FaissFormatError(f"{r.path}: inverted lists hold {sum(s for _, s, _, _ in sec.lists)} vectors, "
                                   f"header says {ntotal}")
        return sec
    raise FaissFormatError(f"{r.path}: index type {cc!r} is not supported (IxFI, IxF2, IwFl)")


@dataclass
class ParsedIndex:
    """What read_index needs: the kind, the header fields and the rows in id order."""
    kind: str                        # "flat" | "ivf"
    d: int
    ntotal: int
    metric: int
    nlist: int
    nprobe: int
    _mm: np.ndarray
    _section: object
    _list_of_id: Optional[np.ndarray] = None     # ivf only: index into _section.lists holding id i
    _pos_of_id: Optional[np.ndarray] = None      # ivf only: position of id i inside that list

    def rows(self, lo: int, hi: int) -> np.ndarray:
        """float32 (hi-lo, d) rows of ids lo..hi-1 (a copy, C-contiguous)."""
        assert 0 <= lo <= hi <= self.ntotal
        d = self.d
        if self.kind == "flat":
            off = self._section.data_offset + lo * d * 4
            return np.frombuffer(self._mm, dtype="<f4", count=(hi - lo) * d, offset=off).reshape(hi - lo, d).astype(np.float32)
        out = np.empty((hi - lo, d), dtype=np.float32)
        which, pos = self._list_of_id[lo:hi], self._pos_of_id[lo:hi]
        for li in np.unique(which):                 # consecutive ids are scattered over the lists
            sel = which == li
            _n, sz, c_off, _i = self._section.lists[int(li)]
            codes = np.frombuffer(self._mm, dtype="<f4", count=sz * d, offset=c_off).reshape(sz, d)
            out[sel] = codes[pos[sel]]
        return out

    def iter_rows(self, step: int = 1 << 15) -> Iterator[np.ndarray]:
        for lo in range(0, self.ntotal, step):
            yield self.rows(lo, min(self.ntotal, lo + step))


def parse(path: str) -> ParsedIndex:
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    r = _Reader(mm, path)
    sec = _read_section(r)
    if isinstance(sec, FlatSection):
        return ParsedIndex("flat", sec.d, sec.ntotal, sec.metric, 0, 0, mm, sec)
    # flatten the inverted lists back into id (= add) order
    list_of_id = np.full(sec.ntotal, -1, dtype=np.int32)
    pos_of_id = np.zeros(sec.ntotal, dtype=np.int64)
    for k, (_li, sz, _c_off, i_off) in enumerate(sec.lists):
        ids = np.frombuffer(mm, dtype="<i8", count=sz, offset=i_off)
        if sz and (ids.min() < 0 or ids.max() >= sec.ntotal):
            raise FaissFormatError(f"{path}: ids outside 0..ntotal-1 (add_with_ids indexes are not supported; "
                                   f"CLI-P uses sequential add, build-index.py:99,107)")
        list_of_id[ids] = k
        pos_of_id[ids] = np.arange(sz, dtype=np.int64)
    if sec.ntotal and list_of_id.min() < 0:
        raise FaissFormatError(f"{path}: inverted lists do not cover every id in 0..{sec.ntotal - 1}")
    return ParsedIndex("ivf", sec.d, sec.ntotal, sec.metric, sec.nlist, sec.nprobe, mm, sec, list_of_id, pos_of_id)


# ---- writing -----------------------------------------------------------------------

def _header(d: int, ntotal: int, metric: int, trained: bool = True) -> bytes:
    assert metric in (METRIC_INNER_PRODUCT, METRIC_L2)
    return struct.pack("<iqqqBi", d, ntotal, _DUMMY, _DUMMY, 1 if trained else 0, metric)


def _flat_cc(metric: int) -> bytes:
    return b"IxFI" if metric == METRIC_INNER_PRODUCT else b"IxF2"


def _write_rows(fh, ntotal: int, d: int, get_rows: Callable[[int, int], np.ndarray], step: int, on_chunk=None) -> None:
    for lo in range(0, ntotal, step):
        hi = min(ntotal, lo + step)
        rows = np.ascontiguousarray(get_rows(lo, hi), dtype="<f4")
        assert rows.shape == (hi - lo, d)
        if on_chunk is not None:
            on_chunk(rows)
        fh.write(rows.tobytes())


def write_flat(path: str, d: int, ntotal: int, get_rows: Callable[[int, int], np.ndarray],
               metric: int = METRIC_INNER_PRODUCT, step: int = 1 << 16) -> None:
    """IndexFlat file: get_rows(lo, hi) returns float32 (hi-lo, d) in id order."""
    with open(path, "wb") as fh:
        fh.write(_flat_cc(metric) + _header(d, ntotal, metric) + struct.pack("<Q", ntotal * d))
        _write_rows(fh, ntotal, d, get_rows, step)


def write_ivf_single_list(path: str, d: int, ntotal: int, get_rows: Callable[[int, int], np.ndarray],
                          nprobe: int = 1, metric: int = METRIC_INNER_PRODUCT, step: int = 1 << 16) -> None:
    """IndexIVFFlat file whose search is exact in faiss too: nlist = 1, one inverted list
    holding every row with ids 0..ntotal-1, quantizer = one centroid (the mean row).

    clipb200 serves IVF by an exact scan (north star), so it has no k-means partition to
    write; a one-list IVF is the faithful on-disk form of that.  nprobe is stored as given (the
    reference sets 32, query-index.py:30; faiss clamps it to nlist when it searches).
    """
    nlist = 1
    with open(path, "wb") as fh:
        fh.write(b"IwFl" + _header(d, ntotal, metric) + struct.pack("<QQ", nlist, max(1, int(nprobe))))
        quant_at = fh.tell()
        centroid = np.zeros((1, d), dtype="<f4")
        fh.write(_flat_cc(metric) + _header(d, nlist, metric) + struct.pack("<Q", nlist * d) + centroid.tobytes())
        cent_data_at = fh.tell() - d * 4
        fh.write(struct.pack("<BQ", 0, 0))                               # direct map: NoMap, empty array
        fh.write(b"ilar" + struct.pack("<QQ", nlist, d * 4))
        if ntotal > 0:                                                   # faiss: "full" iff n_non0 > nlist / 2
            fh.write(b"full" + struct.pack("<QQ", nlist, ntotal))
        else:
            fh.write(b"sprs" + struct.pack("<Q", 0))
        acc = np.zeros(d, dtype=np.float64)

        def on_chunk(rows):
            acc[...] += rows.sum(axis=0, dtype=np.float64)

        _write_rows(fh, ntotal, d, get_rows, step, on_chunk)
        for lo in range(0, ntotal, 1 << 20):
            fh.write(np.arange(lo, min(ntotal, lo + (1 << 20)), dtype="<i8").tobytes())
        if ntotal > 0:
            fh.seek(cent_data_at)
            fh.write((acc / ntotal).astype("<f4").tobytes())
        assert quant_at > 0
