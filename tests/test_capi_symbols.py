"""CPU: the C-ABI library builds, loads and exports every symbol include/clipb200.h
declares; the ctypes table mirrors the header; no compute is called."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "clipb200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib():
    from clipb200 import build
    path = build.build()
    return C.CDLL(path)


def test_header_declares_something():
    syms = declared_symbols()
    assert "cb_flatip_search" in syms and "cb_last_error" in syms


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in clipb200.h but not exported: {missing}"


def test_ctypes_table_matches_header(lib):
    from clipb200 import _native
    assert sorted(_native.SIGNATURES) == declared_symbols()
    assert _native.lib().cb_abi_version() >= 1


def test_no_gpu_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from clipb200 import _native, faiss
    assert _native.device_count() == 0
    with pytest.raises(_native.NativeError) as ei:
        faiss.IndexFlatIP(512)
    assert ei.value.code == _native.CB_ERR_NOGPU
    assert "no CPU fallback" in str(ei.value)


def test_bad_arguments_are_errors_not_crashes(lib):
    from clipb200 import _native
    L = _native.lib()
    h = C.c_void_p()
    assert L.cb_flatip_create(100, 0, 0, C.byref(h)) == _native.CB_ERR_INVALID
    assert "512" in _native.last_error()
    assert L.cb_flatip_create(512, 7, 0, C.byref(h)) == _native.CB_ERR_INVALID
    assert L.cb_flatip_search(None, 1, None, 1, None, None) == _native.CB_ERR_INVALID


def test_sass_is_sm100(lib):
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    from clipb200 import _native
    out = subprocess.run(["cuobjdump", "-lelf", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_tensor_core_kernels_are_tcgen05_tma_tmem(lib):
    """The GEMM and batch-search kernels must really be Blackwell tensor-core code: tcgen05.mma
    (UTCHMMA), TMA loads (UTMALDG), TMEM reads (LDTM), and TMA bulk stores (UTMASTG) in the GEMM
    epilogues that store fp16 tiles.  profiles/r02_sass_digest.txt is this table, committed."""
    import importlib.util
    import shutil
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    spec = importlib.util.spec_from_file_location("sass_digest", os.path.join(ROOT, "profiles", "sass_digest.py"))
    sd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sd)
    d = sd.digest()
    names = sd.demangle(list(d))
    tc = {sd.shorten(names[k]): c for k, c in d.items()
          if "gemm_tcgen05_kernel" in names[k] or "flatip_batch_kernel" in names[k] or "attention_pair_kernel" in names[k]}
    assert len([k for k in tc if k.startswith("gemm_tcgen05_kernel")]) >= 20      # BN x epilogue x NCTA
    assert {"flatip_batch_kernel<1>", "flatip_batch_kernel<2>"} <= set(tc)
    pair = [c for k, c in tc.items() if "attention_pair_kernel" in k]
    assert len(pair) == 1 and pair[0]["UTCHMMA"] >= 12 and pair[0]["UTMASTG"] >= 2   # S (4) + P.V from TMEM (8), TMA stores
    for k, c in tc.items():
        assert c["UTCHMMA"] >= 4 and c["UTMALDG"] >= 2 and c["LDTM"] >= 1 and c["UTCBAR"] >= 2, (k, dict(c))
        assert c["HMMA"] == 0, f"{k} fell back to legacy mma.sync"
    stores = [c["UTMASTG"] for k, c in tc.items() if k.startswith("gemm_tcgen05_kernel")]
    assert sum(1 for x in stores if x >= 1) >= 12
    # the debug probes are compiled out of the product library
    from clipb200 import _native
    assert _native.lib().cb_tuning_set(b"gemm_debug", 1) == _native.CB_ERR_INVALID
    assert _native.lib().cb_tuning_set(b"skip", 1) == _native.CB_ERR_INVALID
    assert _native.lib().cb_tuning_set(b"gemm_bn", -1) == _native.CB_OK


def test_reference_import_lines_resolve_to_clipb200():
    """`import lmdb, clip, faiss` (build-index.py:5-8, query-index.py:6-9) with cli-p_b200/ first on
    PYTHONPATH must bind to this package, in a fresh interpreter, without a GPU."""
    import subprocess
    import sys
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "cli-p_b200"))
    code = ("import lmdb, clip, faiss; "
            "print(lmdb.open.__module__, clip.load.__module__, faiss.IndexFlatIP.__module__, "
            "faiss.METRIC_INNER_PRODUCT, callable(faiss.read_index), callable(clip.tokenize))")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/")
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["clipb200.lmdb", "clipb200.clip", "clipb200.faiss", "0", "True", "True"]
