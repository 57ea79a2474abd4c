"""ctypes binding of libclipb200.so (the C ABI declared in include/clipb200.h).

There is deliberately no fallback: if the shared library is missing or fails to
load, importing a symbol raises -- the product path never routes through the CPU
oracle (that lives under oracle/ and is test infrastructure only).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CLIPB200_LIB: load another build of the same library (profiles/ uses the -DCLIPB200_EXPERIMENTS variant)
LIB_PATH = os.environ.get("CLIPB200_LIB") or os.path.join(_HERE, "libclipb200.so")

CB_OK, CB_ERR_INVALID, CB_ERR_CUDA, CB_ERR_NOGPU, CB_ERR_OOM, CB_ERR_IO = range(6)
CB_F32, CB_F16 = 0, 1


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libclipb200 error {code}: {msg}")
        self.code = code


_lib = None

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int

# name -> (restype, argtypes); mirrors include/clipb200.h one to one
SIGNATURES = {
    "cb_last_error": (C.c_char_p, []),
    "cb_abi_version": (_int, []),
    "cb_device_count": (_int, [C.POINTER(_int)]),
    "cb_launch_count": (_i64, [_int]),
    "cb_tuning_set": (_int, [C.c_char_p, _i64]),
    "cb_tuning_get": (_int, [C.c_char_p, C.POINTER(_i64)]),
    "cb_flatip_device": (_int, [_p]),
    "cb_flatip_p2p_init": (_int, [_p, _int, _int, _i64, _p]),
    "cb_flatip_p2p_connect": (_int, [_p, _p]),
    "cb_flatip_search_p2p_device": (_int, [_p, _i64, _p, _i64, _p, _p, _i64, _p]),
    "cb_flatip_p2p_status": (_int, [_p, C.POINTER(_int)]),
    "cb_flatip_search_p2p": (_int, [_p, _i64, _p, _i64, _p, _p, _i64]),
    "cb_flatip_submit_search_device": (_int, [_p, _i64, _p, _i64, _p, _p, _i64, _p]),
    "cb_flatip_join": (_int, [_p, _p]),
    "cb_sharded_create": (_int, [_int, _int, _int, C.POINTER(_int), C.POINTER(_p)]),
    "cb_sharded_free": (None, [_p]),
    "cb_sharded_ntotal": (_i64, [_p]),
    "cb_sharded_num_shards": (_int, [_p]),
    "cb_sharded_shard": (_p, [_p, _int]),
    "cb_sharded_reserve": (_int, [_p, _i64]),
    "cb_sharded_reset": (_int, [_p]),
    "cb_sharded_add": (_int, [_p, _i64, _p]),
    "cb_sharded_add_device": (_int, [_p, _int, _i64, _p, _int, _p]),
    "cb_sharded_search": (_int, [_p, _i64, _p, _i64, _p, _p]),
    "cb_sharded_search_device": (_int, [_p, _i64, _p, _i64, _p, _p, _p]),
    "cb_sharded_get_rows": (_int, [_p, _i64, _i64, _p]),
    "cb_clip_ln_fold_status": (_int, [_p, C.POINTER(_int), C.POINTER(C.c_double)]),
    "cb_flatip_create": (_int, [_int, _int, _int, C.POINTER(_p)]),
    "cb_flatip_free": (None, [_p]),
    "cb_flatip_ntotal": (_i64, [_p]),
    "cb_flatip_dim": (_int, [_p]),
    "cb_flatip_storage_dtype": (_int, [_p]),
    "cb_flatip_reserve": (_int, [_p, _i64]),
    "cb_flatip_add": (_int, [_p, _i64, _p]),
    "cb_flatip_add_device": (_int, [_p, _i64, _p, _int, _p]),
    "cb_flatip_reset": (_int, [_p]),
    "cb_flatip_search": (_int, [_p, _i64, _p, _i64, _p, _p]),
    "cb_flatip_search_device": (_int, [_p, _i64, _p, _i64, _p, _p, _i64, _p]),
    "cb_topk_merge_device": (_int, [_int, _i64, _i64, _p, _p, _i64, _i64, _p, _p, _p]),
    "cb_flatip_get_rows": (_int, [_p, _i64, _i64, _p]),
    "cb_flatip_device_rows": (_p, [_p]),
    "cb_clip_create": (_int, [_int, _int, _int, C.POINTER(_p)]),
    "cb_clip_set_param": (_int, [_p, C.c_char_p, _p, _i64]),
    "cb_clip_finalize": (_int, [_p]),
    "cb_clip_free": (None, [_p]),
    "cb_clip_encode_image_u8_device": (_int, [_p, _i64, _p, _p, _int, _p]),
    "cb_clip_encode_image_f32_device": (_int, [_p, _i64, _p, _p, _int, _p]),
    "cb_clip_encode_image_u8": (_int, [_p, _i64, _p, _p, _int]),
    "cb_clip_submit_image_u8": (_int, [_p, _i64, _p, _p, _int]),
    "cb_clip_submit_image_u8_device": (_int, [_p, _i64, _p, _p, _int, _p]),
    "cb_clip_join": (_int, [_p, _p]),
    "cb_clip_sync": (_int, [_p]),
    "cb_clip_encode_text_device": (_int, [_p, _i64, _p, _p, _int, _p]),
    "cb_clip_encode_text": (_int, [_p, _i64, _p, _p, _int]),
    "cb_clip_timing": (_int, [_p, _int]),
    "cb_clip_timing_breakdown": (_int, [_p, C.POINTER(C.c_double)]),
    "cb_clip_timing_launches": (_int, [_p, C.POINTER(C.c_double), C.POINTER(C.c_double), _int, C.POINTER(_int)]),
    "cb_clip_timing_read": (_int, [_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_int)]),
    "cb_layernorm_f16_device": (_int, [_p, _p, _p, _p, _int, _int, _int, _p, _p, _int, _p]),
    "cb_attention_f16_device": (_int, [_p, _p, _int, _int, _int, _int, _p]),
    "cb_preprocess_u8_device": (_int, [_p, _p, _int, _p]),
    "cb_preprocess_f32_device": (_int, [_p, _p, _int, _p]),
    "cb_l2norm_f32_device": (_int, [_p, _p, _int, _int, _p]),
    "cb_resize224_u8_device": (_int, [_p, _int, _int, _p, _p]),
    "cb_jpeg_create": (_int, [_int, _int, C.POINTER(_p)]),
    "cb_jpeg_free": (None, [_p]),
    "cb_jpeg_threads": (_int, [_p]),
    "cb_jpeg_decode_files": (_int, [_p, _i64, C.POINTER(C.c_char_p), _p, _p]),
    "cb_jpeg_decode_memory": (_int, [_p, _i64, C.POINTER(_p), C.POINTER(_i64), _p, _p]),
    "cb_gemm_f16_ex_device": (_int, [_int, _int, _int, _p, _p, _p, _p, _p, _int, _p, _int, _p, _p, _p]),
    "cb_gemm_out_slices": (_int, [_int, _int]),
    "cb_gemm_f16_device": (_int, [_int, _int, _int, _p, _p, _p, _p, _p, _p, _int, _int, _p]),
    "cb_flatip_batch_stats": (_int, [_p, C.POINTER(_i64), C.POINTER(_i64)]),
    "cb_flatip_timing": (_int, [_p, _int]),
    "cb_flatip_timing_read": (_int, [_p, C.POINTER(C.c_double), C.POINTER(_int)]),
    "cb_flatip_phase_times": (_int, [_p, C.POINTER(C.c_double)]),
}


def lib() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python __graft_entry__.py build` "
                "(clipb200 has no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)   # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def last_error() -> str:
    return (lib().cb_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != CB_OK:
        raise NativeError(rc, last_error())


def device_count() -> int:
    n = _int(0)
    rc = lib().cb_device_count(C.byref(n))
    return n.value if rc == CB_OK else 0


def tuning_set(name: str, value: int) -> None:
    """Set a process-wide tuning knob (include/clipb200.h: cb_tuning_set); -1 restores the default."""
    check(lib().cb_tuning_set(name.encode(), int(value)))


def tuning_get(name: str) -> int:
    v = _i64(0)
    check(lib().cb_tuning_get(name.encode(), C.byref(v)))
    return int(v.value)


class tuning:
    """Context manager: `with _native.tuning(gemm_bn=192): ...` (restores the previous values)."""

    def __init__(self, **knobs):
        self.knobs, self.prev = knobs, {}

    def __enter__(self):
        for k, v in self.knobs.items():
            self.prev[k] = tuning_get(k)
            tuning_set(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.prev.items():
            tuning_set(k, v)
        return False


def launch_count(reset: bool = False) -> int:
    return int(lib().cb_launch_count(1 if reset else 0))
