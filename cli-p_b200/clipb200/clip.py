"""Drop-in for the subset of `import clip` (openai/CLIP) that CLI-P uses, backed by
libclipb200's hand-written sm_100a kernels.

Reference call sites (under /root/reference):
  model, transform = clip.load("ViT-B/32", device=device, jit=False)   build-index.py:18, query-index.py:21
  model.eval()                                                          build-index.py:20, query-index.py:23
  transform(PIL.Image) -> FloatTensor[3,224,224]                        build-index.py:48
  model.encode_image(FloatTensor[B,3,224,224]) -> Tensor[B,512]         build-index.py:49
  clip.tokenize([str]) -> IntTensor[n,77]                               query-index.py:107
  model.encode_text(IntTensor[n,77]) -> Tensor[n,512]                   query-index.py:108

The model always runs on a CUDA device (there is no CPU path: `device="cpu"`, which
query-index.py:20 forces, is mapped to cuda:0 and the result tensor is returned on the
CPU so the caller's `.detach().cpu().numpy()` chain is unchanged).  Besides the
reference's fp32 NCHW input, encode_image also accepts uint8 [B,224,224,3] batches
(host or device), for which ToTensor+Normalize run on the GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import List, Sequence, Union

import numpy as np
import torch

from . import _native as N
from . import weights as _weights

_MEAN = (0.48145466, 0.4578275, 0.40821073)
_STD = (0.26862954, 0.26130258, 0.27577711)


def available_models() -> List[str]:
    return ["ViT-B/32"]


def weights_fingerprint(state_dict) -> str:
    """Identity of a checkpoint: sha256 over the two projection matrices and the class embedding (fp32 bytes).
    The indexer stamps it into vectors.lmdb so that rows embedded with different weights (for example the
    seeded test weights and the real checkpoint) cannot end up in one store unnoticed."""
    import hashlib
    h = hashlib.sha256()
    for name in ("visual.proj", "visual.class_embedding", "text_projection"):
        h.update(state_dict[name].detach().to(torch.float32).cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


class CLIPB200:
    """Duck type of openai/CLIP's `CLIP` module for the two calls the reference makes."""

    def __init__(self, state_dict, device: int = 0, max_image_batch: int = 256, max_text_batch: int = 64,
                 result_device: Union[str, torch.device, None] = None):
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self.result_device = torch.device(result_device) if result_device is not None else self.device
        self.max_image_batch, self.max_text_batch = max_image_batch, max_text_batch
        self.handle = C.c_void_p()
        N.check(N.lib().cb_clip_create(self.device_index, max_image_batch, max_text_batch, C.byref(self.handle)))
        shapes = _weights.param_shapes()
        for name, shape in shapes.items():
            if name == "logit_scale":
                continue
            t = state_dict[name].detach().to(torch.float32).cpu().contiguous()
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{name}: shape {tuple(t.shape)} != {shape}")
            N.check(N.lib().cb_clip_set_param(self.handle, name.encode(), C.c_void_p(t.data_ptr()), t.numel()))
        N.check(N.lib().cb_clip_finalize(self.handle))
        self.logit_scale = state_dict.get("logit_scale", torch.tensor(2.6592))
        self.weights_id = weights_fingerprint(state_dict)

    def close(self):
        if getattr(self, "handle", None):
            N.lib().cb_clip_free(self.handle)
            self.handle = None

    def ln_fold_status(self):
        """(folded, min cosine of the calibration batch, -2.0 if the ln_fold knob decided)."""
        f, c = C.c_int(0), C.c_double(0)
        N.check(N.lib().cb_clip_ln_fold_status(self.handle, C.byref(f), C.byref(c)))
        return bool(f.value), float(c.value)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # nn.Module surface the reference touches
    def eval(self):
        return self

    def float(self):
        return self

    def to(self, *a, **k):
        return self

    @property
    def dtype(self):
        return torch.float32

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def encode_image(self, image: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """[B,3,224,224] float (what `transform` returns) or [B,224,224,3] uint8 -> [B,512] float32."""
        if image.dtype == torch.uint8:
            assert image.dim() == 4 and tuple(image.shape[1:]) == (224, 224, 3), \
                f"uint8 input must be [B,224,224,3], got {tuple(image.shape)}"
            fn = N.lib().cb_clip_encode_image_u8_device
        else:
            assert image.dim() == 4 and tuple(image.shape[1:]) == (3, 224, 224), \
                f"float input must be [B,3,224,224], got {tuple(image.shape)}"
            image = image.to(torch.float32)
            fn = N.lib().cb_clip_encode_image_f32_device
        with torch.cuda.device(self.device):
            x = image.to(self.device, non_blocking=True).contiguous()
            out = torch.empty((x.shape[0], 512), dtype=torch.float32, device=self.device)
            N.check(fn(self.handle, x.shape[0], C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
                       1 if normalize else 0, self._stream()))
            x.record_stream(torch.cuda.current_stream(self.device))
        return out if self.result_device == self.device else out.to(self.result_device)

    def encode_text(self, text: torch.Tensor, normalize: bool = False) -> torch.Tensor:
        """[n,77] integer tokens -> [n,512] float32."""
        assert text.dim() == 2 and text.shape[1] == 77, f"tokens must be [n,77], got {tuple(text.shape)}"
        with torch.cuda.device(self.device):
            ids = text.to(self.device, dtype=torch.int32, non_blocking=True).contiguous()
            out = torch.empty((ids.shape[0], 512), dtype=torch.float32, device=self.device)
            N.check(N.lib().cb_clip_encode_text_device(self.handle, ids.shape[0], C.c_void_p(ids.data_ptr()),
                                                       C.c_void_p(out.data_ptr()), 1 if normalize else 0,
                                                       self._stream()))
        return out if self.result_device == self.device else out.to(self.result_device)

    # host-buffer entry points (numpy in / numpy out; H2D and D2H inside the call)
    def encode_image_u8_host(self, images: np.ndarray, normalize: bool = True, out: np.ndarray = None) -> np.ndarray:
        assert images.dtype == np.uint8 and images.ndim == 4 and images.shape[1:] == (224, 224, 3)
        images = np.ascontiguousarray(images)
        if out is None:
            out = np.empty((images.shape[0], 512), np.float32)
        N.check(N.lib().cb_clip_encode_image_u8(self.handle, images.shape[0], C.c_void_p(images.ctypes.data),
                                                C.c_void_p(out.ctypes.data), 1 if normalize else 0))
        return out

    def encode_image_batches_host(self, batches, outs=None, normalize: bool = True):
        """Index-time streaming entry point: `batches` is a sequence of uint8 [b,224,224,3]
        arrays/tensors in (preferably pinned) host memory, b <= max_image_batch.  The H2D copy
        of batch i+1 overlaps the forward pass of batch i.  Returns the list of [b,512] float32
        results (host memory), complete on return."""
        L = N.lib()
        keep, results = [], []
        for i, b in enumerate(batches):
            if torch.is_tensor(b):
                assert b.dtype == torch.uint8 and not b.is_cuda and b.is_contiguous()
                ptr, n = b.data_ptr(), b.shape[0]
            else:
                b = np.ascontiguousarray(b)
                assert b.dtype == np.uint8
                ptr, n = b.ctypes.data, b.shape[0]
            assert tuple(b.shape[1:]) == (224, 224, 3)
            if outs is not None:
                o = outs[i]
            else:
                o = torch.empty((n, 512), dtype=torch.float32).pin_memory()
            optr = o.data_ptr() if torch.is_tensor(o) else o.ctypes.data
            keep.append((b, o))
            N.check(L.cb_clip_submit_image_u8(self.handle, n, C.c_void_p(ptr), C.c_void_p(optr), 1 if normalize else 0))
            results.append(o)
        N.check(L.cb_clip_sync(self.handle))
        return results

    def encode_text_host(self, ids: np.ndarray, normalize: bool = True) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        assert ids.ndim == 2 and ids.shape[1] == 77
        out = np.empty((ids.shape[0], 512), np.float32)
        N.check(N.lib().cb_clip_encode_text(self.handle, ids.shape[0], C.c_void_p(ids.ctypes.data),
                                            C.c_void_p(out.ctypes.data), 1 if normalize else 0))
        return out


from .pil_transform import image_to_u8, resize_center_crop  # noqa: E402,F401  (torch-free home: decode workers import it)


def resize_center_crop_device(image_u8: torch.Tensor) -> torch.Tensor:
    """GPU form of resize_center_crop for a uint8 RGB tensor [h, w, 3] on a CUDA device ->
    [224, 224, 3] uint8, bit-identical to the Pillow path (cb_resize224_u8_device)."""
    assert image_u8.is_cuda and image_u8.dtype == torch.uint8 and image_u8.dim() == 3 and image_u8.shape[2] == 3
    src = image_u8.contiguous()
    out = torch.empty((224, 224, 3), dtype=torch.uint8, device=src.device)
    with torch.cuda.device(src.device):
        stream = C.c_void_p(torch.cuda.current_stream(src.device).cuda_stream)
        N.check(N.lib().cb_resize224_u8_device(C.c_void_p(src.data_ptr()), src.shape[0], src.shape[1],
                                               C.c_void_p(out.data_ptr()), stream))
    return out


def _transform(n_px: int = 224):
    """clip._transform on the CPU, as the reference does (build-index.py:48)."""
    mean = torch.tensor(_MEAN).view(3, 1, 1)
    std = torch.tensor(_STD).view(3, 1, 1)

    def transform(image):
        x = torch.from_numpy(image_to_u8(image, n_px).copy()).permute(2, 0, 1).float().div(255.0)
        return (x - mean) / std

    return transform


def load(name: str = "ViT-B/32", device: Union[str, torch.device] = "cuda", jit: bool = False,
         download_root: str = None, max_image_batch: int = 256, max_text_batch: int = 64):
    """clip.load(...) -> (model, transform).  Weights come from $CLIP_WEIGHTS (an OpenAI
    ViT-B-32.pt TorchScript archive or state_dict).  The reference downloads the checkpoint here;
    offline that is impossible, and embedding a photo collection with random weights would fill
    vectors.lmdb with rows that resume-by-key never recomputes -- so without a checkpoint this
    raises.  CLIPB200_SYNTHETIC_WEIGHTS=1 (tests, benches) opts into the seeded synthetic model."""
    if name != "ViT-B/32":
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    dev = torch.device(device)
    path = os.environ.get("CLIP_WEIGHTS")
    if path:
        sd = _weights.load_state_dict(path)
    elif os.environ.get("CLIPB200_SYNTHETIC_WEIGHTS", "0") == "1":
        print("clipb200: CLIPB200_SYNTHETIC_WEIGHTS=1 -- seeded synthetic ViT-B/32 weights (embeddings are "
              "meaningless for real images)", file=sys.stderr)
        sd = _weights.synthetic_state_dict(0)
    else:
        raise RuntimeError(
            "clip.load: set CLIP_WEIGHTS to an OpenAI ViT-B-32.pt (TorchScript archive or state_dict); no "
            "checkpoint can be downloaded offline.  CLIPB200_SYNTHETIC_WEIGHTS=1 opts into seeded synthetic "
            "weights for tests and benches.")
    index = dev.index if (dev.type == "cuda" and dev.index is not None) else (
        torch.cuda.current_device() if dev.type == "cuda" else 0)
    model = CLIPB200(sd, device=index, max_image_batch=max_image_batch, max_text_batch=max_text_batch,
                     result_device=dev if dev.type == "cpu" else None)
    return model, _transform(224)


def tokenize(texts: Union[str, Sequence[str]], context_length: int = 77, truncate: bool = False) -> torch.Tensor:
    """clip.tokenize: [sot] + bpe(text) + [eot], zero padded to 77."""
    from . import bpe
    if isinstance(texts, str):
        texts = [texts]
    tok = bpe.default_tokenizer()
    sot, eot = tok.encoder["<|startoftext|>"], tok.encoder["<|endoftext|>"]
    out = torch.zeros((len(texts), context_length), dtype=torch.int32)
    for i, t in enumerate(texts):
        ids = [sot] + tok.encode(t) + [eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = eot
        out[i, :len(ids)] = torch.tensor(ids, dtype=torch.int32)
    return out
