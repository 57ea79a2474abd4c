"""CLIP ViT-B/32 parameter shapes (openai/CLIP state-dict key names) and seeded
synthetic weights.

The real checkpoint (`ViT-B-32.pt`, fetched by clip.load at
/root/reference/build-index.py:18) is not available offline; a user can point
CLIP_WEIGHTS at it (TorchScript archive or plain state_dict) and the same keys
load.  Tests and benches use `synthetic_state_dict`.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import torch

VISION = dict(width=768, layers=12, heads=12, patch=32, res=224, tokens=50)
TEXT = dict(width=512, layers=12, heads=8, ctx=77, vocab=49408)
EMBED_DIM = 512


def param_shapes() -> "OrderedDict[str, Tuple[int, ...]]":
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    vw, tw = VISION["width"], TEXT["width"]
    s["visual.conv1.weight"] = (vw, 3, 32, 32)
    s["visual.class_embedding"] = (vw,)
    s["visual.positional_embedding"] = (50, vw)
    s["visual.ln_pre.weight"] = (vw,)
    s["visual.ln_pre.bias"] = (vw,)
    for pre, w, L in (("visual.transformer", vw, VISION["layers"]), ("transformer", tw, TEXT["layers"])):
        for i in range(L):
            b = f"{pre}.resblocks.{i}"
            s[f"{b}.ln_1.weight"] = (w,)
            s[f"{b}.ln_1.bias"] = (w,)
            s[f"{b}.attn.in_proj_weight"] = (3 * w, w)
            s[f"{b}.attn.in_proj_bias"] = (3 * w,)
            s[f"{b}.attn.out_proj.weight"] = (w, w)
            s[f"{b}.attn.out_proj.bias"] = (w,)
            s[f"{b}.ln_2.weight"] = (w,)
            s[f"{b}.ln_2.bias"] = (w,)
            s[f"{b}.mlp.c_fc.weight"] = (4 * w, w)
            s[f"{b}.mlp.c_fc.bias"] = (4 * w,)
            s[f"{b}.mlp.c_proj.weight"] = (w, 4 * w)
            s[f"{b}.mlp.c_proj.bias"] = (w,)
    s["visual.ln_post.weight"] = (vw,)
    s["visual.ln_post.bias"] = (vw,)
    s["visual.proj"] = (vw, EMBED_DIM)
    s["token_embedding.weight"] = (TEXT["vocab"], tw)
    s["positional_embedding"] = (TEXT["ctx"], tw)
    s["ln_final.weight"] = (tw,)
    s["ln_final.bias"] = (tw,)
    s["text_projection"] = (tw, EMBED_DIM)
    s["logit_scale"] = ()
    return s


def synthetic_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 CPU tensors; scales follow openai/CLIP's initialisation so activations stay
    O(1) through 12 layers; LayerNorm gains/biases and Linear biases are perturbed so
    every fused epilogue term is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def n(shape, std):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std

    for name, shape in param_shapes().items():
        w = shape[-1] if shape else 1
        if name == "logit_scale":
            t = torch.tensor(2.6592)
        elif name == "visual.conv1.weight":
            t = n(shape, (3 * 32 * 32) ** -0.5)
        elif name in ("visual.class_embedding", "visual.positional_embedding", "visual.proj"):
            t = n(shape, 768 ** -0.5)
        elif name == "token_embedding.weight":
            t = n(shape, 0.02)
        elif name == "positional_embedding":
            t = n(shape, 0.01)
        elif name == "text_projection":
            t = n(shape, 512 ** -0.5)
        elif ".ln_" in name or name.startswith(("ln_final", "visual.ln_")):
            t = 1.0 + n(shape, 0.1) if name.endswith("weight") else n(shape, 0.1)
        elif name.endswith("in_proj_weight"):
            t = n(shape, w ** -0.5)
        elif name.endswith(("out_proj.weight", "c_proj.weight")):
            width = shape[0]
            t = n(shape, (width ** -0.5) * (24 ** -0.5) * (1.0 if name.endswith("out_proj.weight") else 0.5))
        elif name.endswith("c_fc.weight"):
            t = n(shape, (2 * w) ** -0.5)
        elif name.endswith("bias"):
            t = n(shape, 0.02)
        else:
            raise KeyError(name)
        sd[name] = t
    return sd


def load_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """User-supplied OpenAI checkpoint: TorchScript archive or pickled state_dict."""
    try:
        m = torch.jit.load(path, map_location="cpu")
        sd = m.state_dict()
    except RuntimeError:
        sd = torch.load(path, map_location="cpu")
        if hasattr(sd, "state_dict"):
            sd = sd.state_dict()
    want = param_shapes()
    out = {}
    for k, shape in want.items():
        if k not in sd:
            raise KeyError(f"{path}: missing parameter {k}")
        t = sd[k].detach().float().cpu()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{path}: {k} has shape {tuple(t.shape)}, expected {shape} (only ViT-B/32 is supported)")
        out[k] = t.contiguous()
    return out
