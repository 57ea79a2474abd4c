"""ORACLE (test infrastructure, not product code) -- CLIP ViT-B/32 forward in fp32.

CPU restatement of what `model.encode_image` / `model.encode_text` compute when the
reference calls them at /root/reference/build-index.py:49 and
/root/reference/query-index.py:108, plus the L2 normalisation at build-index.py:50 and
query-index.py:13-17, and the `_transform` normalisation step feeding build-index.py:48.

The arithmetic lives in openai/CLIP (`clip/model.py`, `clip/clip.py`), an un-vendored,
un-pinned dependency (setup.sh:22 clones HEAD).  Its published architecture is restated
here functionally over a state dict with the upstream parameter names:

  vision: conv1 (32x32 stride-32 patch embed, no bias) -> [class token ; patches] +
          positional_embedding -> ln_pre -> 12 x residual block -> ln_post(token 0) @ proj
  text  : token_embedding[ids] + positional_embedding -> 12 x residual block with an
          additive causal mask -> ln_final -> row at argmax(ids) @ text_projection
  block : x += out_proj(softmax(q k^T / sqrt(64)) v) over ln_1(x);  x += c_proj(quickgelu(c_fc(ln_2(x))))
          quickgelu(x) = x * sigmoid(1.702 x); LayerNorm eps 1e-5; heads = width / 64.

PARITY UNPINNED against the reference itself (it has no tests or fixtures and neither
`clip` nor its weights are installable offline).  It IS pinned against an independent
implementation of the same architecture, `transformers.CLIPModel`, by
tests/golden/make_golden.py -> tests/golden/clip_golden.npz (max |diff| ~1e-6).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may import
this module.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

MEAN = (0.48145466, 0.4578275, 0.40821073)
STD = (0.26862954, 0.26130258, 0.27577711)


def preprocess_u8(images_u8_hwc: torch.Tensor) -> torch.Tensor:
    """[B,224,224,3] uint8 -> [B,3,224,224] fp32: ToTensor (/255) then Normalize(mean, std).
    For 224x224 inputs CLIP's Resize(224)+CenterCrop(224) is the identity."""
    x = images_u8_hwc.permute(0, 3, 1, 2).float() / 255.0
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    return (x - mean) / std


def _block(x: torch.Tensor, sd: Dict[str, torch.Tensor], pre: str, heads: int, mask) -> torch.Tensor:
    B, L, W = x.shape
    hd = W // heads
    h = F.layer_norm(x, (W,), sd[f"{pre}.ln_1.weight"], sd[f"{pre}.ln_1.bias"], 1e-5)
    qkv = F.linear(h, sd[f"{pre}.attn.in_proj_weight"], sd[f"{pre}.attn.in_proj_bias"])
    q, k, v = qkv.split(W, dim=-1)
    q = q.view(B, L, heads, hd).transpose(1, 2)
    k = k.view(B, L, heads, hd).transpose(1, 2)
    v = v.view(B, L, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if mask is not None:
        s = s + mask
    a = torch.softmax(s, dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, L, W)
    x = x + F.linear(a, sd[f"{pre}.attn.out_proj.weight"], sd[f"{pre}.attn.out_proj.bias"])
    h = F.layer_norm(x, (W,), sd[f"{pre}.ln_2.weight"], sd[f"{pre}.ln_2.bias"], 1e-5)
    m = F.linear(h, sd[f"{pre}.mlp.c_fc.weight"], sd[f"{pre}.mlp.c_fc.bias"])
    m = m * torch.sigmoid(1.702 * m)
    return x + F.linear(m, sd[f"{pre}.mlp.c_proj.weight"], sd[f"{pre}.mlp.c_proj.bias"])


@torch.no_grad()
def encode_image(sd: Dict[str, torch.Tensor], x: torch.Tensor, layers: int = 12) -> torch.Tensor:
    """x: [B,3,224,224] fp32 (what `transform` returns) -> [B,512] fp32, un-normalised."""
    B = x.shape[0]
    W = sd["visual.conv1.weight"].shape[0]
    p = F.conv2d(x, sd["visual.conv1.weight"], None, stride=32)          # [B,768,7,7]
    p = p.reshape(B, W, -1).permute(0, 2, 1)                              # [B,49,768]
    cls = sd["visual.class_embedding"].view(1, 1, W).expand(B, 1, W)
    t = torch.cat([cls, p], dim=1) + sd["visual.positional_embedding"]
    t = F.layer_norm(t, (W,), sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"], 1e-5)
    for i in range(layers):
        t = _block(t, sd, f"visual.transformer.resblocks.{i}", W // 64, None)
    c = F.layer_norm(t[:, 0, :], (W,), sd["visual.ln_post.weight"], sd["visual.ln_post.bias"], 1e-5)
    return c @ sd["visual.proj"]


@torch.no_grad()
def encode_text(sd: Dict[str, torch.Tensor], ids: torch.Tensor, layers: int = 12) -> torch.Tensor:
    """ids: [B,77] integer tokens -> [B,512] fp32, un-normalised."""
    B, L = ids.shape
    W = sd["token_embedding.weight"].shape[1]
    t = sd["token_embedding.weight"][ids.long()] + sd["positional_embedding"][:L]
    mask = torch.full((L, L), float("-inf")).triu_(1)
    for i in range(layers):
        t = _block(t, sd, f"transformer.resblocks.{i}", W // 64, mask)
    t = F.layer_norm(t, (W,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    eot = ids.long().argmax(dim=-1)
    return t[torch.arange(B), eot] @ sd["text_projection"]


def l2_normalize_rows(x: torch.Tensor) -> torch.Tensor:
    """build-index.py:50: x / x.norm(dim=-1, keepdim=True) (no epsilon)."""
    return x / x.norm(dim=-1, keepdim=True)


def normalize_query(v):
    """query-index.py:13-17: whole-array norm; returned unchanged when norm < 1e-9."""
    import numpy as np
    n = np.linalg.norm(v)
    return v if n < 1e-9 else v / n


def synthetic_tokens(n: int, seed: int = 0) -> torch.Tensor:
    """[49406, t_1..t_m, 49407, 0...] with m in [3,20], t in [1000,40000):
    the BPE vocabulary is not available offline."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros((n, 77), dtype=torch.int32)
    for i in range(n):
        m = int(torch.randint(3, 21, (1,), generator=g))
        ids[i, 0] = 49406
        ids[i, 1:1 + m] = torch.randint(1000, 40000, (m,), generator=g, dtype=torch.int32)
        ids[i, 1 + m] = 49407
    return ids


def to_hf_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Key mapping openai/CLIP -> transformers.CLIPModel (used only to pin this oracle
    against that independent implementation)."""
    out = {}

    def blocks(src, dst, L):
        for i in range(L):
            s, d = f"{src}.resblocks.{i}", f"{dst}.encoder.layers.{i}"
            W = sd[f"{s}.ln_1.weight"].shape[0]
            wq, wk, wv = sd[f"{s}.attn.in_proj_weight"].split(W, 0)
            bq, bk, bv = sd[f"{s}.attn.in_proj_bias"].split(W, 0)
            out[f"{d}.self_attn.q_proj.weight"], out[f"{d}.self_attn.q_proj.bias"] = wq, bq
            out[f"{d}.self_attn.k_proj.weight"], out[f"{d}.self_attn.k_proj.bias"] = wk, bk
            out[f"{d}.self_attn.v_proj.weight"], out[f"{d}.self_attn.v_proj.bias"] = wv, bv
            out[f"{d}.self_attn.out_proj.weight"] = sd[f"{s}.attn.out_proj.weight"]
            out[f"{d}.self_attn.out_proj.bias"] = sd[f"{s}.attn.out_proj.bias"]
            for a, b in (("ln_1", "layer_norm1"), ("ln_2", "layer_norm2")):
                out[f"{d}.{b}.weight"] = sd[f"{s}.{a}.weight"]
                out[f"{d}.{b}.bias"] = sd[f"{s}.{a}.bias"]
            for a, b in (("c_fc", "fc1"), ("c_proj", "fc2")):
                out[f"{d}.mlp.{b}.weight"] = sd[f"{s}.mlp.{a}.weight"]
                out[f"{d}.mlp.{b}.bias"] = sd[f"{s}.mlp.{a}.bias"]

    blocks("visual.transformer", "vision_model", 12)
    blocks("transformer", "text_model", 12)
    out["vision_model.embeddings.class_embedding"] = sd["visual.class_embedding"]
    out["vision_model.embeddings.patch_embedding.weight"] = sd["visual.conv1.weight"]
    out["vision_model.embeddings.position_embedding.weight"] = sd["visual.positional_embedding"]
    out["vision_model.pre_layrnorm.weight"] = sd["visual.ln_pre.weight"]
    out["vision_model.pre_layrnorm.bias"] = sd["visual.ln_pre.bias"]
    out["vision_model.post_layernorm.weight"] = sd["visual.ln_post.weight"]
    out["vision_model.post_layernorm.bias"] = sd["visual.ln_post.bias"]
    out["visual_projection.weight"] = sd["visual.proj"].T.contiguous()
    out["text_model.embeddings.token_embedding.weight"] = sd["token_embedding.weight"]
    out["text_model.embeddings.position_embedding.weight"] = sd["positional_embedding"]
    out["text_model.final_layer_norm.weight"] = sd["ln_final.weight"]
    out["text_model.final_layer_norm.bias"] = sd["ln_final.bias"]
    out["text_projection.weight"] = sd["text_projection"].T.contiguous()
    out["logit_scale"] = sd["logit_scale"]
    return out
