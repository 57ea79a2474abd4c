"""Seeded synthetic inputs shared by tests and bench."""
from __future__ import annotations

import numpy as np


def unit_rows(n: int, d: int = 512, seed: int = 0, clip_like: bool = False) -> np.ndarray:
    """n L2-normalised fp32 rows.  clip_like adds a shared direction (real CLIP
    embeddings are anisotropic), which packs scores tightly and stresses top-k."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    if clip_like:
        u = np.random.default_rng(99).standard_normal(d).astype(np.float32)
        u /= np.linalg.norm(u)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        x = 0.7 * u[None, :] + x
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


def device_unit_rows(n: int, d: int, seed: int, device, dtype, chunk: int = 1 << 20):
    """Same recipe generated on the GPU in chunks (10M x 512 never exists in fp32)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n, d), dtype=dtype, device=device)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        x = torch.randn((m, d), generator=g, device=device, dtype=torch.float32)
        x = x / x.norm(dim=1, keepdim=True)
        out[lo:lo + m] = x.to(dtype)
    return out


def synthetic_tokens(n: int, seed: int = 0):
    """[49406, t_1..t_m, 49407, 0...] rows with m in [3,20], t in [1000,40000):
    stand-in for clip.tokenize output while no BPE vocabulary is available offline."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros((n, 77), dtype=torch.int32)
    m = torch.randint(3, 21, (n,), generator=g)
    body = torch.randint(1000, 40000, (n, 20), generator=g, dtype=torch.int32)
    for i in range(n):
        k = int(m[i])
        ids[i, 0] = 49406
        ids[i, 1:1 + k] = body[i, :k]
        ids[i, 1 + k] = 49407
    return ids
