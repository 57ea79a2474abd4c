"""GPU numerics: the tcgen05 GEMM against a plain PyTorch fp32 reference of the same op
(fp16 inputs up-cast to fp32).  Tolerance: fp32 accumulation, fp16 output rounding ->
|err| <= 2e-3 * max(1, |ref|)."""
import ctypes as C

import pytest

pytestmark = pytest.mark.gpu

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESID, EPI_PATCH, EPI_F32 = range(5)


def _gemm(torch, A, W, bias=None, resid=None, pos=None, epi=EPI_BIAS, out=None, out_rows=None):
    from clipb200 import _native as N
    M, K = A.shape
    Nn = W.shape[0]
    if out is None:
        out = torch.empty((out_rows or M, Nn), dtype=torch.float32 if epi == EPI_F32 else torch.float16, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    stream = torch.cuda.current_stream().cuda_stream
    N.check(N.lib().cb_gemm_f16_device(M, Nn, K, p(A), p(W), p(bias), p(resid), p(pos), p(out), Nn, epi,
                                       C.c_void_p(stream)))
    torch.cuda.synchronize()
    return out


def _close(torch, got, ref, tol=2e-3):
    err = (got.float() - ref).abs()
    lim = tol * ref.abs().clamp_min(1.0)
    bad = (err > lim).sum().item()
    assert bad == 0, f"{bad} / {ref.numel()} elements off; max err {err.max().item():.4g} (ref max {ref.abs().max().item():.3g})"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 128), (256, 768, 768), (12800, 768, 768),
                                   (12800, 2304, 768), (12800, 3072, 768), (12800, 768, 3072),
                                   (77, 512, 512), (1000, 1536, 512), (256, 512, 768)])
def test_bias_epilogue(M, N, K):
    import torch
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    A = (torch.randn((M, K), generator=g, device="cuda") * 0.5).half()
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).half()
    b = torch.randn((N,), generator=g, device="cuda")
    out = _gemm(torch, A, W, bias=b)
    ref = A.float() @ W.float().T + b
    _close(torch, out, ref)


@pytest.mark.parametrize("ncta", ["1", "2"])
@pytest.mark.parametrize("bn", ["128", "192", "256"])
def test_all_tile_widths(bn, ncta, monkeypatch):
    """Every (cluster size, N tile) instantiation, incl. an M tail inside a CTA pair."""
    import torch
    monkeypatch.setenv("CLIPB200_GEMM_BN", bn)
    monkeypatch.setenv("CLIPB200_GEMM_NCTA", ncta)
    g = torch.Generator(device="cuda").manual_seed(int(bn))
    M, N, K = 1280 + 77, 768, 768
    A = (torch.randn((M, K), generator=g, device="cuda") * 0.5).half()
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).half()
    out = _gemm(torch, A, W, bias=None)
    _close(torch, out, A.float() @ W.float().T)


def test_quickgelu_epilogue():
    import torch
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 1280, 3072, 768
    A = (torch.randn((M, K), generator=g, device="cuda")).half()
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).half()
    b = torch.randn((N,), generator=g, device="cuda")
    out = _gemm(torch, A, W, bias=b, epi=EPI_BIAS_GELU)
    x = A.float() @ W.float().T + b
    _close(torch, out, x * torch.sigmoid(1.702 * x))


def test_residual_epilogue_in_place():
    import torch
    g = torch.Generator(device="cuda").manual_seed(2)
    M, N, K = 2560, 768, 3072
    A = (torch.randn((M, K), generator=g, device="cuda") * 0.3).half()
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).half()
    b = torch.randn((N,), generator=g, device="cuda")
    x = torch.randn((M, N), generator=g, device="cuda").half()
    ref = x.float() + A.float() @ W.float().T + b
    out = _gemm(torch, A, W, bias=b, resid=x, epi=EPI_BIAS_RESID, out=x)
    assert out.data_ptr() == x.data_ptr()
    _close(torch, out, ref)


def test_patch_embed_epilogue():
    import torch
    g = torch.Generator(device="cuda").manual_seed(3)
    B = 6
    M, N, K = B * 49, 768, 3072
    A = (torch.randn((M, K), generator=g, device="cuda")).half()
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).half()
    pos = torch.randn((50, N), generator=g, device="cuda")
    out = torch.zeros((B * 50, N), dtype=torch.float16, device="cuda")
    _gemm(torch, A, W, pos=pos, epi=EPI_PATCH, out=out)
    ref = (A.float() @ W.float().T).view(B, 49, N) + pos[1:][None]
    got = out.view(B, 50, N)
    _close(torch, got[:, 1:], ref)
    assert (got[:, 0] == 0).all(), "class-token rows must not be touched by the patch GEMM"


def test_f32_projection_epilogue():
    import torch
    g = torch.Generator(device="cuda").manual_seed(4)
    M, N, K = 256, 512, 768
    A = (torch.randn((M, K), generator=g, device="cuda")).half()
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).half()
    out = _gemm(torch, A, W, epi=EPI_F32)
    assert out.dtype == torch.float32
    _close(torch, out, A.float() @ W.float().T, tol=1e-4)


def test_back_to_back_launches_are_deterministic():
    import torch
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 12800, 768, 768
    A = (torch.randn((M, K), generator=g, device="cuda") * 0.5).half()
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).half()
    a = _gemm(torch, A, W).clone()
    for _ in range(3):
        assert torch.equal(_gemm(torch, A, W), a)
