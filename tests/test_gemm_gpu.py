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
@pytest.mark.parametrize("bn", ["64", "128", "192", "256"])
def test_all_tile_widths(bn, ncta):
    """Every (cluster size, N tile) instantiation, incl. an M tail inside a CTA pair."""
    import torch
    if bn == "64" and ncta == "2":
        pytest.skip("the 64-wide tile exists for single-CTA tiles only (M <= 128 row blocks)")
    from clipb200 import _native
    g = torch.Generator(device="cuda").manual_seed(int(bn))
    M, N, K = 1280 + 77, 768, 768
    A = (torch.randn((M, K), generator=g, device="cuda") * 0.5).half()
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).half()
    with _native.tuning(gemm_bn=int(bn), gemm_ncta=int(ncta)):
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


@pytest.mark.parametrize("outlier", [0.0, 60.0])
@pytest.mark.parametrize("M,W,epi", [(12800, 768, EPI_BIAS), (1280 + 50, 768, EPI_BIAS_GELU), (77 * 9, 512, EPI_BIAS),
                                     (50, 768, EPI_BIAS_GELU), (77, 512, EPI_BIAS), (128, 768, EPI_BIAS)])
def test_layernorm_folded_into_gemm(M, W, epi, outlier):
    """Producer GEMM (residual epilogue) emits per-row statistics of x; consumer GEMM applies
    LayerNorm in its epilogue from gamma-folded weights.  Reference: LN(x) @ W^T + b in fp32."""
    import torch
    from clipb200 import _native as N
    g = torch.Generator(device="cuda").manual_seed(M + W)
    L = N.lib()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    # producer: x = x0 + att @ Wo^T + bo, plus statistics
    att = (torch.randn((M, W), generator=g, device="cuda") * 0.5).half()
    Wo = (torch.randn((W, W), generator=g, device="cuda") * W ** -0.5).half()
    bo = torch.randn((W,), generator=g, device="cuda") * 0.1
    x = (torch.randn((M, W), generator=g, device="cuda") * 2 + 0.7)
    x[:, 5] += outlier          # real CLIP residual streams carry a few huge-magnitude channels
    x[:, 301] -= outlier * 0.5
    x = x.half()
    x_ref = (x.float() + att.float() @ Wo.float().T + bo).half()
    slices = L.cb_gemm_out_slices(M, W)
    stats = torch.zeros((M, slices, 2), device="cuda")
    N.check(L.cb_gemm_f16_ex_device(M, W, W, p(att), p(Wo), p(bo), p(x), p(x), EPI_BIAS_RESID, None, 0, None,
                                    p(stats), st))
    torch.cuda.synchronize()
    assert ((x.float() - x_ref.float()).abs() <= 2e-2 + 1e-3 * x_ref.float().abs()).all()
    tot = stats.sum(dim=1)
    # statistics are taken from the fp32 values just before the fp16 store: they differ from the
    # stored row by its rounding noise (<= 2^-10 relative on the sum of squares, worst case)
    assert torch.allclose(tot[:, 0], x.float().sum(1), atol=1e-1, rtol=2e-4)
    assert torch.allclose(tot[:, 1], (x.float() ** 2).sum(1), atol=2e-1, rtol=1.5e-3)
    # consumer: y = LN(x) @ W1^T + b1 (optionally QuickGELU), LayerNorm folded
    Nn = 3 * W
    W1 = torch.randn((Nn, W), generator=g, device="cuda") * W ** -0.5
    b1 = torch.randn((Nn,), generator=g, device="cuda") * 0.1
    gamma = 1 + 0.1 * torch.randn((W,), generator=g, device="cuda")
    beta = 0.1 * torch.randn((W,), generator=g, device="cuda")
    Wp = (W1 * gamma[None, :]).half()
    colsum = Wp.float().sum(1).contiguous()
    bp = (b1 + W1 @ beta).contiguous()
    y = torch.empty((M, Nn), dtype=torch.float16, device="cuda")
    N.check(L.cb_gemm_f16_ex_device(M, Nn, W, p(x), p(Wp), p(bp), None, p(y), epi, p(stats), slices, p(colsum),
                                    None, st))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.float(), (W,), gamma, beta, 1e-5) @ W1.T + b1
    if epi == EPI_BIAS_GELU:
        ref = ref * torch.sigmoid(1.702 * ref)
    err = (y.float() - ref).abs()
    assert err.max().item() <= 3e-2 and err.mean().item() <= 3e-3, (err.max().item(), err.mean().item())


@pytest.mark.parametrize("split", [-1, 1, 2, 4, 8])
@pytest.mark.parametrize("M,N,K,epi", [(50, 2304, 768, EPI_BIAS), (77, 2048, 512, EPI_BIAS_GELU), (1, 768, 3072, EPI_BIAS_RESID),
                                       (128, 512, 2048, EPI_BIAS_RESID), (98, 768, 3072, EPI_PATCH), (3, 512, 768, EPI_F32)])
def test_single_row_block_split_k(M, N, K, epi, split):
    """M <= 128 (a single query): the cluster split-K kernel, every epilogue, every split factor (split = -1:
    the library's own choice): against fp32 torch, bit-identical across repeated launches (partials are
    summed in rank order), and within fp16 rounding of the general kernel (gemm_skinny = 0)."""
    import torch
    from clipb200 import _native
    g = torch.Generator(device="cuda").manual_seed(M * 31 + N + K + epi)
    A = (torch.randn((M, K), generator=g, device="cuda") * 0.5).half()
    W = (torch.randn((N, K), generator=g, device="cuda") * K ** -0.5).half()
    b = torch.randn((N,), generator=g, device="cuda") if epi in (EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESID) else None
    pos = torch.randn((50, N), generator=g, device="cuda") if epi == EPI_PATCH else None
    x0 = torch.randn((M, N), generator=g, device="cuda").half() if epi == EPI_BIAS_RESID else None
    rows = M + M // 49 + 1 if epi == EPI_PATCH else M

    def run():
        out = torch.zeros((rows, N), dtype=torch.float32 if epi == EPI_F32 else torch.float16, device="cuda")
        x = x0.clone() if x0 is not None else None
        return _gemm(torch, A, W, bias=b, resid=x, pos=pos, epi=epi, out=x if x is not None else out)

    acc = A.float() @ W.float().T
    if epi in (EPI_BIAS, EPI_BIAS_GELU):
        ref = acc + b
        if epi == EPI_BIAS_GELU:
            ref = ref * torch.sigmoid(1.702 * ref)
    elif epi == EPI_BIAS_RESID:
        ref = acc + b + x0.float()
    elif epi == EPI_PATCH:
        ref = torch.zeros((rows, N), device="cuda")
        r = torch.arange(M, device="cuda")
        ref[r + r // 49 + 1] = acc + pos[1 + r % 49]
    else:
        ref = acc
    with _native.tuning(gemm_skinny=split):
        got = run()
        again = run()
    assert torch.equal(got, again), "split-K result depends on the order the CTAs finished"
    _close(torch, got, ref, tol=4e-3 if epi == EPI_BIAS_GELU else (1e-4 if epi == EPI_F32 else 2e-3))
    with _native.tuning(gemm_skinny=0):
        general = run()
    assert (got.float() - general.float()).abs().max().item() <= 2e-3 * max(1.0, ref.abs().max().item())
